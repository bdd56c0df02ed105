"""Host-side utterance mixing of the UniSpeech-SAT / WavLM data path, with the reference's semantics AND its numpy RNG consumption
order (src/fairseq/data/audio/utterance_mixing_dataset.py:373-438, utterance branch): seeding `np.random` identically reproduces
the reference's mixed batches bit for bit (tests/test_api_cpu.py against a fixture generated from the reference source text by
tools/make_mixing_golden.py).  Like the span sampler this is collater glue that the reference also runs on the host; it sits
in front of the GPU hot path (BASELINE config #4 applies it to every training batch).

Per utterance i, with probability `mixing_prob`: `mixing_num` times pick another utterance c of the batch (with replacement), a
chunk length c_len <= mixing_max_len, a source window in c and a destination window in i, and add the chunk scaled to a random
SNR in [-5, 5] dB relative to utterance i's current power; then re-normalise utterance i when `normalize` is set.
The noise-corpus branch of the reference (`mixing_noise`, h5py files) is not part of this function.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def mix_utterances(source: torch.Tensor, mixing_prob: float = 0.5, mixing_num: int = 1, mixing_max_len: int = -1,
                   normalize: bool = False) -> torch.Tensor:
    """In place on `source` (float [B, T] on the host), returns it.  `mixing_max_len` < 0: chunks up to T // 2, else T // value."""
    assert source.device.type == "cpu" and source.dim() == 2
    B, T = source.shape
    limit = T // 2 if mixing_max_len < 0 else T // mixing_max_len
    limit = min(limit, T)
    for i in range(B):
        if not (np.random.random() < mixing_prob):
            continue
        for c in np.random.choice(range(B), mixing_num, replace=True):
            c_len = np.random.randint(0, limit + 1)
            c_end = np.random.randint(c_len, T + 1)
            s_end = np.random.randint(c_len, T + 1)
            c_start, s_start = c_end - c_len, s_end - c_len
            ref_pow = np.mean(source[i].numpy() ** 2)
            mix_pow = np.mean(source[c].numpy() ** 2)
            if mix_pow == 0:
                scale = 0
            else:
                snr = np.random.uniform(-5, 5)
                scale = (ref_pow / (mix_pow * 10 ** (snr / 10))) ** 0.5
            source[i, s_start:s_end] += source[c, c_start:c_end].clone() * scale
        if normalize:
            with torch.no_grad():
                source[i] = F.layer_norm(source[i], source[i].shape)
    return source
