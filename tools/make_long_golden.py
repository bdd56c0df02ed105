"""Long-sequence fixtures: the UNMODIFIED reference (oracle/_ref, see oracle/build_ref.py) at T = 499 and T = 1499 frames,
ragged batches, both LayerNorm orders -- the regime where the relative-position buckets take the LOG branch (|delta| >= 80,
WavLM/modules.py:417-443) and attention spans 4-12 key tiles.  Only a subsample of the frames is stored (every `STEP`-th row of
every layer), which keeps the fixture small while pinning the oracle at full sequence length.

    python -m oracle.build_ref && python tools/make_long_golden.py        (authoring container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from oracle import wavlm_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STEP = 23
CASES = {
    "long_postln_T499": (lambda: O.tiny_config(pre_ln=False), 2, 160000, [160000, 101234]),
    "long_preln_T1499": (lambda: O.tiny_config(pre_ln=True), 2, 480000, [480000, 160480]),
}


def main():
    torch.manual_seed(0)
    for name, (mk, B, L, lengths) in CASES.items():
        cfg = mk()
        sd = O.deterministic_state_dict(cfg)
        m = build_ref.build_model(cfg, sd)
        wav, pmask = O.deterministic_waveform(B, L, seed=11, lengths=lengths)
        with torch.no_grad():
            (x, lr), fpm = m.extract_features(wav.clone(), padding_mask=pmask, mask=False, ret_layer_results=True,
                                              output_layer=cfg.encoder_layers)
            xf, _ = m.extract_features(wav.clone(), padding_mask=pmask, mask=False)
        T = xf.shape[1]
        rows = np.arange(0, T, STEP)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), rows=rows, frame_padding_mask=fpm.numpy(),
                            x_final=xf[:, rows].numpy(), layer_results=np.stack([t[0][rows].numpy() for t in lr]))
        print(name, "T", T, "rows", len(rows), "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__":
    main()
