"""On-device span masking and utterance mixing against the reference's rules.

The reference samples with numpy's RNG on the host (WavLM/WavLM.py:35-159; utterance_mixing_dataset.py:373-438); the device versions
use the library's counter-based generator, so parity is (i) EXACT for every deterministic rule -- span counts, span length, masked
frames inside the valid region, equal masked count per row, the mixing arithmetic given the drawn plan -- and (ii) STATISTICAL for the
random part: the masked fraction and the distribution of span starts are compared with the bit-exact host port of the reference
sampler (`unispeech_b200.masking`, itself pinned to the reference by tests/golden/mask_indices.npz)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_span_mask_rules_and_statistics(cuda_device):
    from unispeech_b200.datapath import span_mask_device
    from unispeech_b200.masking import compute_mask_indices
    dev = cuda_device
    B, T, p, L = 48, 749, 0.65, 10
    lengths = torch.randint(300, T + 1, (B,), generator=torch.Generator().manual_seed(1))
    lengths[0] = T
    pad = torch.arange(T)[None, :] >= lengths[:, None]
    fracs, starts_hist = [], np.zeros(10)
    for seed in range(6):
        m = span_mask_device(B, T, dev, p, L, min_masks=2, padding_mask=pad.to(dev), seed=seed).cpu()
        assert m.dtype == torch.bool and m.shape == (B, T)
        assert not (m & pad).any()                                   # never inside the padded tail
        per_row = m.sum(1)
        assert int(per_row.min()) == int(per_row.max()) > 0          # every row trimmed to the same number of masked frames
        fracs.append(per_row[0].item() / lengths.min().item())
        # span starts of the longest row (before trimming they are uniform on [0, sz - L)); after trimming the masked frames
        # still spread uniformly: decile histogram of masked positions
        pos = torch.nonzero(m[0]).squeeze(1).numpy()
        starts_hist += np.histogram(pos, bins=10, range=(0, T))[0]
    # host port of the reference sampler on the same batch geometry
    np.random.seed(0)
    ref_fracs = []
    for _ in range(6):
        r = compute_mask_indices((B, T), pad, p, L, "static", 0, min_masks=2)
        assert int(r.sum(1).min()) == int(r.sum(1).max())
        ref_fracs.append(r[0].sum() / lengths.min().item())
    assert abs(np.mean(fracs) - np.mean(ref_fracs)) < 0.05 * np.mean(ref_fracs), (np.mean(fracs), np.mean(ref_fracs))
    h = starts_hist / starts_hist.sum()
    assert h.min() > 0.06 and h.max() < 0.14, h                      # ~uniform over the utterance
    # determinism: same seed, same mask; different seed, different mask
    a = span_mask_device(B, T, dev, p, L, 2, pad.to(dev), seed=3)
    b = span_mask_device(B, T, dev, p, L, 2, pad.to(dev), seed=3)
    c = span_mask_device(B, T, dev, p, L, 2, pad.to(dev), seed=4)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_span_mask_untrimmed_counts_follow_the_reference_law(cuda_device):
    """Without padding every row has sz = T and the reference draws ONE span count: all rows hold count spans of L frames whose
    union (overlaps allowed) is what gets masked; count = floor(p T / L + u) takes the two values around p T / L."""
    from unispeech_b200 import ops
    from unispeech_b200 import dropout as DR
    dev = cuda_device
    B, T, p, L = 32, 999, 0.8, 10
    seen = set()
    for seed in range(8):
        mask = torch.empty(B, T, dtype=torch.uint8, device=dev)
        counts = torch.empty(B, dtype=torch.int32, device=dev)
        ops.span_mask(None, B, T, p, L, 2, DR.site_key(seed, 0x7F000003), mask, counts)
        torch.cuda.synchronize()
        c = counts.cpu()
        k = p * T / L
        # union of `count` spans of length L: at most count * L frames, and (overlaps) not much less than the independent-draw mean
        exp_cov = T * (1 - (1 - L / (T - L)) ** int(k))
        assert c.max().item() <= (int(k) + 1) * L
        assert abs(c.float().mean().item() - exp_cov) < 0.06 * exp_cov, (c.float().mean().item(), exp_cov)
        seen.add(int(round(c.float().mean().item())))
    assert len(seen) > 1


def test_mixing_matches_reference_arithmetic(cuda_device):
    """Given the SAME drawn plan, the device mixer reproduces the reference arithmetic (chunk from utterance c scaled to the drawn
    SNR relative to utterance i's power, then re-normalisation of the mixed utterances) on a batch where no source utterance was
    itself mixed before it is used (so the in-place order of the reference cannot matter)."""
    from unispeech_b200.datapath import draw_mix_plan, mix_utterances_device
    dev = cuda_device
    B, T = 6, 48000
    g = torch.Generator().manual_seed(5)
    src = torch.randn(B, T, generator=g) * torch.tensor([0.5, 1.0, 2.0, 0.1, 1.5, 0.7])[:, None]
    np.random.seed(11)
    plan = draw_mix_plan(B, T, mixing_prob=0.6, mixing_max_len=-1)
    assert any(p_[0] >= 0 for p_ in plan) and any(p_[0] < 0 for p_ in plan)
    # reference arithmetic on the host (utterance_mixing_dataset.py:415-435), reading the ORIGINAL batch
    want = src.clone()
    for i, (c, n, cs, ss, snr) in enumerate(plan):
        if c < 0:
            continue
        ref_pow, mix_pow = np.mean(src[i].numpy() ** 2), np.mean(src[c].numpy() ** 2)
        scale = (ref_pow / (mix_pow * 10 ** (snr / 10))) ** 0.5
        want[i, ss:ss + n] += src[c, cs:cs + n] * scale
        want[i] = torch.nn.functional.layer_norm(want[i], want[i].shape)
    got = mix_utterances_device(src.to(dev), plan, normalize=True).cpu()
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4), (got - want).abs().max().item()
    # the plan generator consumes numpy's RNG in the reference's order: the bit-exact host port draws the same plan
    from unispeech_b200.mixing import mix_utterances
    np.random.seed(11)
    host = mix_utterances(src.clone(), mixing_prob=0.6, mixing_num=1, mixing_max_len=-1, normalize=True)
    unmixed = [i for i, p_ in enumerate(plan) if p_[0] < 0]
    assert all(torch.equal(host[i], src[i]) for i in unmixed)         # same utterances left alone => same draws
