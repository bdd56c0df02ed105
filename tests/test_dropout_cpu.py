"""CPU-only checks of the training-mode dropout plumbing: the numpy restatement the parity tests use (oracle.HashDropout)
against the C code the kernels compile (host entry points of csrc/dropout.cuh), the key derivation, and mask statistics."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import wavlm_oracle as O


def _lib():
    from unispeech_b200 import _lib, build
    build.build()
    lib = _lib.load()
    for n in ("b200s_dropout_bits", "b200s_dropout_row_key", "b200s_dropout_threshold16"):
        getattr(lib, n).restype = ctypes.c_uint32
    lib.b200s_attn_dropout_mask_words.restype = ctypes.c_longlong
    return lib


def test_numpy_restatement_matches_compiled_formulas():
    lib = _lib()
    rng = np.random.default_rng(0)
    u32 = ctypes.c_uint32
    for _ in range(200):
        k0, k1, ctr = (int(v) for v in rng.integers(0, 2 ** 32, 3))
        want = lib.b200s_dropout_bits(u32(k0), u32(k1), u32(ctr))
        got = int(O._drop_bits(np.uint64(k0), np.uint64(k1), np.array([ctr], dtype=np.uint64))[0])
        assert got == want
        row = int(rng.integers(0, 2 ** 32))
        r0 = int(O._fmix32((np.uint64(k0) + np.array([row], dtype=np.uint64) * np.uint64(0x9E3779B1)) & O._M32)[0])
        r1 = int(O._fmix32(np.uint64(k1) ^ ((np.array([row], dtype=np.uint64) * np.uint64(0x85EBCA6B)) & O._M32))[0])
        assert r0 == lib.b200s_dropout_row_key(u32(k0), u32(row), 0)
        assert r1 == lib.b200s_dropout_row_key(u32(k1), u32(row), 1)
    for p in (0.0, 0.05, 0.1, 0.25, 0.5, 0.9, 1e-6):
        assert O.drop_threshold16(p) == lib.b200s_dropout_threshold16(ctypes.c_float(p)), p


def test_attention_rows_of_the_restatement_use_the_row_keys():
    """keep_attn(row, j) == half (j & 1) of bits(row_k0, row_k1, j >> 1) evaluated through the compiled formulas."""
    lib = _lib()
    u32 = ctypes.c_uint32
    d = O.HashDropout(99)
    B, H, T, p = 2, 3, 37, 0.3
    keep = d.keep_attn(O.HashDropout.layer_site(1, 3), B, H, T, p)
    k0, k1 = (int(v) for v in d.key(O.HashDropout.layer_site(1, 3)))
    thr = O.drop_threshold16(p)
    rng = np.random.default_rng(1)
    for _ in range(300):
        b, h, i, j = int(rng.integers(B)), int(rng.integers(H)), int(rng.integers(T)), int(rng.integers(T))
        row = (b * H + h) * T + i
        bits = lib.b200s_dropout_bits(u32(lib.b200s_dropout_row_key(u32(k0), u32(row), 0)),
                                      u32(lib.b200s_dropout_row_key(u32(k1), u32(row), 1)), u32(j >> 1))
        half = (bits >> 16) if (j & 1) else (bits & 0xFFFF)
        assert bool(keep[b, h, i, j]) == (half >= thr)


def test_site_keys_match_product_derivation():
    from unispeech_b200 import dropout as DR
    d = O.HashDropout(0x1234_5678_9ABC_DEF0)
    for site in (DR.SITE_INPUT, DR.SITE_ENCODER, DR.layer_site(0, DR.L_DROPOUT1), DR.layer_site(7, DR.L_ATTENTION),
                 DR.layer_site(23, DR.L_DROPOUT3)):
        assert tuple(int(v) for v in d.key(site)) == DR.site_key(0x1234_5678_9ABC_DEF0, site)
    assert DR.layer_site(3, DR.L_ACTIVATION) == O.HashDropout.layer_site(3, 1)
    keys = {DR.site_key(5, s) for s in range(200)}
    assert len(keys) == 200


def test_mask_statistics():
    """Keep rate = 1 - p to sampling error, no correlation between neighbours, sites or seeds."""
    d = O.HashDropout(7)
    for p in (0.05, 0.1, 0.5):
        k = d.keep_rows(4, 512, 768, p).astype(np.float64)
        n = k.size
        assert abs(k.mean() - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n) + 1e-5, (p, k.mean())
        z = (k - k.mean()) / k.std()
        for a, b in ((z[:, :-1], z[:, 1:]), (z[:-1], z[1:]), (z[:, ::2], z[:, 1::2])):
            assert abs((a * b).mean()) < 5 / np.sqrt(a.size)
        k2 = d.keep_rows(5, 512, 768, p).astype(np.float64)
        k3 = O.HashDropout(8).keep_rows(4, 512, 768, p).astype(np.float64)
        for other in (k2, k3):
            assert abs(((other - other.mean()) / other.std() * z).mean()) < 5 / np.sqrt(n)
    a = d.keep_attn(7, 2, 4, 150, 0.1).astype(np.float64)
    assert abs(a.mean() - 0.9) < 4 * np.sqrt(0.09 / a.size)
    za = (a - a.mean()) / a.std()
    assert abs((za[:, :, :-1] * za[:, :, 1:]).mean()) < 5 / np.sqrt(za[:, :, 1:].size)   # adjacent query rows
    assert abs((za[:, :-1] * za[:, 1:]).mean()) < 5 / np.sqrt(za[:, 1:].size)             # adjacent heads
    assert d.keep_rows(4, 8, 64, 0.0).all()


def test_oracle_dropout_is_unbiased_and_scaled_like_f_dropout():
    d = O.HashDropout(3)
    x = torch.ones(4, 50, 64)
    y = d.rows_btc(9, x, 0.25)
    vals = sorted(y.unique().tolist())
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1.0 / 0.75) < 1e-6
    assert abs(y.mean().item() - 1.0) < 0.02
    assert torch.equal(d.rows_tbc(9, x.transpose(0, 1), 0.25).transpose(0, 1), y)
    assert d.rows_btc(9, x, 0.0) is x


def test_mask_word_count_and_training_mode_is_accepted():
    lib = _lib()
    from unispeech_b200 import ops
    from unispeech_b200.dropout import DropState
    from unispeech_b200.wavlm import WavLM, WavLMConfig
    for B, T, H in ((1, 1, 1), (2, 128, 2), (3, 129, 4), (16, 749, 12), (8, 1499, 16)):
        assert lib.b200s_attn_dropout_mask_words(B, T, H) == ops.attn_dropout_mask_words(B, T, H)
    cfg = O.tiny_config(dropout=0.1, attention_dropout=0.1)
    m = WavLM(WavLMConfig(vars(cfg))).train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # (not NotImplementedError: dropout is part of the path)
        m.extract_features(torch.zeros(1, 4000))
    assert DropState.for_model(cfg, False, None) is None
    assert DropState.for_model(O.tiny_config(), True, None) is None
    torch.manual_seed(5)
    a = DropState.for_model(cfg, True, None).seed
    torch.manual_seed(5)
    assert DropState.for_model(cfg, True, None).seed == a
    assert DropState.for_model(cfg, True, 42).seed == 42
    with pytest.raises(ValueError):
        DropState(0, 1.0, 0.0, 0.0, 0.0)
