"""CPU: the oracle restatement (oracle/wavlm_oracle.py) against fixtures generated from the unmodified reference
modules by tools/make_golden.py.  Bit-level agreement is not expected (different op order), 1e-4 abs is."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import wavlm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    "tiny_postln_ragged": (lambda: O.tiny_config(pre_ln=False), 2, 8000, [8000, 5000]),
    "tiny_preln_ragged": (lambda: O.tiny_config(pre_ln=True), 2, 6400, [6400, 4321]),
    "tiny_postln_nomask": (lambda: O.tiny_config(pre_ln=False), 1, 7777, None),
    "tiny_preln_norelpos": (lambda: O.tiny_config(pre_ln=True, relative_position_embedding=False, gru_rel_pos=False),
                            2, 4000, [4000, 3000]),
    "base2l_halfsec": (lambda: O.base_config(encoder_layers=2), 1, 8000, None),
    "large2l_halfsec": (lambda: O.large_config(encoder_layers=2), 1, 8000, None),
}


def test_all_fixtures_are_covered():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
    names.remove("mask_indices")  # span-sampler fixture, covered by tests/test_api_cpu.py
    assert names == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_matches_reference(name):
    mk, B, L, lengths = CASES[name]
    cfg = mk()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sd = O.deterministic_state_dict(cfg)
    wav, pmask = O.deterministic_waveform(B, L, seed=1, lengths=lengths)
    pm = pmask if lengths is not None else None
    with torch.no_grad():
        conv = O.conv_feature_extractor(sd, wav, cfg)
        res = O.extract_features(sd, wav, cfg, padding_mask=pm)
        res_l = O.extract_features(sd, wav, cfg, padding_mask=pm, output_layer=cfg.encoder_layers)
    assert np.abs(conv.numpy() - g["conv_out"]).max() < 1e-4
    assert np.abs(res["features"].numpy() - g["features"]).max() < 1e-4
    assert np.abs(res["x"].numpy() - g["x_final"]).max() < 2e-4
    lr = np.stack([t.numpy() for t in res_l["layer_results"]])
    assert lr.shape == g["layer_results"].shape
    assert np.abs(lr - g["layer_results"]).max() < 2e-4
    if "frame_padding_mask" in g:
        assert np.array_equal(res["padding_mask"].numpy(), g["frame_padding_mask"])


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if n.startswith("tiny")])
def test_gradients_match_reference(name):
    mk, B, L, lengths = CASES[name]
    cfg = mk()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sd = {k: v.clone().requires_grad_(True) for k, v in O.deterministic_state_dict(cfg).items()}
    wav, pmask = O.deterministic_waveform(B, L, seed=1, lengths=lengths)
    pm = pmask if lengths is not None else None
    mi = torch.from_numpy(g["mask_indices"]) if "mask_indices" in g else None
    res = O.extract_features(sd, wav, cfg, padding_mask=pm, mask_indices=mi)
    assert np.abs(res["x"].detach().numpy() - g["masked_x"]).max() < 2e-4
    loss = O.probe_loss(res["x"], res["padding_mask"], seed=2)
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    norms = dict(zip(g["grad_norm_keys"].tolist(), g["grad_norms"].tolist()))
    for k, n in norms.items():
        got = sd[k].grad.double().norm().item()
        assert abs(got - n) < 2e-3 * max(n, 1e-3) + 1e-5, (k, got, n)
    for key in g.files:
        if key.startswith("grad:"):
            k = key[5:]
            ref = g[key]
            err = np.abs(sd[k].grad.numpy() - ref).max()
            assert err < 1e-3 * max(1.0, np.abs(ref).max()), (k, err)


def test_bucket_lut_is_toeplitz_restatement():
    """bias[h,i,j] = E[bucket(j-i), h]: the LUT form used by the CUDA path equals the dense reference form."""
    T, H = 37, 3
    E = O.hash_uniform("E", (320, H))
    dense = O.compute_bias(T, E, 320, 800)
    lut = O.bucket_lut(T, 320, 800)
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    tab = E[lut]  # [2T-1, H]
    assert torch.equal(dense, tab[(j - i) + T - 1].permute(2, 0, 1))


def test_bucket_values_known_answers():
    """SURVEY.md S7 known answers for num_buckets=320, max_distance=800."""
    d = torch.tensor([0, -1, -79, -80, -100, -400, -800, -5000, 1, 79, 80, 100, 400, 800, 5000])
    b = O.relative_positions_bucket(d, 320, 800).tolist()
    assert b[0] == 0 and b[1] == 1 and b[2] == 79 and b[3] == 80
    assert b[6] == 159 and b[7] == 159
    assert b[8] == 161 and b[9] == 239 and b[10] == 240 and b[11] == 247 and b[12] == 295 and b[13] == 319 and b[14] == 319


def test_frame_counts():
    cfg = O.base_config()
    assert [O.num_frames(16000 * s, cfg) for s in (4, 10, 15, 20, 30)] == [199, 499, 749, 999, 1499]
