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


LONG_CASES = {
    "long_postln_T499": (lambda: O.tiny_config(pre_ln=False), 2, 160000, [160000, 101234]),
    "long_preln_T1499": (lambda: O.tiny_config(pre_ln=True), 2, 480000, [480000, 160480]),
}


@pytest.mark.parametrize("name", sorted(LONG_CASES))
def test_long_sequence_rows_match_reference(name):
    """The oracle against the unmodified reference at T = 499 / 1499 frames (tools/make_long_golden.py): ragged batches, every
    layer, a subsample of the frames.  This is the regime of BASELINE.json's configs: the relative-position buckets take the
    LOG branch (|delta| >= 80, WavLM/modules.py:417-443; 319 distinct buckets at T = 1499) and attention spans many key tiles."""
    mk, B, L, lengths = LONG_CASES[name]
    cfg = mk()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sd = O.deterministic_state_dict(cfg)
    wav, pmask = O.deterministic_waveform(B, L, seed=11, lengths=lengths)
    with torch.no_grad():
        res = O.extract_features(sd, wav, cfg, padding_mask=pmask)
        res_l = O.extract_features(sd, wav, cfg, padding_mask=pmask, output_layer=cfg.encoder_layers)
    rows = g["rows"]
    fpm = g["frame_padding_mask"]
    assert np.array_equal(res["padding_mask"].numpy(), fpm)
    valid = ~fpm[:, rows]                                    # [B, rows]
    dx = np.abs(res["x"][:, rows].numpy() - g["x_final"])
    assert dx[valid].max() < 2e-4, dx[valid].max()
    lr = np.stack([t[rows].numpy() for t in res_l["layer_results"]])    # [n+1, rows, B, D]
    assert lr.shape == g["layer_results"].shape
    dl = np.abs(lr - g["layer_results"])
    assert dl[:, valid.T].max() < 2e-4, dl[:, valid.T].max()
    # padded frames too: the reference computes them (garbage in, but deterministic) and so does the restatement
    assert dl.max() < 1e-3, dl.max()


def test_all_fixtures_are_covered():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))
    names.remove("mask_indices")  # span-sampler fixture, covered by tests/test_api_cpu.py
    names.remove("train_heads")   # compute_nce / clip_grad_norm_ / Adam fixture, covered by test_training_heads_match_reference_code
    names.remove("utterance_mixing")  # host data-path fixture, covered by tests/test_api_cpu.py
    names.remove("sat_heads")     # UniSpeech-SAT utterance-contrastive fixture, covered by test_sat_utterance_contrastive_branch_...
    names.remove("w2v_heads")     # wav2vec 2.0 contrastive-head fixture, covered by tests/test_w2v_oracle_cpu.py
    names.remove("vox_real_large2l")  # real-speech ragged fixture, covered by tests/test_vox_real.py
    for n in LONG_CASES:          # long-sequence fixtures (subsampled rows), covered by test_long_sequence_rows_match_reference
        names.remove(n)
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


def test_training_heads_match_reference_code():
    """compute_nce / clip_grad_norm_ / Adam restatements against outputs of the reference's own source text
    (tools/make_head_golden.py executes src/fairseq/models/wavlm/wavlm.py:426-438, utils.py:338-381 and optim/adam.py:100-228)."""
    import torch.nn.functional as F
    g = np.load(os.path.join(GOLD, "train_heads.npz"))
    S, C, Dp = (int(v) for v in g["nce_shape"])
    temp = float(g["nce_temp"])
    proj = O.hash_uniform("g.proj", (S, Dp), -1.0, 1.0)
    E = O.hash_uniform("g.emb", (C, Dp), 0.0, 1.0)
    tgt = (O.hash_uniform("g.tgt", (S,), 0.0, 1.0) * C).long().clamp(max=C - 1)
    logits = O.compute_nce(proj, torch.index_select(E, 0, tgt), E.unsqueeze(1).expand(-1, S, -1), temp)
    want = torch.from_numpy(g["nce_logits"])
    assert torch.equal(torch.isinf(logits), torch.isinf(want))
    fin = ~torch.isinf(want)
    assert (logits[fin] - want[fin]).abs().max().item() < 1e-5
    loss, ss, _ = O.wavlm_criterion([logits], [], 1.0, 0.0)
    assert abs(loss.item() - float(g["nce_loss"])) < 1e-3 and ss == S
    # the fused formulation used by the kernels (plain CE over the C classes) gives the same number
    z = F.normalize(proj, dim=-1) @ F.normalize(E, dim=-1).t() / temp
    assert abs(F.cross_entropy(z, tgt, reduction="sum").item() - float(g["nce_loss"])) < 1e-3
    # optimizer: three steps of clip (max_norm 1.5) + Adam(lr 3e-3, betas (0.9, 0.98), eps 1e-6, weight_decay 0.01)
    shapes = [(5, 7), (3,), (2, 3, 4)]
    params = [O.hash_uniform(f"g.p{i}", s, -1.0, 1.0) for i, s in enumerate(shapes)]
    state = {}
    for step in range(3):
        grads = [O.hash_uniform(f"g.g{step}.{i}", s, -1.0, 1.0) * (0.5 + step) for i, s in enumerate(shapes)]
        norm, coef = O.clip_coefficient(grads, 1.5)
        assert abs(norm.item() - float(g["adam_norms"][step])) < 1e-4
        O.adam_step(params, [x * coef for x in grads], state, 3e-3, (0.9, 0.98), 1e-6, 0.01)
    for i, p in enumerate(params):
        assert torch.allclose(p, torch.from_numpy(g[f"adam_p{i}"]), rtol=1e-5, atol=1e-6), i


def test_sat_utterance_contrastive_branch_matches_reference_code():
    """Oracle groundwork for BASELINE config #4: sample_instances / compute_nce / compute_pred_spk / eval-mode Gumbel quantizer
    restatements against outputs of the reference's own source text (tools/make_sat_golden.py)."""
    g = np.load(os.path.join(GOLD, "sat_heads.npz"))
    B, T, C, Dp, temp = 3, 14, 16, 8, 0.1
    for i, (use_q, n_inst, cross, seed) in enumerate(g["cases"].tolist()):
        tag = f"sat{int(use_q)}{n_inst}{cross}"
        spk_x = O.hash_uniform(tag + ".x", (B, T, C), -1.0, 1.0)
        mask = torch.from_numpy(g[f"mask_{i}"])
        pad = torch.zeros(B, T, dtype=torch.bool)
        sw, sb = O.hash_uniform(tag + ".sw", (Dp, C), -0.5, 0.5), O.hash_uniform(tag + ".sb", (Dp,), -0.1, 0.1)
        quant, pq = None, None
        if use_q:
            groups, num_vars, vq_dim = 2, 5, 12
            quant = dict(weight_proj_w=O.hash_uniform(tag + ".qw", (groups * num_vars, C), -1.0, 1.0),
                         weight_proj_b=O.hash_uniform(tag + ".qb", (groups * num_vars,), -0.1, 0.1),
                         vars_=O.hash_uniform(tag + ".qv", (1, groups * num_vars, vq_dim // groups), 0.0, 1.0), groups=groups,
                         num_vars=num_vars)
            pq = (O.hash_uniform(tag + ".pw", (Dp, vq_dim), -0.5, 0.5), O.hash_uniform(tag + ".pb", (Dp,), -0.1, 0.1))
        torch.manual_seed(int(seed))
        loss, mean_t, acc, q = O.sat_utterance_contrastive_loss(spk_x, pad, mask, sw, sb, int(n_inst), int(cross), temp, quant, pq)
        want = g[f"out_{i}"]
        assert abs(loss.item() - want[0]) < 1e-5, (i, loss.item(), want[0])
        assert abs(mean_t.item() - want[1]) < 1e-6 and abs(acc.item() - want[2]) < 1e-6
        if use_q:
            assert abs(q["code_perplexity"].item() - want[3]) < 1e-4 and abs(q["prob_perplexity"].item() - want[4]) < 1e-4
            assert q["num_vars"] == int(want[5])
