"""The oracle's wav2vec 2.0 contrastive head (oracle/wavlm_oracle.py: w2v_sample_negatives, w2v_compute_preds, w2v_contrastive_loss)
against tests/golden/w2v_heads.npz -- numbers produced by executing the reference's own source text for `sample_negatives`,
`compute_preds` and `GumbelVectorQuantizer` (tools/make_w2v_golden.py).  Inputs are hash-generated, so they are rebuilt here."""
import os

import numpy as np
import pytest
import torch

from oracle import wavlm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "w2v_heads.npz")


def _inputs(use_quantizer, n_neg, cross):
    B, T, C, D, Dp = 3, 15, 16, 24, 8
    tag = f"w2v{int(use_quantizer)}{n_neg}{cross}"
    fp = (O.hash_uniform(tag + ".fw", (Dp, D), -0.5, 0.5), O.hash_uniform(tag + ".fb", (Dp,), -0.1, 0.1))
    qd = None
    if use_quantizer:
        groups, num_vars, vq_dim = 2, 4, 12
        qd = dict(weight_proj_w=O.hash_uniform(tag + ".qw", (groups * num_vars, C), -1.0, 1.0),
                  weight_proj_b=O.hash_uniform(tag + ".qb", (groups * num_vars,), -0.1, 0.1),
                  vars_=O.hash_uniform(tag + ".qv", (1, groups * num_vars, vq_dim // groups), 0.0, 1.0), groups=groups, num_vars=num_vars)
        pq = (O.hash_uniform(tag + ".pw", (Dp, vq_dim), -0.5, 0.5), O.hash_uniform(tag + ".pb", (Dp,), -0.1, 0.1))
    else:
        pq = (O.hash_uniform(tag + ".pw", (Dp, C), -0.5, 0.5), O.hash_uniform(tag + ".pb", (Dp,), -0.1, 0.1))
    x_enc = O.hash_uniform(tag + ".x", (B, T, D), -1.0, 1.0)
    unmasked = O.hash_uniform(tag + ".u", (B, T, C), -1.0, 1.0)
    mask = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):
        mask[b, torch.tensor([1 + b, 3 + b, 4 + b, 7 + b, 9 + b, 11 + b])] = True
    return tag, x_enc, unmasked, mask, fp, pq, qd


@pytest.mark.parametrize("use_quantizer,n_neg,cross,seed", [(True, 5, 0, 1), (True, 3, 2, 2), (False, 4, 0, 3)])
def test_w2v_head_oracle_matches_reference_numbers(use_quantizer, n_neg, cross, seed):
    g = np.load(GOLD)
    tag, x_enc, unmasked, mask, fp, pq, qd = _inputs(use_quantizer, n_neg, cross)
    torch.manual_seed(seed)
    got = O.w2v_contrastive_loss(x_enc, unmasked, mask, fp, pq, n_neg, cross, 0.1, quantizer=qd)
    want = torch.from_numpy(g[f"{tag}.logits"])
    fin = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got["logits"]), fin) and int((~fin).sum()) == int(g[f"{tag}.n_neg_is_pos"])
    assert (got["logits"][fin] - want[fin]).abs().max().item() < 1e-5
    assert abs(float(got["loss"]) - float(g[f"{tag}.loss"])) < 1e-4 * abs(float(g[f"{tag}.loss"]))
    assert got["correct"] == int(g[f"{tag}.correct"]) and got["sample_size"] == want.shape[0]
    if use_quantizer:
        assert abs(float(got["q"]["prob_perplexity"]) - float(g[f"{tag}.prob_ppl"])) < 1e-4
        assert abs(float(got["q"]["code_perplexity"]) - float(g[f"{tag}.code_ppl"])) < 1e-4
    # the sampler's index list is the reference's (same generator, same calls)
    torch.manual_seed(seed)
    B = x_enc.shape[0]
    num = int(mask[0].sum())
    _, idx = O.w2v_sample_negatives(torch.zeros(B, num, 4), num, n_neg, cross)
    assert np.array_equal(idx.numpy(), g[f"{tag}.neg_idxs"])
