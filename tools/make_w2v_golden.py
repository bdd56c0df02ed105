"""Pin the oracle's restatement of the wav2vec 2.0 contrastive head to the reference source text (authoring container only).

As tools/make_sat_golden.py: fairseq cannot be imported here, so `Wav2Vec2Model.sample_negatives` / `.compute_preds`
(src/fairseq/models/wav2vec/wav2vec2.py:474-553) and `GumbelVectorQuantizer` (src/fairseq/modules/gumbel_vector_quantizer.py) are
extracted with `ast` and executed unmodified; the forward tail (:621-723) and the infonce criterion
(criterions/wav2vec_criterion.py:44-118) are then applied with plain torch calls in the reference's order.  Only numbers are
committed (tests/golden/w2v_heads.npz)."""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wavlm_oracle as O  # noqa: E402

REF = "/root/reference/src/fairseq"


def find(node, name):
    for sub in ast.walk(node):
        if isinstance(sub, (ast.FunctionDef, ast.ClassDef)) and sub.name == name:
            return sub
    raise KeyError(name)


def source_of(path, *names):
    src = open(path).read()
    node = ast.parse(src)
    for n in names:
        node = find(node, n)
    return textwrap.dedent(ast.get_source_segment(src, node, padded=True))


def build_case(use_quantizer: bool, n_neg: int, cross: int, seed: int):
    B, T, C, D, Dp, temp = 3, 15, 16, 24, 8, 0.1
    w2v = f"{REF}/models/wav2vec/wav2vec2.py"
    ns = {"torch": torch, "F": F, "nn": nn, "buffered_arange": lambda n: torch.arange(n), "is_xla_tensor": lambda t: False,
          "index_put": lambda t, m, v: t.masked_fill(m, v) if not isinstance(v, torch.Tensor) else t.masked_fill(m, float(v))}
    exec(source_of(w2v, "Wav2Vec2Model", "sample_negatives"), ns)
    exec(source_of(w2v, "Wav2Vec2Model", "compute_preds"), ns)
    self = types.SimpleNamespace(n_negatives=n_neg, cross_sample_negatives=cross, logit_temp=temp)
    self.sample_negatives = types.MethodType(ns["sample_negatives"], self)
    self.compute_preds = types.MethodType(ns["compute_preds"], self)
    tag = f"w2v{int(use_quantizer)}{n_neg}{cross}"
    final_proj, gq, qpar = nn.Linear(D, Dp), None, None
    final_proj.weight.data = O.hash_uniform(tag + ".fw", (Dp, D), -0.5, 0.5)
    final_proj.bias.data = O.hash_uniform(tag + ".fb", (Dp,), -0.1, 0.1)
    if use_quantizer:
        gns = {"torch": torch, "nn": nn, "F": F}
        exec(source_of(f"{REF}/modules/gumbel_vector_quantizer.py", "GumbelVectorQuantizer"), gns)
        groups, num_vars, vq_dim = 2, 4, 12   # few codes: equal (positive, negative) pairs DO occur, the -inf branch is exercised
        gq = gns["GumbelVectorQuantizer"](dim=C, num_vars=num_vars, temp=(2.0, 0.5, 0.999), groups=groups, combine_groups=False,
                                          vq_dim=vq_dim, time_first=True)
        gq.weight_proj.weight.data = O.hash_uniform(tag + ".qw", (groups * num_vars, C), -1.0, 1.0)
        gq.weight_proj.bias.data = O.hash_uniform(tag + ".qb", (groups * num_vars,), -0.1, 0.1)
        gq.vars.data = O.hash_uniform(tag + ".qv", (1, groups * num_vars, vq_dim // groups), 0.0, 1.0)
        gq.eval()
        project_q = nn.Linear(vq_dim, Dp)
        qpar = (groups, num_vars, vq_dim)
    else:
        project_q = nn.Linear(C, Dp)
    project_q.weight.data = O.hash_uniform(tag + ".pw", tuple(project_q.weight.shape), -0.5, 0.5)
    project_q.bias.data = O.hash_uniform(tag + ".pb", (Dp,), -0.1, 0.1)
    x_enc = O.hash_uniform(tag + ".x", (B, T, D), -1.0, 1.0)
    unmasked = O.hash_uniform(tag + ".u", (B, T, C), -1.0, 1.0)
    mask = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):
        mask[b, torch.tensor([1 + b, 3 + b, 4 + b, 7 + b, 9 + b, 11 + b])] = True
    # ---- the reference's forward tail (wav2vec2.py:621-723), its own methods doing the work
    with torch.no_grad():
        y = unmasked[mask].view(B, -1, C)
        q = None
        if gq is not None:
            q = gq(y, produce_targets=False)
            y = project_q(q["x"])
        else:
            y = project_q(y)
        torch.manual_seed(seed)
        negs, neg_idxs = self.sample_negatives(y, y.size(1))
        x = final_proj(x_enc[mask].view(B, -1, D))
        logits = self.compute_preds(x, y, negs)
        lg = logits.transpose(0, 2).reshape(-1, logits.size(-1) if False else logits.size(0)).float()   # get_logits (:741-745)
        tgt = lg.new_zeros(lg.size(0), dtype=torch.long)
        loss = F.cross_entropy(lg, tgt, reduction="sum")
        mx, mn = lg.argmax(-1) == 0, lg.argmin(-1) == 0
        out = {"loss": loss.numpy(), "logits": lg.numpy(), "neg_idxs": neg_idxs.numpy(), "correct": np.int64(mx.long().sum() - (mx & mn).long().sum()),
               "n_neg_is_pos": np.int64(torch.isinf(lg).sum())}
        if q is not None:
            out["prob_ppl"] = q["prob_perplexity"].numpy()
            out["code_ppl"] = q["code_perplexity"].numpy()
    # ---- the oracle on the same inputs / seed
    qd = None
    if gq is not None:
        qd = dict(weight_proj_w=gq.weight_proj.weight.data, weight_proj_b=gq.weight_proj.bias.data, vars_=gq.vars.data,
                  groups=qpar[0], num_vars=qpar[1])
    torch.manual_seed(seed)
    got = O.w2v_contrastive_loss(x_enc, unmasked, mask, (final_proj.weight.data, final_proj.bias.data),
                                 (project_q.weight.data, project_q.bias.data), n_neg, cross, temp, quantizer=qd)
    fin = torch.isfinite(torch.from_numpy(out["logits"]))
    assert torch.equal(torch.isfinite(got["logits"]), fin)
    d = (got["logits"][fin] - torch.from_numpy(out["logits"])[fin]).abs().max().item()
    print(f"{tag}: loss ref {float(out['loss']):.6f} oracle {float(got['loss']):.6f}  max |dlogit| {d:.2e}  -inf entries {int(out['n_neg_is_pos'])}")
    assert d < 1e-5 and abs(float(got["loss"]) - float(out["loss"])) < 1e-4 * abs(float(out["loss"])) and got["correct"] == int(out["correct"])
    return tag, out


def main():
    res = {}
    for uq, n, c, seed in ((True, 5, 0, 1), (True, 3, 2, 2), (False, 4, 0, 3)):
        tag, out = build_case(uq, n, c, seed)
        for k, v in out.items():
            res[f"{tag}.{k}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "w2v_heads.npz"), **res)
    print("wrote tests/golden/w2v_heads.npz", {k: np.asarray(v).shape for k, v in res.items()})


if __name__ == "__main__":
    main()
