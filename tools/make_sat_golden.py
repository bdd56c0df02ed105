"""Pin the oracle's restatement of the UniSpeech-SAT utterance-contrastive branch to the reference source text.

Like tools/make_head_golden.py: the fairseq package cannot be imported here, so the relevant functions / classes are extracted
from /root/reference with `ast` and executed unmodified (authoring container only):
  * UniSpeechSATModel.sample_instances, .compute_nce     src/fairseq/models/unispeech_sat/unispeech_sat.py:487-557
  * the nested compute_pred_spk of UniSpeechSATModel.forward                                               :701-736
  * GumbelVectorQuantizer (eval mode)                     src/fairseq/modules/gumbel_vector_quantizer.py
Inputs are hash-generated; torch.manual_seed fixes the instance sampling.  Only numbers are committed (tests/golden/sat_heads.npz).
"""
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


def build_case(use_quantizer: bool, n_inst: int, cross: int, seed: int):
    B, T, C, Dp, temp = 3, 14, 16, 8, 0.1
    sat = f"{REF}/models/unispeech_sat/unispeech_sat.py"
    ns = {"torch": torch, "F": F, "nn": nn, "buffered_arange": lambda n: torch.arange(n)}
    exec(source_of(sat, "UniSpeechSATModel", "sample_instances"), ns)
    exec(source_of(sat, "UniSpeechSATModel", "compute_nce"), ns)
    self = types.SimpleNamespace(n_instances=n_inst, cross_sample_instances=cross, target_glu=None, logit_temp=temp, quantizer=None,
                                 project_q=None)
    self.sample_instances = types.MethodType(ns["sample_instances"], self)
    self.compute_nce = types.MethodType(ns["compute_nce"], self)
    tag = f"sat{int(use_quantizer)}{n_inst}{cross}"
    spk_proj = nn.Linear(C, Dp)
    spk_proj.weight.data = O.hash_uniform(tag + ".sw", (Dp, C), -0.5, 0.5)
    spk_proj.bias.data = O.hash_uniform(tag + ".sb", (Dp,), -0.1, 0.1)
    qpar = None
    if use_quantizer:
        gns = {"torch": torch, "nn": nn, "F": F}
        exec(source_of(f"{REF}/modules/gumbel_vector_quantizer.py", "GumbelVectorQuantizer"), gns)
        groups, num_vars, vq_dim = 2, 5, 12
        gq = gns["GumbelVectorQuantizer"](dim=C, num_vars=num_vars, temp=(2.0, 0.5, 0.999), groups=groups, combine_groups=False,
                                          vq_dim=vq_dim, time_first=True)
        gq.weight_proj.weight.data = O.hash_uniform(tag + ".qw", (groups * num_vars, C), -1.0, 1.0)
        gq.weight_proj.bias.data = O.hash_uniform(tag + ".qb", (groups * num_vars,), -0.1, 0.1)
        gq.vars.data = O.hash_uniform(tag + ".qv", (1, groups * num_vars, vq_dim // groups), 0.0, 1.0)
        gq.eval()
        pq = nn.Linear(vq_dim, Dp)
        pq.weight.data = O.hash_uniform(tag + ".pw", (Dp, vq_dim), -0.5, 0.5)
        pq.bias.data = O.hash_uniform(tag + ".pb", (Dp,), -0.1, 0.1)
        self.quantizer, self.project_q = gq, pq
        qpar = (groups, num_vars, vq_dim)
    fn_ns = dict(ns, self=self)
    exec(source_of(sat, "UniSpeechSATModel", "forward", "compute_pred_spk"), fn_ns)
    compute_pred_spk = fn_ns["compute_pred_spk"]
    # same number of masked frames per utterance (the reference's .view(B, -1, C) needs it)
    spk_x = O.hash_uniform(tag + ".x", (B, T, C), -1.0, 1.0)
    mask = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):
        mask[b, torch.tensor([1 + b, 3 + b, 4 + b, 7 + b, 9 + b])] = True
    pad = torch.zeros(B, T, dtype=torch.bool)
    b_pos = torch.arange(B).unsqueeze(1).expand(B, T)
    masked = ~pad & mask
    x_m = spk_x[masked].view(B, -1, C)
    torch.manual_seed(seed)
    loss, q, mean_t, acc = compute_pred_spk(x_m, spk_proj(x_m), b_pos[masked].view(B, -1))
    out = {"loss": float(loss.detach()), "mean_targets": float(mean_t), "acc": float(acc)}
    if q is not None:
        out.update(code_perplexity=float(q["code_perplexity"]), prob_perplexity=float(q["prob_perplexity"]), num_vars=int(q["num_vars"]))
    return dict(B=B, T=T, C=C, Dp=Dp, temp=temp, qpar=qpar, tag=tag, seed=seed, mask=mask.numpy()), out


def main():
    cases = [(False, 4, 0, 11), (False, 3, 5, 12), (True, 0, 6, 13), (True, 2, 2, 14)]
    store = {"cases": np.array([[int(q), n, c, s] for q, n, c, s in cases])}
    for i, (q, n, c, s) in enumerate(cases):
        meta, out = build_case(q, n, c, s)
        store[f"mask_{i}"] = meta["mask"]
        store[f"out_{i}"] = np.array([out["loss"], out["mean_targets"], out["acc"], out.get("code_perplexity", -1.0),
                                      out.get("prob_perplexity", -1.0), out.get("num_vars", -1)], dtype=np.float64)
        print(i, meta["tag"], out)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sat_heads.npz"), **store)


if __name__ == "__main__":
    main()
