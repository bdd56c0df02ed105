"""Pin the oracle's restatements of the pre-training head and the optimizer step to the REFERENCE CODE ITSELF.

The fairseq package of /root/reference cannot be imported here (no omegaconf / hydra; Python-3.12-incompatible dataclasses,
SURVEY.md section 8c), but the three pieces the oracle restates are plain torch code.  This script (run in the authoring
container only) extracts their source text from the reference tree with `ast`, executes it unmodified, feeds it hash-generated
inputs and commits inputs' seeds + outputs as tests/golden/train_heads.npz:
  * WavLMModel.compute_nce            src/fairseq/models/wavlm/wavlm.py:426-438
  * utils.clip_grad_norm_             src/fairseq/utils.py:338-381
  * Adam (torch.optim.Optimizer)      src/fairseq/optim/adam.py:100-228
Nothing from the reference is copied into the repository: only numbers.
"""
import ast
import math
import os
import sys
import textwrap
import types
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wavlm_oracle as O  # noqa: E402

REF = "/root/reference/src/fairseq"


def grab(path, cls, name):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if cls is None and isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            return ast.get_source_segment(src, node)
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    return textwrap.dedent(ast.get_source_segment(src, sub, padded=True))
    raise KeyError((path, cls, name))


def main():
    out = {}
    # ---- compute_nce on a small masked-prediction problem
    ns = {"torch": torch}
    exec(grab(f"{REF}/models/wavlm/wavlm.py", "WavLMModel", "compute_nce"), ns)
    S, C, Dp, temp = 23, 17, 32, 0.1
    proj = O.hash_uniform("g.proj", (S, Dp), -1.0, 1.0)
    E = O.hash_uniform("g.emb", (C, Dp), 0.0, 1.0)
    tgt = (O.hash_uniform("g.tgt", (S,), 0.0, 1.0) * C).long().clamp(max=C - 1)
    y = torch.index_select(E, 0, tgt)
    negs = E.unsqueeze(1).expand(-1, S, -1)
    logits = ns["compute_nce"](types.SimpleNamespace(logit_temp=temp), proj, y, negs)
    loss = F.cross_entropy(logits.float(), logits.new_zeros(S, dtype=torch.long), reduction="sum")  # wavlm_criterion.py:70
    out.update(nce_logits=logits.numpy(), nce_loss=np.float64(loss.item()), nce_shape=np.array([S, C, Dp]), nce_temp=np.float64(temp))

    # ---- clip_grad_norm_ + Adam for three steps
    ns2 = {"torch": torch, "warnings": warnings, "multi_tensor_l2norm_available": False, "multi_tensor_total_norm": None}
    exec(grab(f"{REF}/utils.py", None, "clip_grad_norm_"), ns2)
    ns3 = {"torch": torch, "math": math}
    exec(grab(f"{REF}/optim/adam.py", None, "Adam"), ns3)
    shapes = [(5, 7), (3,), (2, 3, 4)]
    params = [torch.nn.Parameter(O.hash_uniform(f"g.p{i}", s, -1.0, 1.0)) for i, s in enumerate(shapes)]
    opt = ns3["Adam"](params, lr=3e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    norms = []
    for step in range(3):
        for i, p in enumerate(params):
            p.grad = O.hash_uniform(f"g.g{step}.{i}", tuple(p.shape), -1.0, 1.0) * (0.5 + step)
        norms.append(float(ns2["clip_grad_norm_"](params, 1.5)))
        opt.step()
    out["adam_norms"] = np.array(norms)
    for i, p in enumerate(params):
        out[f"adam_p{i}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train_heads.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
