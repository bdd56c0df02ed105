"""Generate tests/golden/utterance_mixing.npz from the reference's own `mixing_collated_audios`
(src/fairseq/data/audio/utterance_mixing_dataset.py:373-438), extracted with `ast` and executed unmodified (the fairseq package
cannot be imported here).  Authoring container only; only numbers are committed."""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wavlm_oracle as O  # noqa: E402

path = "/root/reference/src/fairseq/data/audio/utterance_mixing_dataset.py"
src = open(path).read()
fn = None
for node in ast.walk(ast.parse(src)):
    if isinstance(node, ast.FunctionDef) and node.name == "mixing_collated_audios":
        fn = textwrap.dedent(ast.get_source_segment(src, node, padded=True))
ns = {"np": np, "torch": torch, "F": F}
exec(fn, ns)
# (B, T, mixing_prob, mixing_num, mixing_max_len, normalize, seed)
CASES = [(4, 4000, 0.5, 1, -1, 0, 0), (6, 2500, 1.0, 2, 4, 1, 1), (3, 999, 0.7, 1, -1, 1, 2), (5, 1600, 0.2, 3, 2, 0, 3)]
out = {"cases": np.array(CASES, dtype=np.float64)}
for i, (B, T, prob, num, mlen, norm, seed) in enumerate(CASES):
    self = types.SimpleNamespace(mixing_max_len=mlen, mixing_prob=prob, mixing_noise=False, mixing_noise_prob=0.0, mixing_num=num,
                                 normalize=bool(norm))
    wav = O.hash_uniform(f"mix{i}", (B, T), -1.0, 1.0)
    wav[-1, T - T // 3:] = 0.0   # a zero-padded tail, as collated batches have
    np.random.seed(seed)
    out[f"mixed_{i}"] = ns["mixing_collated_audios"](self, wav).numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "utterance_mixing.npz"), **out)
print("wrote", len(CASES), "cases")
