"""Generate tests/golden/mask_indices.npz from the UNMODIFIED reference sampler (/root/reference/WavLM/WavLM.py:35-159).
Authoring container only.  Each case seeds numpy, calls the reference compute_mask_indices and stores the bool mask."""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/WavLM")
import WavLM as ref  # noqa: E402

# (B, T, mask_prob, mask_length, pad_from (-1: no padding mask), seed, no_overlap)
CASES = [
    (4, 749, 0.65, 10, -1, 0, 0),
    (4, 749, 0.8, 10, 500, 1, 0),
    (2, 199, 0.65, 10, 120, 2, 0),
    (8, 999, 0.8, 10, -1, 3, 0),
    (3, 49, 0.65, 10, 30, 4, 0),
    (16, 1499, 0.65, 10, 1000, 5, 0),
]
out = {"cases": np.array(CASES, dtype=np.float64)}
for i, (B, T, prob, length, pad_from, seed, no_overlap) in enumerate(CASES):
    pm = None
    if pad_from >= 0:
        pm = torch.zeros(B, T, dtype=torch.bool)
        pm[-1, pad_from:] = True
    np.random.seed(seed)
    out[f"mask_{i}"] = ref.compute_mask_indices((B, T), pm, prob, length, "static", 0, min_masks=2,
                                                no_overlap=bool(no_overlap), min_space=1)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mask_indices.npz"), **out)
print("wrote", len(CASES), "cases")
