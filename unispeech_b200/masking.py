"""Host-side span-mask sampler with the reference's semantics AND its numpy RNG consumption order, so that seeding
`np.random` identically reproduces the reference's masks (tests/test_masking.py checks this against fixtures generated
from /root/reference/WavLM/WavLM.py:35-159 `compute_mask_indices`).  This is input-pipeline glue that the reference also
runs on the host; it is not part of the GPU hot path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def _span_lengths(kind: str, count: int, mask_length: int, mask_other: float):
    if kind == "static":
        return np.full(count, mask_length)
    if kind == "uniform":
        return np.random.randint(mask_other, mask_length * 2 + 1, size=count)
    if kind == "normal":
        return [max(1, int(round(v))) for v in np.random.normal(mask_length, mask_other, size=count)]
    if kind == "poisson":
        return [int(round(v)) for v in np.random.poisson(mask_length, size=count)]
    raise ValueError("unknown mask selection " + kind)


def _place_without_overlap(lengths, sz: int, min_space: int) -> np.ndarray:
    """Longest spans first; each span lands uniformly inside a free segment chosen with probability proportional to its
    size, and splits it (keeping `min_space` unmasked elements around the span)."""
    chosen = []
    free = [(0, sz)]
    shortest = min(lengths)
    for length in sorted(lengths, reverse=True):
        room = np.fromiter((hi - lo if hi - lo >= length + min_space else 0 for lo, hi in free), np.int64)
        total = np.sum(room)
        if total == 0:
            break
        pick = np.random.choice(len(free), p=room / total)
        lo, hi = free.pop(pick)
        start = np.random.randint(lo, hi - length)
        chosen.extend(range(start, start + length))
        if start - lo - min_space >= shortest:
            free.append((lo, start - min_space + 1))
        if hi - start - shortest - min_space > shortest:
            free.append((start + length + min_space, hi))
    return np.asarray(chosen)


def compute_mask_indices(shape: Tuple[int, int], padding_mask: Optional[torch.Tensor], mask_prob: float, mask_length: int,
                         mask_type: str = "static", mask_other: float = 0.0, min_masks: int = 0, no_overlap: bool = False,
                         min_space: int = 0) -> np.ndarray:
    """bool [B, T] array of masked frames: about `mask_prob * T / mask_length` spans of `mask_length` frames per row
    (probabilistic rounding), never starting inside the padded tail, every row trimmed to the same number of masked
    frames."""
    bsz, all_sz = shape
    out = np.full((bsz, all_sz), False)
    shared_count = max(min_masks, int(mask_prob * all_sz / float(mask_length) + np.random.rand()))
    rows = []
    for b in range(bsz):
        if padding_mask is not None:
            sz = all_sz - int(padding_mask[b].long().sum().item())
            count = max(min_masks, int(mask_prob * sz / float(mask_length) + np.random.rand()))
        else:
            sz, count = all_sz, shared_count
        lengths = _span_lengths(mask_type, count, mask_length, mask_other)
        if sum(lengths) == 0:
            lengths[0] = min(mask_length, sz - 1)
        if no_overlap:
            idc = _place_without_overlap(lengths, sz, min_space)
        else:
            shortest = min(lengths)
            if sz - shortest <= count:
                shortest = sz - count - 1
            starts = np.random.choice(sz - shortest, count, replace=False)
            idc = np.asarray([starts[j] + o for j in range(len(starts)) for o in range(lengths[j])])
        rows.append(np.unique(idc[idc < sz]))
    keep = min(len(r) for r in rows)
    for b, idc in enumerate(rows):
        if len(idc) > keep:
            idc = np.random.choice(idc, keep, replace=False)
        out[b, idc] = True
    return out
