"""Data-parallel gradient exchange: ONE collective per step on the flat fp32 gradient buffer.

Mirrors `LegacyDistributedDataParallel.all_reduce_grads` (src/fairseq/legacy_distributed_data_parallel.py:76-165, called from
src/fairseq/trainer.py:781-785): gradients are divided by the world size and summed across ranks.  The reference packs
every gradient into a temporary flat buffer and unpacks it afterwards (two extra full-gradient copies); here the kernels
already accumulate into the flat buffer, so the exchange is a single in-place `all_reduce` (NCCL over NVLink/NVSwitch on
the GPU box, gloo in the CPU tests).  One process per GPU, launched with torch.distributed.run.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def all_reduce_grads(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place average of the flat gradient buffer over the process group (no-op for a single process)."""
    if not dist.is_available() or not dist.is_initialized():
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    flat.div_(world)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous utterance shard of a global batch for this rank (utterances are independent: SURVEY.md section 8e)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
