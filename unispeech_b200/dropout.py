"""Host side of training-mode dropout: one 64-bit seed per forward pass, one (key0, key1) pair per dropout site.

The kernels (csrc/dropout.cuh) decide keep/drop for an element from (key0, key1, row, column) alone, so nothing but the
seed is carried from the forward to the backward pass.  Sites follow the reference: `dropout_input` after
post_extract_proj (WavLM/WavLM.py:350), the encoder-level `F.dropout` after pos_conv [+ LayerNorm] (:584), and per layer
dropout1 / dropout2 (activation) / dropout3 (:659-661,702-738) and the dropout on the attention probabilities
(WavLM/modules.py:551).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

_M64 = (1 << 64) - 1

SITE_INPUT = 0
SITE_ENCODER = 1
L_DROPOUT1, L_ACTIVATION, L_DROPOUT3, L_ATTENTION = 0, 1, 2, 3


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def layer_site(layer: int, which: int) -> int:
    return 4 + 4 * layer + which


def site_key(seed: int, site: int) -> Tuple[int, int]:
    z = splitmix64((seed ^ (site * 0xD1342543DE82EF95)) & _M64)
    return z & 0xFFFFFFFF, z >> 32


class DropState:
    """Dropout probabilities of one training forward pass and the seed its masks derive from."""

    __slots__ = ("seed", "p", "p_attn", "p_act", "p_input")

    def __init__(self, seed: int, p: float, p_attn: float, p_act: float, p_input: float):
        for name, v in (("dropout", p), ("attention_dropout", p_attn), ("activation_dropout", p_act), ("dropout_input", p_input)):
            if not 0.0 <= v < 1.0:
                raise ValueError(f"{name}={v} must be in [0, 1)")
        self.seed, self.p, self.p_attn, self.p_act, self.p_input = seed & _M64, p, p_attn, p_act, p_input

    def key(self, site: int) -> Tuple[int, int]:
        return site_key(self.seed, site)

    @staticmethod
    def for_model(cfg, training: bool, fixed_seed: Optional[int]) -> Optional["DropState"]:
        """None unless the module is in training mode and some probability is non-zero.  The seed comes from torch's CPU
        generator (reproducible under torch.manual_seed, no device sync) unless the caller pinned one."""
        if not training:
            return None
        ps = (cfg.dropout, cfg.attention_dropout, cfg.activation_dropout, cfg.dropout_input)
        if not any(v > 0 for v in ps):
            return None
        seed = fixed_seed if fixed_seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        return DropState(seed, *ps)
