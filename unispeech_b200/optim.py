"""Fused optimizer step for the fp32 masters of a `unispeech_b200.WavLM`.

`FusedAdam` mirrors the surface the reference trainer drives (src/fairseq/optim/fairseq_optimizer.py and
fp16_optimizer.py: `multiply_grads`, `clip_grad_norm`, `step`, `zero_grad`, `set_lr/get_lr`; update rule of
src/fairseq/optim/adam.py:150-228) but runs on the model's flat gradient buffer: one launch for the global gradient norm,
one for scale + clip + Adam + (optionally) zeroing the gradients, no host synchronisation and no per-tensor kernels.
The data-parallel allreduce (`parallel.all_reduce_grads`) runs on the same flat buffer right before it.
"""
from __future__ import annotations

import struct
from typing import Tuple

import torch

from . import ops

_CHUNK = 2048


class FusedAdam:
    def __init__(self, model, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, exclude=()):
        """`exclude`: parameters to leave alone.  fairseq's Adam skips a parameter whose `.grad` is None (adam.py:160-161); here
        every parameter has a view of the flat gradient buffer, so one the backward pass never reaches would see a ZERO gradient
        and still be weight-decayed.  The frozen feature extractor (`feature_grad_mult <= 0`: the conv stack runs under no_grad,
        WavLM/WavLM.py:333-339) is excluded automatically; name anything else that is frozen."""
        eng = model._engine
        if eng is None or eng.flat is None:
            raise RuntimeError("run one forward pass on the GPU (or call model._engine_for(device)) before building the optimizer: "
                               "it updates the parameters through the engine's flat gradient buffer")
        self.model, self.eng = model, eng
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        flat = eng.flat
        dev = flat.flat.device
        self.g = flat.flat
        self.exp_avg = torch.zeros_like(self.g)      # state["exp_avg"], adam.py:183
        self.exp_avg_sq = torch.zeros_like(self.g)   # state["exp_avg_sq"], adam.py:185
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._step = 0
        self._multiply_factor = 1.0
        self._max_norm = 0.0
        self._have_norm = False
        skip = {id(p) for p in exclude}
        if getattr(model, "feature_grad_mult", 1.0) <= 0:
            skip |= {id(p) for p in model.feature_extractor.parameters()}
        recs, chunks, self._ptrs = [], 0, []
        for p in flat.params:
            if not p.requires_grad or id(p) in skip:
                continue
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam needs contiguous fp32 master parameters")
            recs.append(struct.pack("<Qqqq", p.data_ptr(), flat.offsets[id(p)], p.numel(), chunks))
            chunks += (p.numel() + _CHUNK - 1) // _CHUNK
            self._ptrs.append((p, p.data_ptr()))
        self._n, self._chunks = len(recs), chunks
        self._table = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev)

    # ---- fairseq optimizer surface
    def get_lr(self) -> float:
        return self.lr

    def set_lr(self, lr: float):
        self.lr = float(lr)

    def multiply_grads(self, c: float):
        """Deferred like FP16Optimizer.multiply_grads (fp16_optimizer.py:190-192): folded into the update kernel."""
        self._multiply_factor *= float(c)

    def clip_grad_norm(self, max_norm: float) -> torch.Tensor:
        """Global gradient norm (device tensor, no sync); gradients are clipped inside `step()` (fp16_optimizer.py:194-214)."""
        self._sumsq.zero_()
        # only the tensors this optimizer owns: gradients the backward kernels leave in the flat buffer for excluded / frozen
        # parameters do not count (fairseq's clip_grad_norm_ sees the optimizer's parameters, src/fairseq/utils.py:338-345)
        ops.sumsq_table(self._table, self._n, self._chunks, self.g, self._sumsq)
        self._max_norm = float(max_norm)
        self._have_norm = True
        return self._sumsq.sqrt().float() * abs(self._multiply_factor)

    def step(self, zero_grad: bool = False):
        for p, ptr in self._ptrs:
            if p.data_ptr() != ptr:
                raise RuntimeError("a parameter was re-allocated after the optimizer was built")
        self._step += 1
        use_clip = self._have_norm and self._max_norm > 0
        ops.adam_step(self._table, self._n, self._chunks, self.g, self.exp_avg, self.exp_avg_sq,
                      self._sumsq if use_clip else None, self._multiply_factor, self._max_norm if use_clip else 0.0, self.lr,
                      self.betas[0], self.betas[1], self.eps, self.weight_decay, self._step, zero_grad)
        self._multiply_factor, self._have_norm = 1.0, False
        self.eng.prepared_version = None  # the masters changed behind autograd's version counters: re-derive the bf16 operands

    def zero_grad(self):
        ops.memset_zero(self.g)
        self._multiply_factor, self._have_norm = 1.0, False

    def state_dict(self):
        return {"step": self._step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr, "betas": self.betas,
                "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self._step = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
