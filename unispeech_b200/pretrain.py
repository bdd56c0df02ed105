"""Pre-training surface of the fairseq WavLM model: encoder + masked-prediction head + criterion on the B200 kernels.

Mirrors:
  * `WavLMModel.forward(source, target_list, padding_mask, mask, features_only, output_layer)`
    -- src/fairseq/models/wavlm/wavlm.py:465-576 (state_dict keys `final_proj.*`, `label_embs_concat` as there, :328-345)
  * `WavLMCriterion.get_loss`  -- src/fairseq/criterions/wavlm_criterion.py:52-138: sum-reduced cross entropy over the masked
    (x pred_masked_weight) and unmasked (x pred_nomask_weight) frames, `sample_size`, `features_pen` extra loss, accuracy counts.
The reference materialises `[C+1, S, final_dim]` expanded targets and `[S, C+1]` logits per label set; here the logits are one
tcgen05 GEMM against the row-normalised label embeddings and the softmax / cross entropy / backward operand come from one
row kernel (csrc/nce.cu), so the loss is a scalar produced on the device with no host synchronisation.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .engine import BF
from .wavlm import WavLM, WavLMConfig, _on_forward_stream


class WavLMPretrainConfig(WavLMConfig):
    """WavLMConfig + the pre-training fields of the fairseq dataclass (src/fairseq/models/wavlm/wavlm.py:50,104-135,223-230)."""

    def __init__(self, cfg=None):
        self.label_rate = 50
        self.sample_rate = 16000
        self.final_dim = 256
        self.untie_final_proj = False
        self.logit_temp = 0.1
        self.target_glu = False
        self.skip_masked = False
        self.skip_nomask = False
        super().__init__(cfg)


def _rows(n, C, dtype, dev):
    return torch.empty(n, C, dtype=dtype, device=dev)


class _MaskedPredictionFn(torch.autograd.Function):
    """loss = sum over label sets of weight * CE(cos(final_proj(x[idx]), label_embs) / temp, target).  x: bf16 [B*T, D]."""

    @staticmethod
    def forward(ctx, x2d, w, b, label_embs, model, idx, targets, weight, stats):
        ctx.fwd_stream = torch.cuda.current_stream()
        eng = model._engine
        dev = x2d.device
        S, D = idx.numel(), x2d.shape[1]
        Dp, n_sets = model.final_dim, len(model.num_classes)
        Dt = w.shape[0]
        untie = model.untie_final_proj
        wp, wpT = torch.empty(Dt, D, dtype=BF, device=dev), torch.empty(D, Dt, dtype=BF, device=dev)
        ops.prep_linear(w, Dt, D, 1.0, wp, D, wpT, Dt)
        xs = _rows(S, D, BF, dev)
        ops.gather_rows(x2d, D, idx, S, D, xs, D)
        proj = _rows(S, Dt, BF, dev)
        ops.gemm_rows(xs, 0, D, S, 1, D, wp, Dt, proj, 0, Dt, L.make_epilogue(bias=b))
        loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        sets, off = [], 0
        for i, C in enumerate(model.num_classes):
            Cpad = (C + 63) // 64 * 64
            E = label_embs[off:off + C]
            en, en_t = _rows(Cpad, Dp, BF, dev), _rows(Dp, Cpad, BF, dev)
            invn = torch.empty(C, dtype=torch.float32, device=dev)
            ops.nce_prep(E, C, Cpad, Dp, en, en_t, invn)
            proj_i = proj[:, i * Dp:(i + 1) * Dp] if untie else proj
            zraw = _rows(S, Cpad, BF, dev)
            ops.gemm_rows(proj_i, 0, Dt, S, 1, Dp, en, Cpad, zraw, 0, Cpad, None)
            G = _rows(S, Cpad, BF, dev)
            pn = torch.empty(S, dtype=torch.float32, device=dev)
            rvec = torch.empty(S, dtype=torch.float32, device=dev)
            part = torch.zeros(1, dtype=torch.float64, device=dev)
            correct = torch.zeros(1, dtype=torch.int32, device=dev)
            ops.nce_ce(proj_i, Dt, Dp, zraw, Cpad, targets[i], S, C, Cpad, model.logit_temp, weight, G, Cpad, pn, rvec, part, correct)
            loss_sum += part
            stats.append(dict(loss=part, correct=correct, count=S))
            sets.append((off, C, Cpad, en_t, invn, G, pn, rvec))
            off += C
        ctx.model, ctx.eng, ctx.idx, ctx.sets = model, eng, idx, sets
        ctx.xs, ctx.proj, ctx.wpT, ctx.shape = xs, proj, wpT, (x2d.shape[0], D, S, Dp, Dt, untie)
        ctx.save_for_backward(w, b, label_embs)
        return loss_sum.float().reshape(())

    @staticmethod
    @_on_forward_stream
    def backward(ctx, dloss):
        model, eng, idx = ctx.model, ctx.eng, ctx.idx
        w, b, label_embs = ctx.saved_tensors
        rows, D, S, Dp, Dt, untie = ctx.shape
        dev = ctx.xs.device
        g = eng.g
        dproj = _rows(S, Dt, BF, dev)
        scale_bf, scale_f = dloss.to(BF), dloss.float()
        for i, (off, C, Cpad, en_t, invn, G, pn, rvec) in enumerate(ctx.sets):
            G.mul_(scale_bf)     # upstream gradient of the scalar loss (device scalar, no sync): everything below is linear in G
            rvec.mul_(scale_f)
            proj_i = ctx.proj[:, i * Dp:(i + 1) * Dp] if untie else ctx.proj
            first = untie or i == 0
            tgt = (dproj[:, i * Dp:(i + 1) * Dp] if untie else dproj) if first else _rows(S, Dp, BF, dev)
            ops.gemm_rows(G, 0, Cpad, S, 1, Cpad, en_t, Dp, tgt, 0, tgt.stride(0), None)          # G En
            ops.nce_dproj(tgt, tgt.stride(0), proj_i, Dt, S, Dp, pn, rvec)
            if not first:
                dproj.add_(tgt)  # tied final_proj shared by several label sets
            d_en = torch.zeros(Cpad, Dp, dtype=torch.float32, device=dev)
            ops.gemm_wgrad(G, 0, Cpad, proj_i, 0, Dt, S, 1, Cpad, Dp, d_en, Dp)                    # G^T proj
            ops.nce_dlabel(d_en, label_embs[off:off + C], invn, C, Dp, g(model.label_embs_concat)[off:off + C])
        ops.colsum(dproj, 0, Dt, S, 1, Dt, g(model.final_proj.bias))
        ops.gemm_wgrad(dproj, 0, Dt, ctx.xs, 0, D, S, 1, Dt, D, g(model.final_proj.weight), D)
        dxs = _rows(S, D, BF, dev)
        ops.gemm_rows(dproj, 0, Dt, S, 1, Dt, ctx.wpT, D, dxs, 0, D, None)
        dx = torch.zeros(rows, D, dtype=BF, device=dev)
        ops.scatter_add_rows(dxs, D, idx, S, D, dx, D)
        ctx.sets = ctx.xs = ctx.proj = None
        return dx, None, None, None, None, None, None, None, None


class WavLMForPretraining(WavLM):
    """`WavLM` + `final_proj` / `label_embs_concat` (same state_dict keys as the fairseq model) and the fused criterion."""

    def __init__(self, cfg: WavLMPretrainConfig, num_classes: List[int]):
        super().__init__(cfg)
        self._want_features_pen = True
        if cfg.target_glu:
            raise NotImplementedError("target_glu is not implemented in the fused masked-prediction head")
        D = cfg.encoder_embed_dim
        self.final_dim = cfg.final_dim if cfg.final_dim > 0 else D
        assert self.final_dim % 64 == 0, "final_dim must be a multiple of 64"
        self.num_classes = [int(c) for c in num_classes]
        assert all(0 < c <= 1024 for c in self.num_classes), "label sets of 1..1024 classes are supported"
        self.untie_final_proj = bool(cfg.untie_final_proj)
        self.logit_temp = float(cfg.logit_temp)
        self.skip_masked, self.skip_nomask = bool(cfg.skip_masked), bool(cfg.skip_nomask)
        self.feat2tar_ratio = cfg.label_rate * 320 / cfg.sample_rate  # label_rate * feature_ds_rate / sample_rate, wavlm.py:277-279
        self.final_proj = nn.Linear(D, self.final_dim * (len(self.num_classes) if self.untie_final_proj else 1))
        self.label_embs_concat = nn.Parameter(torch.empty(sum(self.num_classes), self.final_dim))
        nn.init.uniform_(self.label_embs_concat)  # wavlm.py:345

    # ---- reference helpers
    def forward_targets(self, T: int, target_list: List[torch.Tensor]) -> List[torch.Tensor]:
        """Label sub-sampling of wavlm.py:440-451 (features are never trimmed here: the labels must cover every frame)."""
        targ_tsz = min(t.size(1) for t in target_list)
        if self.feat2tar_ratio * T > targ_tsz:
            raise NotImplementedError(f"labels ({targ_tsz} per utterance) are shorter than the {T} feature frames: trimming the "
                                      "features is not implemented")
        inds = (torch.arange(T).float() * self.feat2tar_ratio).long()
        return [t[:, inds.to(t.device)] for t in target_list]

    def remove_pretraining_modules(self):
        self.final_proj = None
        self.label_embs_concat = None

    # ---- BaseFairseqModel surface the trainer / criterion touch (src/fairseq/models/fairseq_model.py:102-158, wavlm.py:599-627)
    def set_num_updates(self, num_updates: int):
        self.num_updates = num_updates

    def upgrade_state_dict_named(self, state_dict, name):
        return state_dict

    def get_extra_losses(self, net_output):
        """(losses, names) exactly as wavlm.py:611-620: the feature penalty when the forward pass produced it."""
        extra_losses, names = [], []
        if net_output.get("features_pen") is not None:
            extra_losses.append(net_output["features_pen"])
            names.append("features_pen")
        return extra_losses, names

    def get_logits(self, net_output, is_masked=True):
        raise NotImplementedError("the [S, C+1] logit lists of the reference are never materialised on this path: use "
                                  "`model.criterion(net_output, ...)`, which returns the same loss / sample_size / accuracy counts as "
                                  "WavLMCriterion.get_loss")

    get_targets = get_logits

    def forward(self, source, target_list=None, padding_mask=None, mask=True, features_only=False, output_layer=None,
                mask_indices=None):
        """fairseq WavLMModel.forward.  With `features_only=False` the result carries everything the criterion needs
        (`x`, `padding_mask`, `mask_indices`, the frame-aligned `target_list`, `features_pen`); logits are never materialised."""
        self.extract_features(source, padding_mask=padding_mask, mask=mask, output_layer=output_layer, mask_indices=mask_indices)
        res = self._last
        out = {"x": res["x"], "padding_mask": res["padding_mask"], "features": res["features"],
               "layer_results": res["layer_results"]}
        if features_only:
            return out
        T = res["x"].shape[1]
        out["mask_indices"] = res["mask_indices"]
        out["padding_mask_host"] = res.get("padding_mask_host")  # host copy of the frame mask when the caller's mask was on the host
        out["target_list"] = self.forward_targets(T, target_list) if target_list is not None else None
        out["features_pen"] = self._last_pen  # mean(features^2) after GradMultiply, wavlm.py:477-484 (kernel, inside _ConvFn)
        return out

    def criterion(self, net_output: Dict, pred_masked_weight: float = 1.0, pred_nomask_weight: float = 0.0,
                  loss_weights: Optional[List[float]] = None):
        """WavLMCriterion.get_loss (wavlm_criterion.py:52-138): returns (loss, sample_size, logging_output) with `loss` a
        device scalar; logging values stay device tensors (call `.item()` when you log)."""
        x = net_output["x"]
        B, T, D = x.shape
        dev = x.device
        mi, pm, targets = net_output["mask_indices"], net_output["padding_mask"], net_output["target_list"]
        assert mi is not None and targets is not None, "forward(..., target_list=..., mask=True) must run first"
        mi_h = mi.cpu() if mi.device.type != "cpu" else mi
        # frame selection happens on the host, like the reference's collater-side masks: with a host padding mask there is no
        # device round trip at all (a device-only mask costs one synchronising copy here)
        pm_h = net_output.get("padding_mask_host")
        if pm_h is None:
            pm_h = torch.zeros(B, T, dtype=torch.bool) if pm is None else (pm.cpu() if pm.device.type != "cpu" else pm)
        x2d = x.reshape(B * T, D)
        if x2d.dtype != BF or not x2d.is_contiguous():
            x2d = x2d.to(BF).contiguous()
        loss, sample_size, log = 0.0, 0, {}
        plans = []
        if not self.skip_masked and pred_masked_weight > 0:
            plans.append(("m", torch.logical_and(~pm_h, mi_h), pred_masked_weight))
        if not self.skip_nomask and pred_nomask_weight > 0:
            plans.append(("u", torch.logical_and(~pm_h, ~mi_h), pred_nomask_weight))
        for tag, sel, wgt in plans:
            idx_h = torch.nonzero(sel.reshape(-1), as_tuple=False).squeeze(1)
            if idx_h.numel() == 0:
                continue
            idx = idx_h.to(torch.int32).to(dev, non_blocking=True)
            tg = [t.reshape(-1).to(dev)[idx.long()].to(torch.int32).contiguous() if t.device.type != "cpu"
                  else t.reshape(-1)[idx_h].to(torch.int32).to(dev, non_blocking=True) for t in targets]
            stats = []
            part = _MaskedPredictionFn.apply(x2d, self.final_proj.weight, self.final_proj.bias, self.label_embs_concat, self, idx,
                                             tg, float(wgt), stats)
            loss = loss + part
            sample_size += idx_h.numel()
            for i, st in enumerate(stats):
                log[f"loss_{tag}_{i}"] = st["loss"] / wgt
                log[f"correct_{tag}_{i}"] = st["correct"]
                log[f"count_{tag}_{i}"] = st["count"]
        if loss_weights is not None and net_output.get("features_pen") is not None:
            coef = loss_weights[0]
            if coef != 0:
                p = coef * net_output["features_pen"].float() * sample_size  # wavlm_criterion.py:100-103
                loss = loss + p
                log["loss_features_pen"] = p.detach()
        log.update(ntokens=sample_size, sample_size=sample_size, nsentences=B)
        log["loss"] = loss.detach() if torch.is_tensor(loss) else loss
        return loss, sample_size, log
