"""`FairseqEncoder`-style wrappers the reference's fine-tuning models put around the pre-trained encoder:

  * `HubertEncoder.forward(source, padding_mask, tbc=True)`   src/fairseq/models/hubert/hubert_asr.py:314-340
        -> {"encoder_out": T x B x C, "encoder_padding_mask": B x T, "padding_mask": B x T}
  * `Wav2VecEncoder.forward(source, padding_mask, tbc=True)`  src/fairseq/models/wav2vec/wav2vec2_asr.py:390-421
        -> {"encoder_out": T x B x C, "encoder_padding_mask": T x B, "padding_mask": B x T, "layer_results": [...]}
    (it indexes the DICT form of extract_features: `res["x"]`, which is what `WavLM.forward(features_only=True)` returns)
  * `reorder_encoder_out`, `max_positions`, `set_num_updates`, `upgrade_state_dict_named`   fairseq_encoder.py:26-92

Both keep the reference's freeze logic (`freeze_finetune_updates`: the encoder runs under no_grad until that many updates) and
`apply_mask` (span masking only in training).  They are host glue around `unispeech_b200.wavlm.WavLM`: every tensor they return is a
view of the kernels' output.  The output projection (`proj`: CTC vocabulary / decoder width) and `final_dropout > 0` are the "next"
row after the encoder and raise instead of falling back to PyTorch kernels.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
import torch.nn as nn

from .wavlm import WavLM


class _EncoderBase(nn.Module):
    def __init__(self, w2v_model: WavLM, apply_mask: bool = False, freeze_finetune_updates: int = 0, final_dropout: float = 0.0,
                 output_dim: Optional[int] = None):
        super().__init__()
        d = w2v_model.cfg.encoder_embed_dim
        if output_dim is not None and output_dim != d:
            raise NotImplementedError("the output projection (CTC vocabulary / decoder width) is not built on the B200 kernels yet; "
                                      "project `encoder_out` in the caller")
        if final_dropout > 0:
            raise NotImplementedError("final_dropout > 0 is not wired to the dropout kernel in this wrapper yet")
        if hasattr(w2v_model, "remove_pretraining_modules"):
            w2v_model.remove_pretraining_modules()  # hubert_asr.py:290 / wav2vec2_asr.py:355
        self.w2v_model = w2v_model
        self.apply_mask = apply_mask
        self.freeze_finetune_updates = freeze_finetune_updates
        self.num_updates = 0
        self.proj = None

    def set_num_updates(self, num_updates: int):
        self.num_updates = num_updates

    def _context(self):
        ft = self.freeze_finetune_updates <= self.num_updates
        return contextlib.ExitStack() if ft else torch.no_grad()

    def reorder_encoder_out(self, encoder_out, new_order):
        if encoder_out["encoder_out"] is not None:
            encoder_out["encoder_out"] = encoder_out["encoder_out"].index_select(1, new_order)
        if encoder_out["encoder_padding_mask"] is not None:
            encoder_out["encoder_padding_mask"] = encoder_out["encoder_padding_mask"].index_select(0, new_order)
        return encoder_out

    def max_positions(self):
        return None

    def upgrade_state_dict_named(self, state_dict, name):
        return state_dict


class HubertEncoder(_EncoderBase):
    """hubert_asr.py:314-340: tuple form of `extract_features`; `encoder_padding_mask` is B x T."""

    def forward(self, source, padding_mask, tbc: bool = True, **kwargs):
        with self._context():
            x, padding_mask = self.w2v_model.extract_features(source=source, padding_mask=padding_mask,
                                                              mask=self.apply_mask and self.training)
            if tbc:
                x = x.transpose(0, 1)  # B x T x C -> T x B x C
        return {"encoder_out": x, "encoder_padding_mask": padding_mask, "padding_mask": padding_mask}


class Wav2VecEncoder(_EncoderBase):
    """wav2vec2_asr.py:390-421: dict form (`res["x"]`), `encoder_padding_mask` is T x B, `layer_results` passed through."""

    def forward(self, source, padding_mask, tbc: bool = True, **kwargs):
        with self._context():
            res = self.w2v_model(source=source, padding_mask=padding_mask, mask=self.apply_mask and self.training,
                                 features_only=True)
            x, padding_mask = res["x"], res["padding_mask"]
            if tbc:
                x = x.transpose(0, 1)
        return {"encoder_out": x,
                "encoder_padding_mask": padding_mask.transpose(0, 1) if padding_mask is not None else None,
                "padding_mask": padding_mask, "layer_results": res["layer_results"]}
