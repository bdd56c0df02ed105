"""Recipe that makes the UNMODIFIED reference available as `oracle/_ref/` (test / baseline infrastructure, never the product).

The reference path is two pure-Python files, `/root/reference/WavLM/{WavLM,modules}.py` (no C/C++ on this path, nothing to
compile).  `build()` copies them byte for byte from where they lie into `oracle/_ref/` (git-ignored, NOT gpurun-ignored: it
travels to the GPU box like a built .so) and records their SHA-256.  `/root/reference` only exists in the authoring container;
on the GPU box the prebuilt copy is used as is.  `load()` imports that copy and applies the ONE harness patch the rest of the
test infrastructure uses (tools/make_golden.py): `TransformerEncoder.extract_features` is replaced by an out-of-place
restatement of WavLM/WavLM.py:572-612, because the reference's in-place `x += x_conv` / `x[padding_mask] = 0` break autograd on
torch >= 2; its forward is asserted bit-identical to the unpatched one before any fixture is generated (tools/make_golden.py).
Used by: `bench.py --impl reference` / `cpu_baseline` (kind "reference"), tools/make_*golden*.py.  Nothing under
`unispeech_b200/` imports this.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/WavLM"
REF_DST = os.path.join(HERE, "_ref")
FILES = ("WavLM.py", "modules.py")


def build() -> bool:
    """Copy the reference sources into oracle/_ref (when /root/reference is present).  Returns True if oracle/_ref is usable."""
    if os.path.isdir(REF_SRC):
        os.makedirs(REF_DST, exist_ok=True)
        meta = {}
        for f in FILES:
            shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DST, f))
            meta[f] = hashlib.sha256(open(os.path.join(REF_DST, f), "rb").read()).hexdigest()
        json.dump({"source": REF_SRC, "sha256": meta}, open(os.path.join(REF_DST, "MANIFEST.json"), "w"), indent=1)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES)


def _patched_extract_features(self, x, padding_mask=None, streaming_mask=None, tgt_layer=None):
    """Out-of-place restatement of TransformerEncoder.extract_features (WavLM/WavLM.py:572-612); forward bit-identical."""
    import numpy as np
    import torch.nn.functional as F
    if padding_mask is not None:
        x = x.masked_fill(padding_mask.unsqueeze(-1), 0.0)
    x_conv = self.pos_conv(x.transpose(1, 2)).transpose(1, 2)
    x = x + x_conv
    if not self.layer_norm_first:
        x = self.layer_norm(x)
    x = F.dropout(x, p=self.dropout, training=self.training)
    x = x.transpose(0, 1)
    layer_results = []
    z = None
    if tgt_layer is not None:
        layer_results.append((x, z))
    r = None
    pos_bias = None
    for i, layer in enumerate(self.layers):
        dropout_probability = np.random.random()
        if not self.training or (dropout_probability > self.layerdrop):
            x, z, pos_bias = layer(x, self_attn_padding_mask=padding_mask, need_weights=False,
                                   self_attn_mask=streaming_mask, pos_bias=pos_bias)
        if tgt_layer is not None:
            layer_results.append((x, z))
        if i == tgt_layer:
            r = x
            break
    if r is not None:
        x = r
    x = x.transpose(0, 1)
    return x, layer_results


_MOD = None


def load(patch_autograd: bool = True):
    """Import oracle/_ref/WavLM.py (it does `from modules import ...`, WavLM.py:20) and return the module."""
    global _MOD
    if _MOD is None:
        if not available():
            raise RuntimeError("oracle/_ref is missing: run `python -m oracle.build_ref` where /root/reference exists")
        warnings.filterwarnings("ignore")
        if REF_DST not in sys.path:
            sys.path.insert(0, REF_DST)
        spec = importlib.util.spec_from_file_location("_unispeech_ref_wavlm", os.path.join(REF_DST, "WavLM.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod._orig_extract_features = mod.TransformerEncoder.extract_features
        _MOD = mod
    _MOD.TransformerEncoder.extract_features = _patched_extract_features if patch_autograd else _MOD._orig_extract_features
    return _MOD


def build_model(cfg, state_dict, train: bool = False):
    """Reference `WavLM(WavLMConfig(cfg))` with the given state_dict loaded (strict)."""
    ref = load()
    m = ref.WavLM(ref.WavLMConfig(dict(vars(cfg))))
    m.load_state_dict(state_dict, strict=True)
    return m.train() if train else m.eval()


if __name__ == "__main__":
    print("oracle/_ref available:", build())
