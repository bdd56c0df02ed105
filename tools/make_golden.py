"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (/root/reference/WavLM) on CPU.

Run in the authoring container only (the GPU box has no /root/reference):
    python tools/make_golden.py
Parameters and inputs are the hash-generated ones from oracle/wavlm_oracle.py, so fixtures only hold outputs.
The one harness-level patch: TransformerEncoder.extract_features is replaced by an out-of-place restatement of
WavLM/WavLM.py:572-612 because the reference's in-place `x += x_conv` breaks autograd on torch >= 2
(forward verified bit-identical below before any fixture is written).
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/WavLM")
warnings.filterwarnings("ignore")

import WavLM as ref  # noqa: E402  (reference module)
from oracle import wavlm_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def patched_extract_features(self, x, padding_mask=None, streaming_mask=None, tgt_layer=None):
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


def build_ref(cfg):
    rc = ref.WavLMConfig(dict(vars(cfg)))
    m = ref.WavLM(rc)
    sd = O.deterministic_state_dict(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    m.eval()
    return m, sd


def run_case(name, cfg, B, L, lengths=None, with_mask=False, grads=True):
    torch.manual_seed(0)
    m, sd = build_ref(cfg)
    wav, pmask = O.deterministic_waveform(B, L, seed=1, lengths=lengths)
    pm = pmask if lengths is not None else None
    T = O.num_frames(L, cfg)
    mask_idx = None
    if with_mask:
        mask_idx = (O.hash_uniform("maskidx", (B, T)) > 0.3)

    # 1) unpatched reference forward (no mask; the reference's apply_mask uses host numpy RNG)
    with torch.no_grad():
        (x0, lr0), pm0 = m.extract_features(wav.clone(), padding_mask=pm, mask=False, ret_layer_results=True,
                                           output_layer=cfg.encoder_layers)
        xfinal0, _ = m.extract_features(wav.clone(), padding_mask=pm, mask=False)
    # 2) patched forward must be bit-identical
    orig = ref.TransformerEncoder.extract_features
    ref.TransformerEncoder.extract_features = patched_extract_features
    try:
        with torch.no_grad():
            (x1, lr1), _ = m.extract_features(wav.clone(), padding_mask=pm, mask=False, ret_layer_results=True,
                                              output_layer=cfg.encoder_layers)
        assert torch.equal(x0, x1), "out-of-place patch changed the forward"
        # `features` (ret_conv=True) is taken under the out-of-place patch: the unpatched reference returns a tensor
        # that its own in-place `x += x_conv` / `x[padding_mask] = 0` has already overwritten (WavLM.py:574-579).
        with torch.no_grad():
            feats0, _ = m.extract_features(wav.clone(), padding_mask=pm, mask=False, ret_conv=True)
        out = {
            "x_final": xfinal0.numpy(), "features": feats0.numpy(),
            "layer_results": np.stack([t[0].numpy() for t in lr0]),  # [n+1, T, B, D]
        }
        if pm0 is not None:
            out["frame_padding_mask"] = pm0.numpy()
        # conv extractor output alone
        with torch.no_grad():
            out["conv_out"] = m.feature_extractor(wav).numpy()
        # 3) masked forward + gradients through the patched encoder (mask indices injected, not sampled)
        if grads:
            m.zero_grad()
            feats = m.feature_extractor(wav)
            feats = m.layer_norm(feats.transpose(1, 2))
            fpm = m.forward_padding_mask(feats, pm) if pm is not None else None
            x = m.post_extract_proj(feats) if m.post_extract_proj is not None else feats
            if mask_idx is not None:
                x = torch.where(mask_idx.unsqueeze(-1), m.mask_emb, x)
            x, _ = m.encoder(x, padding_mask=fpm)
            loss = O.probe_loss(x, fpm, seed=2)
            loss.backward()
            out["masked_x"] = x.detach().numpy()
            out["loss"] = np.array(loss.item(), dtype=np.float64)
            if mask_idx is not None:
                out["mask_indices"] = mask_idx.numpy()
            gsum = {}
            for k, p in m.named_parameters():
                if p.grad is not None:
                    gsum[k] = p.grad
            # full gradients for a handful of small / structurally interesting parameters, norms for all
            keep = [k for k in gsum if any(s in k for s in (
                "relative_attention_bias", "grep_", "mask_emb", "weight_g", "pos_conv.0.bias",
                "layers.0.fc1.bias", "layers.1.self_attn.q_proj.bias", "conv_layers.0.0.weight",
                "conv_layers.0.2", "layer_norm.weight", "post_extract_proj.bias"))]
            for k in keep:
                out["grad:" + k] = gsum[k].numpy()
            out["grad_norm_keys"] = np.array(sorted(gsum.keys()))
            out["grad_norms"] = np.array([gsum[k].double().norm().item() for k in sorted(gsum.keys())])
    finally:
        ref.TransformerEncoder.extract_features = orig
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: v for k, v in out.items()})
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith("grad:")},
          os.path.getsize(path) // 1024, "KiB")


def main():
    os.makedirs(OUT, exist_ok=True)
    # tiny post-LN / GroupNorm extractor ("Base-like"), ragged batch with padding mask + masking
    run_case("tiny_postln_ragged", O.tiny_config(pre_ln=False), B=2, L=8000, lengths=[8000, 5000], with_mask=True)
    # tiny pre-LN / LayerNorm extractor ("Large-like"), ragged
    run_case("tiny_preln_ragged", O.tiny_config(pre_ln=True), B=2, L=6400, lengths=[6400, 4321], with_mask=True)
    # tiny without padding mask, odd length (L % T != 0 path of forward_padding_mask not taken; extra samples)
    run_case("tiny_postln_nomask", O.tiny_config(pre_ln=False), B=1, L=7777, lengths=None, with_mask=False)
    # rel-pos / gate switched off (UniSpeech-SAT shipped cfg has relative_position_embedding False)
    run_case("tiny_preln_norelpos", O.tiny_config(pre_ln=True, relative_position_embedding=False, gru_rel_pos=False),
             B=2, L=4000, lengths=[4000, 3000], with_mask=False)
    # real WavLM-Base dims, 2 layers, 0.5 s: final output only (keeps the fixture small)
    run_case("base2l_halfsec", O.base_config(encoder_layers=2), B=1, L=8000, grads=False)
    # real WavLM-Large dims, 2 layers
    run_case("large2l_halfsec", O.large_config(encoder_layers=2), B=1, L=8000, grads=False)


if __name__ == "__main__":
    main()
