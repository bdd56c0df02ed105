"""The reference's fine-tuning wrappers around the encoder (HubertEncoder / Wav2VecEncoder) on the GPU,
WITHOUT stubs: real `unispeech_b200.WavLM`, `final_dropout` and the output projection `proj` on the kernels, forward and backward
against the fp32 oracle (same weights, same counter-based dropout mask).  src/fairseq/models/hubert/hubert_asr.py:314-340,
src/fairseq/models/wav2vec/wav2vec2_asr.py:390-421."""
import pytest
import torch
import torch.nn.functional as F

from oracle import wavlm_oracle as O

pytestmark = pytest.mark.gpu
SITE_FINAL = 0x7F000001


def _oracle_tail(x_btc, w, b, p, seed):
    if p > 0:
        x_btc = O.HashDropout(seed).rows_btc(SITE_FINAL, x_btc, p)
    return F.linear(x_btc, w, b) if w is not None else x_btc


@pytest.mark.parametrize("kind,pre_ln,V,p", [("hubert", False, 40, 0.1), ("wav2vec", True, 32, 0.0), ("hubert", True, None, 0.2)])
def test_encoder_wrapper_forward_backward(cuda_device, kind, pre_ln, V, p):
    from unispeech_b200.fairseq_encoder import HubertEncoder, Wav2VecEncoder
    from unispeech_b200.wavlm import WavLM, WavLMConfig
    dev = cuda_device
    cfg = O.tiny_config(pre_ln=pre_ln)
    sd = O.deterministic_state_dict(cfg)
    m = WavLM(WavLMConfig(vars(cfg)))
    m.load_state_dict(sd, strict=True)
    Enc = HubertEncoder if kind == "hubert" else Wav2VecEncoder
    enc = Enc(m, apply_mask=False, final_dropout=p, output_dim=V).to(dev).train()
    enc.dropout_seed = 77
    D = cfg.encoder_embed_dim
    if V is not None:
        pw = O.hash_uniform("wrap.w", (V, D), -0.1, 0.1)
        pb = O.hash_uniform("wrap.b", (V,), -0.1, 0.1)
        with torch.no_grad():
            enc.proj.weight.copy_(pw)
            enc.proj.bias.copy_(pb)
    wav, pmask = O.deterministic_waveform(2, 8000, seed=1, lengths=[8000, 6100])
    out = enc(wav.to(dev), pmask.to(dev))
    y = out["encoder_out"]  # T x B x C'
    T = O.num_frames(8000, cfg)
    C_out = V if V is not None else D
    assert y.shape == (T, 2, C_out)
    if kind == "hubert":
        assert out["encoder_padding_mask"].shape == (2, T) and out["padding_mask"] is out["encoder_padding_mask"]
    else:
        assert out["encoder_padding_mask"].shape == (T, 2) and out["padding_mask"].shape == (2, T)
    # oracle
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    wr = pw.clone().requires_grad_(True) if V is not None else None
    br = pb.clone().requires_grad_(True) if V is not None else None
    ref = O.extract_features(sdr, wav, cfg, padding_mask=pmask)
    fpm = ref["padding_mask"]
    want = _oracle_tail(ref["x"], wr, br, p, 77).transpose(0, 1)
    valid = (~fpm).t()
    d = (y.detach().float().cpu() - want.detach())[valid]
    assert torch.isfinite(y.float()).all()
    assert d.abs().max().item() < 0.12 and d.abs().mean().item() < 0.02, (d.abs().max().item(), d.abs().mean().item())
    # backward through the wrapper into the encoder and the projection
    R = O.hash_uniform("wrap.R", tuple(want.shape)).masked_fill(~valid.unsqueeze(-1), 0.0)
    (y.float() * R.to(dev)).sum().backward()
    (want * R).sum().backward()
    torch.cuda.synchronize()
    checks = {"encoder.layers.0.fc1.weight": (dict(m.named_parameters())["encoder.layers.0.fc1.weight"].grad, sdr["encoder.layers.0.fc1.weight"].grad),
              "post_extract_proj.weight": (dict(m.named_parameters())["post_extract_proj.weight"].grad, sdr["post_extract_proj.weight"].grad)}
    if V is not None:
        checks["proj.weight"] = (enc.proj.weight.grad, wr.grad)
        checks["proj.bias"] = (enc.proj.bias.grad, br.grad)
    for k, (got, ref_g) in checks.items():
        got, ref_g = got.detach().double().cpu(), ref_g.double()
        cos = ((got * ref_g).sum() / (got.norm() * ref_g.norm() + 1e-30)).item()
        rel = abs(got.norm().item() - ref_g.norm().item()) / ref_g.norm().item()
        assert cos > 0.995 and rel < 0.06, (k, cos, rel)


def test_frozen_encoder_gets_no_gradient(cuda_device):
    """`freeze_finetune_updates`: until that many updates the encoder runs under no_grad and only `proj` trains (hubert_asr.py:322-326)."""
    from unispeech_b200.fairseq_encoder import HubertEncoder
    from unispeech_b200.wavlm import WavLM, WavLMConfig
    cfg = O.tiny_config()
    m = WavLM(WavLMConfig(vars(cfg)))
    m.load_state_dict(O.deterministic_state_dict(cfg), strict=True)
    enc = HubertEncoder(m, freeze_finetune_updates=5, output_dim=16).to(cuda_device).train()
    wav, _ = O.deterministic_waveform(1, 6400, seed=2)
    out = enc(wav.to(cuda_device), None)
    out["encoder_out"].float().pow(2).sum().backward()
    assert enc.proj.weight.grad is not None and float(enc.proj.weight.grad.abs().sum()) > 0
    assert all(p.grad is None or float(p.grad.abs().sum()) == 0.0 for p in m.parameters())
