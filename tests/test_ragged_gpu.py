"""BASELINE.json configs[4] (WavLM-Large, variable-length batch with a padding mask) at FULL size through a size-independent
property the domain offers: an utterance's valid frames -- and the parameter gradients of a loss on those frames -- do not depend
on what else is in the batch or on how far the batch is padded (no BatchNorm; LayerNorm extractor; padded frames are zeroed
before pos_conv and masked as attention keys, WavLM/WavLM.py:311-321,574-575, WavLM/modules.py:549-556).  The CPU oracle cannot
run 24 x 1024 at T = 1499 in seconds; the small-size oracle parity of the same path is tests/test_model_gpu.py."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
SR = 16000


def test_large_full_size_batch_and_padding_invariance(cuda_device):
    from oracle import wavlm_oracle as O   # config + frame arithmetic only
    from unispeech_b200.wavlm import WavLM, WavLMConfig
    dev = cuda_device
    cfg = O.large_config()
    torch.manual_seed(11)
    m = WavLM(WavLMConfig(vars(cfg))).to(dev).train()      # dropout 0 in this config: train mode only enables the backward
    n1 = 160480                                             # 10.03 s
    La, Lb = 30 * SR, 15 * SR
    Ta, Tb = O.num_frames(La, cfg), O.num_frames(Lb, cfg)
    assert (Ta, Tb) == (1499, 749) and La // Ta == Lb // Tb == 320
    n_valid = math.ceil(n1 / 320)                           # frames the reference's frame mask keeps (forward_padding_mask)
    g = torch.Generator().manual_seed(5)
    u0 = torch.nn.functional.layer_norm(torch.randn(La, generator=g), (La,))
    u1 = torch.nn.functional.layer_norm(torch.randn(n1, generator=g), (n1,))
    A = torch.zeros(2, La)
    A[0], A[1, :n1] = u0, u1
    pmA = torch.zeros(2, La, dtype=torch.bool)
    pmA[1, n1:] = True
    Bw = torch.zeros(1, Lb)
    Bw[0, :n1] = u1
    pmB = torch.zeros(1, Lb, dtype=torch.bool)
    pmB[0, n1:] = True
    R = torch.randn(n_valid, cfg.encoder_embed_dim, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    names = ["encoder.layers.0.fc1.weight", "encoder.layers.23.self_attn.out_proj.weight", "encoder.layers.11.self_attn.q_proj.weight",
             "encoder.layers.0.self_attn.relative_attention_bias.weight", "encoder.layers.5.self_attn.grep_linear.weight",
             "encoder.pos_conv.0.weight_v", "post_extract_proj.weight", "feature_extractor.conv_layers.1.0.weight",
             "feature_extractor.conv_layers.0.0.weight", "encoder.layer_norm.weight"]
    params = dict(m.named_parameters())

    def run(wav, pm, row):
        if m._engine is not None and m._engine.flat is not None:
            m.grad_buffer().zero_()
        x, fpm = m.extract_features(wav.to(dev), padding_mask=pm)
        assert int((~fpm[row]).sum()) == n_valid
        loss = (x[row, :n_valid].float() * R).sum()
        loss.backward()
        torch.cuda.synchronize()
        return x[row, :n_valid].detach().float(), loss.item(), {k: params[k].grad.detach().double().clone() for k in names}

    xa, la, ga = run(A, pmA, 1)
    xb, lb, gb = run(Bw, pmB, 0)
    assert torch.isfinite(xa).all() and torch.isfinite(xb).all()
    d = (xa - xb).abs()
    scale = xa.abs().max().item()
    assert d.max().item() < 0.06 * max(1.0, scale) and d.mean().item() < 0.01, (d.max().item(), d.mean().item(), scale)
    assert abs(la - lb) < 0.02 * abs(lb) + 1.0, (la, lb)
    bad = []
    for k in names:
        cos = ((ga[k] * gb[k]).sum() / (ga[k].norm() * gb[k].norm() + 1e-30)).item()
        rel = abs(ga[k].norm().item() - gb[k].norm().item()) / (gb[k].norm().item() + 1e-30)
        if cos < 0.995 or rel > 0.05:
            bad.append((k, round(cos, 5), round(rel, 4)))
    assert not bad, bad
