"""Fused optimizer step against the oracle's restatement of fairseq's clip_grad_norm_ + Adam."""
import struct

import pytest
import torch

from oracle import wavlm_oracle as O

pytestmark = pytest.mark.gpu


def _table(params, offsets, device):
    recs, chunks = [], 0
    for p, off in zip(params, offsets):
        recs.append(struct.pack("<Qqqq", p.data_ptr(), off, p.numel(), chunks))
        chunks += (p.numel() + 2047) // 2048
    return torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(device), len(recs), chunks


@pytest.mark.parametrize("max_norm,mult,wd,zero", [(0.0, 1.0, 0.0, False), (0.5, 1.0, 0.01, True), (1e6, 0.125, 0.1, False),
                                                  (2.0, 3.0, 0.0, True)])
def test_adam_step_kernel_vs_reference(cuda_device, max_norm, mult, wd, zero):
    from unispeech_b200 import ops
    dev = cuda_device
    torch.manual_seed(17)
    shapes = [(768, 768), (3,), (1,), (2049,), (512, 10), (4097,), (8, 64), (1, 12, 1, 1), (2048,), (5, 7, 3)]
    params = [torch.randn(s, device=dev) for s in shapes]
    offsets, total = [], 0
    for p in params:
        offsets.append(total)
        total += (p.numel() + 3) // 4 * 4
    g = torch.zeros(total, device=dev)
    m = torch.zeros(total, device=dev)
    v = torch.zeros(total, device=dev)
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    table, n, chunks = _table(params, offsets, dev)
    ref_p = [p.detach().cpu().clone() for p in params]
    state = {}
    lr, betas, eps = 3e-3, (0.9, 0.98), 1e-6
    for step in range(1, 4):
        grads = [torch.randn(s, device=dev) * (0.1 * step) for s in shapes]
        for gr, off in zip(grads, offsets):
            g[off:off + gr.numel()] = gr.reshape(-1)
        sumsq.zero_()
        ops.sumsq_f32(g, g.numel(), sumsq)
        ops.adam_step(table, n, chunks, g, m, v, sumsq if max_norm > 0 else None, mult, max_norm, lr, betas[0], betas[1], eps, wd,
                      step, zero)
        torch.cuda.synchronize()
        cg = [x.cpu() for x in grads]
        norm, coef = O.clip_coefficient(cg, max_norm, mult)
        assert abs(sumsq.sqrt().item() * abs(mult) - norm.item()) < 1e-4 * max(1.0, norm.item())
        O.adam_step(ref_p, [x * coef for x in cg], state, lr, betas, eps, wd)
        for p, r, s_ in zip(params, ref_p, shapes):
            assert torch.allclose(p.cpu(), r, rtol=2e-5, atol=2e-6), (s_, step, (p.cpu() - r).abs().max().item())
        for off, mr, vr in zip(offsets, state["exp_avg"], state["exp_avg_sq"]):
            assert torch.allclose(m[off:off + mr.numel()].cpu(), mr.reshape(-1), rtol=2e-5, atol=1e-7)
            assert torch.allclose(v[off:off + vr.numel()].cpu(), vr.reshape(-1), rtol=2e-5, atol=1e-9)
        if zero:
            assert g.abs().max().item() == 0.0
        else:
            assert g.abs().max().item() > 0.0


def test_fused_adam_on_model(cuda_device):
    """One optimisation step through the public surface: norm / clip / update match the restatement on the model's own
    gradients, and the next forward pass runs on the UPDATED parameters (bf16 operands re-derived)."""
    from unispeech_b200.optim import FusedAdam
    from unispeech_b200.wavlm import WavLM, WavLMConfig
    dev = cuda_device
    cfg = O.tiny_config(pre_ln=False)
    m = WavLM(WavLMConfig(vars(cfg)))
    m.load_state_dict(O.deterministic_state_dict(cfg), strict=True)
    m = m.to(dev).train()
    wav, pmask = O.deterministic_waveform(2, 8000, seed=1, lengths=[8000, 5000])
    x, fpm = m.extract_features(wav.to(dev), padding_mask=pmask.to(dev))
    O.probe_loss(x.float(), fpm, seed=2).backward()
    opt = FusedAdam(m, lr=1e-2, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    names = [k for k, p in m.named_parameters() if p.requires_grad]
    params = dict(m.named_parameters())
    before = {k: params[k].detach().cpu().clone() for k in names}
    grads = {k: params[k].grad.detach().cpu().clone() for k in names}
    opt.multiply_grads(0.5)
    gn = opt.clip_grad_norm(1.0)
    want_norm, coef = O.clip_coefficient([grads[k] for k in names], 1.0, 0.5)
    assert abs(gn.item() - want_norm.item()) < 1e-3 * want_norm.item()
    opt.step(zero_grad=True)
    torch.cuda.synchronize()
    ref = [before[k].clone() for k in names]
    O.adam_step(ref, [grads[k] * coef for k in names], {}, 1e-2, (0.9, 0.98), 1e-6, 0.01)
    for k, r in zip(names, ref):
        got = params[k].detach().cpu()
        assert torch.allclose(got, r, rtol=1e-4, atol=1e-5), (k, (got - r).abs().max().item())
        assert not torch.equal(got, before[k]) or grads[k].abs().max().item() == 0, k
    assert m.grad_buffer().abs().max().item() == 0.0
    # the next forward uses the updated masters
    sd_new = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        m.eval()
        y, _ = m.extract_features(wav.to(dev), padding_mask=pmask.to(dev))
        want = O.extract_features(sd_new, wav, cfg, padding_mask=pmask)
        old = O.extract_features(O.deterministic_state_dict(cfg), wav, cfg, padding_mask=pmask)
    valid = ~want["padding_mask"]
    err = (y.float().cpu() - want["x"])[valid].abs().max().item()
    moved = (old["x"] - want["x"])[valid].abs().max().item()
    assert err < 0.12, err
    assert moved > 2 * err, (moved, err)
