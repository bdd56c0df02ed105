"""Component-wise error report of the fused attention backward against autograd (run on the GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from unispeech_b200 import ops
from test_kernels_gpu import _attn_ref, bf

dev = torch.device("cuda:0")
for (B, T, H, bias, padded) in [(1, 64, 1, False, False), (1, 128, 1, False, False), (1, 128, 1, True, False), (2, 100, 2, True, True), (1, 300, 2, True, False), (1, 749, 3, True, False)]:
    torch.manual_seed(T + 1)
    D = H * 64
    qkv = bf(torch.randn(B, T, 3 * D, device=dev))
    gate = (torch.rand(B, H, T, device=dev) * 2 + 0.2) if bias else None
    tab = torch.randn(H, 2 * T - 1, device=dev) if bias else None
    pad = None
    if padded:
        pad = torch.zeros(B, T, device=dev, dtype=torch.uint8)
        pad[0, T - T // 3:] = 1
    out = torch.empty(B, T, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev)
    ops.attn_fwd(qkv, gate, tab, pad, out, lse, B, T, H, 0.125)
    dout = bf(torch.randn(B, T, D, device=dev))
    res = {}
    for fused in (False, True):
        delta = torch.empty(B, H, T, device=dev)
        dqkv = torch.zeros(B, T, 3 * D, device=dev, dtype=torch.bfloat16)
        dgate = torch.zeros(B, H, T, device=dev) if bias else None
        dtab = torch.zeros(H, 2 * T - 1, device=dev) if bias else None
        if fused:
            dq_acc = torch.zeros(B, T, D, device=dev)
            ops.attn_bwd_fused(qkv, out, dout, gate, tab, pad, lse, delta, dq_acc, dqkv, dgate, dtab, B, T, H, 0.125)
        else:
            ops.attn_bwd(qkv, out, dout, gate, tab, pad, lse, delta, dqkv, dgate, dtab, B, T, H, 0.125)
        torch.cuda.synchronize()
        res[fused] = (dqkv.float(), dgate, dtab)
    qr = qkv.float().requires_grad_(True)
    gr = gate.clone().requires_grad_(True) if bias else None
    tr = tab.clone().requires_grad_(True) if bias else None
    ref = _attn_ref(qr, gr, tr, pad, B, T, H, 0.125)
    ref.backward(dout.float())
    g = qr.grad
    line = f"B{B} T{T} H{H} bias{int(bias)} pad{int(padded)}:"
    for fused in (False, True):
        dqkv, dgate, dtab = res[fused]
        e = [(dqkv[..., i * D:(i + 1) * D] - g[..., i * D:(i + 1) * D]).abs().max().item() for i in range(3)]
        line += f"  [{'fused' if fused else 'old'}] dq {e[0]:.4f} dk {e[1]:.4f} dv {e[2]:.4f} (scale {g.abs().max().item():.2f})"
        if bias:
            line += f" dgate {(dgate - gr.grad).abs().max().item():.4f}/{gr.grad.abs().max().item():.2f} dtab {(dtab - tr.grad).abs().max().item():.4f}/{tr.grad.abs().max().item():.2f}"
    print(line, flush=True)
    if T <= 128 and not bias:
        dq_f = res[True][0][..., :D]
        bad = (dq_f - g[..., :D]).abs()
        print("   dq err by row block of 16:", [round(bad[0, i:i + 16].max().item(), 3) for i in range(0, T, 16)])
        print("   dq err by col block of 8:", [round(bad[0, :, i:i + 8].max().item(), 3) for i in range(0, 64, 8)])
