"""Micro-benchmark of the attention kernels at the WavLM-Base (16 x 749, 12 heads) and -Large (8 x 999, 16 heads) shapes.
    python tools/bench_attn.py [--reps 10] [--only base|large] [--dropout 0.1]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unispeech_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default=None)
ap.add_argument("--dropout", type=float, default=0.0, help="also time the kernels with dropout on the probabilities")
args = ap.parse_args()
dev = torch.device("cuda:0")
for name, B, T, H in (("base", 16, 749, 12), ("large", 8, 999, 16)):
    if args.only and args.only != name:
        continue
    D = H * 64
    torch.manual_seed(0)
    qkv = torch.randn(B, T, 3 * D, device=dev).to(torch.bfloat16)
    gate = torch.rand(B, H, T, device=dev) * 2 + 0.2
    tab = torch.randn(H, 2 * T - 1, device=dev)
    pad = torch.zeros(B, T, device=dev, dtype=torch.uint8)
    out = torch.empty(B, T, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev)
    dout = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
    delta = torch.empty(B, H, T, device=dev)
    dqkv = torch.zeros(B, T, 3 * D, device=dev, dtype=torch.bfloat16)
    dgate = torch.zeros(B, H, T, device=dev)
    dtab = torch.zeros(H, 2 * T - 1, device=dev)
    dq_acc = torch.zeros(B, T, D, device=dev)
    fns = {
        "fwd": lambda: ops.attn_fwd(qkv, gate, tab, pad, out, lse, B, T, H, 0.125),
        "bwd_fused": lambda: ops.attn_bwd_fused(qkv, out, dout, gate, tab, pad, lse, delta, dq_acc, dqkv, dgate, dtab, B, T, H, 0.125),
        "bwd_2kernel": lambda: ops.attn_bwd(qkv, out, dout, gate, tab, pad, lse, delta, dqkv, dgate, dtab, B, T, H, 0.125),
    }
    # the same kernels without the gated relative-position bias (HuBERT / wav2vec2 encoders): what the bias path costs
    fns["fwd_nobias"] = lambda: ops.attn_fwd(qkv, None, None, pad, out, lse, B, T, H, 0.125)
    fns["bwd_fused_nobias"] = lambda: ops.attn_bwd_fused(qkv, out, dout, None, None, pad, lse, delta, dq_acc, dqkv, None, None, B, T, H, 0.125)
    if args.dropout > 0:
        words = torch.empty(ops.attn_dropout_mask_words(B, T, H), dtype=torch.int32, device=dev)
        fns["fwd_dropout"] = lambda: ops.attn_fwd_dropout(qkv, gate, tab, pad, out, lse, B, T, H, 0.125, args.dropout, (123, 456), words)
        fns["bwd_fused_dropout"] = lambda: ops.attn_bwd_fused_dropout(qkv, out, dout, gate, tab, pad, lse, delta, dq_acc, dqkv, dgate,
                                                                      dtab, B, T, H, 0.125, args.dropout, words)
    fl = 4.0 * B * H * T * T * 64
    for k, fn in fns.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        mult = 1.0 if k.startswith("fwd") else 2.5
        print(f"{name:6s} {k:18s} {ms*1e3:9.1f} us   {fl*mult/ms/1e9:8.1f} TFLOP/s (algorithmic)", flush=True)
