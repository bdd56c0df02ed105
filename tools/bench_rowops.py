"""Micro-benchmark of the HBM-bound kernels added for training: dropout row kernel and the fused Adam step / gradient norm,
at the WavLM-Base and -Large sizes.  Reports achieved GB/s on the ALGORITHMIC bytes against MEASURED_PEAKS.json hbm_gbs.
    python tools/bench_rowops.py [--reps 20] [--only base|large]"""
import argparse, json, os, struct, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unispeech_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only", default=None)
ap.add_argument("--norms", action="store_true", help="LayerNorm / gate / column-sum kernels only (skip dropout and Adam)")
args = ap.parse_args()
dev = torch.device("cuda:0")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6500.0


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, B, T, D, F, nparam in (("base", 16, 749, 768, 3072, 94_381_936), ("large", 8, 999, 1024, 4096, 315_453_120)):
    if args.only and args.only != name:
        continue
    x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
    res = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    h = torch.randn(B, T, F, device=dev).to(torch.bfloat16)
    rows = [
        ("dropout_rows [B,T,D] y=drop(x)", lambda: ops.dropout_rows(x, T * D, D, None, 0, 0, y, T * D, D, T, B, D, 0.1, (1, 2)), 2 * 2 * B * T * D),
        ("dropout_rows [B,T,D] y=res+drop(x)", lambda: ops.dropout_rows(x, T * D, D, res, T * D, D, y, T * D, D, T, B, D, 0.1, (1, 2)), 3 * 2 * B * T * D),
        ("dropout_rows [B,T,F] in place", lambda: ops.dropout_rows(h, T * F, F, None, 0, 0, h, T * F, F, T, B, F, 0.1, (1, 2)), 2 * 2 * B * T * F),
    ]
    H = D // 64
    f32 = lambda *sh: torch.randn(*sh, device=dev)
    gam, bet, mean, rstd = f32(D), f32(D), f32(B * T), f32(B * T).abs() + 0.5
    dgam, dbet, csum = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    gw, gb, ga_ = f32(8, 64) * 0.1, f32(8) * 0.1, f32(H)
    gate, dgate = torch.empty(B, H, T, device=dev), f32(B, H, T)
    dgw, dgb, dga = torch.zeros(8, 64, device=dev), torch.zeros(8, device=dev), torch.zeros(H, device=dev)
    qkv = torch.randn(B * T, 3 * D, device=dev).to(torch.bfloat16)
    csum3 = torch.zeros(3 * D, device=dev)
    E = 2 * B * T * D
    rows += [
        ("layer_norm_fwd", lambda: ops.layer_norm_fwd(x, T * D, D, gam, bet, y, T * D, D, mean, rstd, T, B, D), 2 * E),
        ("layer_norm_gate_fwd", lambda: ops.layer_norm_gate_fwd(x, T * D, D, gam, bet, y, T * D, D, mean, rstd, T, B, D, gw, gb, ga_, H, gate), 2 * E),
        ("layer_norm_bwd (+dres, colsum)", lambda: ops.layer_norm_bwd(x, T * D, D, res, T * D, D, mean, rstd, gam, bet, res, T * D, D, y, T * D, D, dgam, dbet, csum, T, B, D), 4 * E),
        ("layer_norm_bwd (plain)", lambda: ops.layer_norm_bwd(x, T * D, D, res, T * D, D, mean, rstd, gam, bet, None, 0, 0, y, T * D, D, dgam, dbet, None, T, B, D), 3 * E),
        ("gate_bwd", lambda: ops.gate_bwd(x, T * D, D, T, B, H, gw, gb, ga_, dgate, y, T * D, D, dgw, dgb, dga), 2 * E),
        ("colsum [B*T, D]", lambda: ops.colsum(x, T * D, D, T, B, D, csum), E),
        ("colsum [B*T, 3D]", lambda: ops.colsum(qkv, 0, 3 * D, B * T, 1, 3 * D, csum3), 3 * E),
    ]
    # one LayerNorm (+GELU) of the `layer_norm` conv stack: conv layer 1 of the large workload (rows = B x 31999 at 20 s, 512 channels)
    Tc, Cc = (16000 * (20 if name == "large" else 15) // 320 * 32) - 1, 512
    xc = torch.randn(B, Tc, Cc, device=dev).to(torch.bfloat16)
    yc, dc = torch.empty_like(xc), torch.randn(B, Tc, Cc, device=dev).to(torch.bfloat16)
    gc, bc, mc, rc = f32(Cc), f32(Cc), f32(B * Tc), f32(B * Tc).abs() + 0.5
    dgc, dbc = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    Ec = 2 * B * Tc * Cc
    rows += [
        (f"conv-stack LN+GELU fwd [{B}x{Tc}, 512]", lambda: ops.layer_norm_fwd(xc, Tc * Cc, Cc, gc, bc, yc, Tc * Cc, Cc, mc, rc, Tc, B, Cc, True), 2 * Ec),
        (f"conv-stack LN+GELU bwd [{B}x{Tc}, 512]", lambda: ops.layer_norm_bwd(dc, Tc * Cc, Cc, xc, Tc * Cc, Cc, mc, rc, gc, bc, None, 0, 0, yc, Tc * Cc, Cc, dgc, dbc, None, Tc, B, Cc, True), 3 * Ec),
    ]
    if args.norms:
        rows = rows[3:]
        for label, fn, nbytes in rows:
            ms = timeit(fn, args.reps)
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(f"{name:6s} {label:38s} {ms*1e3:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.0f} GB/s  = {gbs/peak:5.2f} of measured HBM peak {peak:.0f}", flush=True)
        continue
    # Adam over one flat tensor of the model's parameter count (the table has one record per parameter tensor in the model;
    # here a handful of large records: the kernel's cost is per element)
    n = (nparam + 3) // 4 * 4
    p_ = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev) * 0.01
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    parts, recs, chunks, off = 16, [], 0, 0
    per = n // parts // 4 * 4
    for i in range(parts):
        ne = per if i < parts - 1 else n - off
        recs.append(struct.pack("<Qqqq", p_.data_ptr() + 4 * off, off, ne, chunks))
        chunks += (ne + 2047) // 2048
        off += ne
    table = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev)
    step = [0]

    def adam():
        step[0] += 1
        ops.adam_step(table, parts, chunks, g, m, v, sumsq, 1.0, 1.0, 1e-4, 0.9, 0.98, 1e-6, 0.01, step[0], False)

    rows.append((f"sumsq_f32 ({n/1e6:.0f} M grads)", lambda: ops.sumsq_f32(g, n, sumsq), 4 * n))
    rows.append((f"adam_step ({n/1e6:.0f} M params)", adam, 28 * n))
    for label, fn, nbytes in rows:
        ms = timeit(fn, args.reps)
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(f"{name:6s} {label:38s} {ms*1e3:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.0f} GB/s  = {gbs/peak:5.2f} of measured HBM peak {peak:.0f}", flush=True)
