"""Which thread count gives the best CPU-oracle throughput on this host? (run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import wavlm_oracle as O
cfg = O.base_config()
sd = {k: v.clone().requires_grad_(True) for k, v in O.deterministic_state_dict(cfg).items()}
secs = 5
wav, _ = O.deterministic_waveform(1, secs * 16000, seed=3)
pm = torch.zeros(1, secs * 16000, dtype=torch.bool)
print("os.cpu_count", os.cpu_count(), "torch default threads", torch.get_num_threads(), "affinity", len(os.sched_getaffinity(0)))
for th in (8, 16, 32, 64, 128):
    torch.set_num_threads(th)
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        res = O.extract_features(sd, wav, cfg, padding_mask=pm)
        O.probe_loss(res["x"], res["padding_mask"], seed=2).backward()
        for v in sd.values():
            v.grad = None
        ts.append(time.perf_counter() - t0)
    print(th, "threads:", [round(t, 2) for t in ts], "s  ->", round(secs / min(ts), 2), "audio-s/s", flush=True)
