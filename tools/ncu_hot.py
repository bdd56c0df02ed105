"""Summarise the source page of an ncu report: top SASS instructions by stall samples + stall-reason totals.
usage: python tools/ncu_hot.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# possibly several kernels: split on "Kernel Name" rows
blocks, cur = [], None
for l in lines:
    if l.startswith('"Kernel Name"'):
        cur = [l]; blocks.append(cur)
    elif cur is not None:
        cur.append(l)
for b in blocks:
    print(b[0][:200])
    rows = list(csv.DictReader(b[1:]))
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    reasons = collections.Counter()
    for r in rows:
        for k, v in r.items():
            if k.startswith("stall_") and "Not Issued" not in k and v:
                reasons[k] += int(v)
    print("total samples", tot, " reasons:", ", ".join(f"{k[6:]}={v}" for k, v in reasons.most_common(10)))
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:topn]
    for i in sorted(idx):
        r = rows[i]
        top = sorted(((int(v), k[6:]) for k, v in r.items() if k.startswith("stall_") and "Not Issued" not in k and v and int(v) > 0), reverse=True)[:3]
        print(f"{i:5d} {int(r['# Samples']):6d} {100*int(r['# Samples'])/max(tot,1):5.1f}%  {r['Source'].strip()[:90]:90s} {top}")
