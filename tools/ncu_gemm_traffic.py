"""Turn an ncu CSV log (metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the GEMM launches of ONE
bench step) into profiles/gemm_traffic_<model>.json, which bench.py reports as roofline.traffic.
usage: python tools/ncu_gemm_traffic.py gpurun_out/gemm_traffic.csv base "<command that produced it>"
"""
import csv, json, sys, collections
path, model, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(path).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
per = collections.defaultdict(dict)
names = {}
for r in csv.DictReader(lines[i:]):
    per[r["ID"]][r["Metric Name"]] = float(r["Metric Value"])
    names[r["ID"]] = r["Kernel Name"]
n = len(per)
rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in per.values())
wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in per.values())
ns = sum(v.get("gpu__time_duration.sum", 0.0) for v in per.values())
unit_scale = 1.0
out = {"model": model, "launches": n, "dram_bytes_read_per_step": rd, "dram_bytes_write_per_step": wr,
       "dram_bytes_per_launch": (rd + wr) / max(n, 1), "gpu_time_ns_sum_under_ncu": ns,
       "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the {n} tcgen05 GEMM launches of one step ({cmd})"}
json.dump(out, open(f"profiles/gemm_traffic_{model}.json", "w"), indent=1)
print(json.dumps(out, indent=1))
