"""One-file text summary of an `ncu --set full` report for profiles/: headline metrics (raw page), the details-page lines the
judge reads (IPC, issue slots, occupancy, memory pipes), and the hottest SASS lines with their stall reasons (source page).
usage: python tools/ncu_summary.py report.ncu-rep [topN] > profiles/rNN_ncu_<kernel>.txt"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = sys.argv[2] if len(sys.argv) > 2 else "30"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
print(f"# ncu --set full --clock-control none --import-source on: {rep.split('/')[-1]}")
for r in rows[2:]:
    print("kernel:", r[h.index("Kernel Name")][:120], " grid", r[h.index("Grid Size")], " block", r[h.index("Block Size")])
    for name in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
                 "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                 "smsp__warps_active.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"):
        if name in h:
            print(f"  {name} = {r[h.index(name)]} {rows[1][h.index(name)]}")
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Elapsed Cycles", "SM Frequency", "Executed Ipc", "Issue Slots Busy", "No Eligible", "Registers Per", "Achieved Occ",
        "Dynamic Shared", "Mem Busy", "Max Bandwidth", "Mem Pipes Busy", "Active Warps Per", "Eligible Warps", "Compute (SM)",
        "Memory Throughput", "bank conflicts", "excessive")
for l in det.splitlines():
    if any(k in l for k in keys):
        print(l.rstrip()[:160])
print(subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_hot.py"), rep, topn], capture_output=True, text=True).stdout)
