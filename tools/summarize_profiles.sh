#!/bin/bash
# Turn the gpurun_out/ captures of tools/profile_gpu.sh into the text summaries kept under profiles/.  Usage: bash tools/summarize_profiles.sh TAG OUTTAG
TAG=$1; OUT=${2:-$1}
for name in gemm_fc1 attn_fwd attn_bwd_fused ln_fwd ln_bwd conv0_fwd conv0_bwd; do
  rep=gpurun_out/prof_${name}_$TAG.ncu-rep
  [ -f $rep ] || continue
  ( echo "# ncu --set full --clock-control none --import-source on, one launch: $name ($TAG)"
    ncu -i $rep --page details 2>/dev/null | grep -E "^  [a-z_:A-Z<>, ()0-9]+\(|Duration|Elapsed Cycles|SM Frequency|Executed Ipc|Issue Slots Busy|highest-utilized|Memory Throughput|L2 Cache Throughput|DRAM Throughput|No Eligible|L2 Hit|Registers Per|Dynamic Shared|Cluster Size|Block Size|Grid Size|bank conflicts|Achieved Occupancy"
    ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for h,u,v in zip(rows[0],rows[1],rows[2]):
    if h in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','smsp__inst_executed.sum','dram__bytes_read.sum.per_second','dram__bytes_write.sum.per_second','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread'): print(h,u,v)
"
    python tools/ncu_hot.py $rep 20 ) > profiles/r01_ncu_${name}_$OUT.txt
done
ls profiles
