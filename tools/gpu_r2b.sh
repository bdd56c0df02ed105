#!/bin/bash
# Round-2 GPU call B: ncu captures of the rewritten attention kernels at the Large shape, then the test suite and a bench line.
TAG=${1:-x}
mkdir -p gpurun_out
cap() {  # name regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_${name}_$TAG "$@" > gpurun_out/ncu_${name}_$TAG.log 2>&1
  echo "ncu $name exit $?"
}
cap attn_fwd attn_fwd_kernel 2 python tools/bench_attn.py --reps 1 --only large
cap attn_bwd attn_bwd_fused_kernel 2 python tools/bench_attn.py --reps 1 --only large
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --no-also > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
