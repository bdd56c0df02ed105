#!/bin/bash
# Final-check GPU call: the driver's own commands (GPU test suite with -x, smoke, default bench), the Base line, a quick
# attention-dropout micro-benchmark, a launch list of one Large step and three ncu captures of the reworked memory-bound kernels.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 500 python bench.py > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
timeout 200 python bench.py --model base --no-also > gpurun_out/bench_base_$TAG.json 2> gpurun_out/bench_base_$TAG.err; echo "bench base exit $?"; cut -c1-250 gpurun_out/bench_base_$TAG.json
timeout 100 python tools/bench_attn.py --reps 10 --dropout 0.1 > gpurun_out/bench_attn_$TAG.txt 2>&1; cat gpurun_out/bench_attn_$TAG.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_large_$TAG.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1; echo "launch list exit $?"
cap() {  # name regex skip
  local name=$1 rx=$2 skip=$3
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -s $skip -c 1 \
      -o gpurun_out/prof_${name}_$TAG python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${name}_$TAG.log 2>&1
  echo "ncu $name exit $?"
}
cap ln_bwd2 ln_bwd_kernel 10
cap gate_bwd2 gate_bwd_kernel 3
cap conv0_dconv conv0_ln_bwd_dconv_kernel 0
ls -la gpurun_out | tail -16
