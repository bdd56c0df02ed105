#!/bin/bash
# Run on the GPU box (through gpurun): bench lines, ncu launch list + GEMM DRAM traffic of ONE warmed-up bench step, and
# `--set full` captures of the dominant kernels.  Outputs land in gpurun_out/ (summaries are copied into profiles/ afterwards
# with tools/ncu_hot.py / tools/ncu_gemm_traffic.py).  Usage: bash tools/profile_gpu.sh [tag]
TAG=${1:-x}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_base_$TAG.json 2> gpurun_out/bench_base_$TAG.err
python bench.py --model large --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err
python tools/bench_gemm.py --reps 20 > gpurun_out/bench_gemm_$TAG.txt 2>&1
python tools/bench_attn.py > gpurun_out/bench_attn_$TAG.txt 2>&1
# launch list of one step (cold-cache, serialised times: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:gemm_bf16 --csv --log-file gpurun_out/gemm_traffic_$TAG.csv \
    python bench.py --ncu-step --warmup 3 > /dev/null 2>&1
# full captures: one launch each
cap() {  # name regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_${name}_$TAG "$@" > /dev/null 2>&1
}
cap gemm_fc1 gemm_bf16_pair 3 python tools/bench_gemm.py --reps 1 --only fc1
cap attn_fwd attn_fwd 1 python tools/bench_attn.py --reps 1 --only base
cap attn_bwd_fused attn_bwd_fused 1 python tools/bench_attn.py --reps 1 --only base
cap ln_fwd ln_fwd_kernel 5 python bench.py --ncu-step --warmup 3
cap ln_bwd ln_bwd_kernel 5 python bench.py --ncu-step --warmup 3
cap conv0_fwd conv0_gn_fwd_apply 0 python bench.py --ncu-step --warmup 3
cap conv0_bwd conv0_gn_bwd_pass 0 python bench.py --ncu-step --warmup 3
ls -la gpurun_out | tail -20
