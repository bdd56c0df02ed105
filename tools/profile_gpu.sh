#!/bin/bash
# Run on the GPU box (through gpurun): GEMM micro-benchmark, ncu launch list of ONE warmed-up bench step, and one
# `--set full` capture of the dominant GEMM.  Outputs land in gpurun_out/ (copy the summaries you keep into profiles/).
set -x
mkdir -p gpurun_out
python tools/bench_gemm.py --reps 10 2>&1 | tee gpurun_out/bench_gemm.txt
# launch list of one step (cold-cache, serialised times: compare SHARES, not absolutes)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --ncu-step --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
wc -l gpurun_out/launches.csv
# one full capture of the dominant GEMM (fc1 shape, pair kernel)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/prof_gemm_fc1 \
    python tools/bench_gemm.py --reps 1 --only fc1 > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
ls -la gpurun_out
