#!/bin/bash
# Round-2 GPU call C: full GPU suite (no -x), attention micro-benchmark, bench line (Large, with `also`).
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu_$TAG.log | cut -c1-220
timeout 200 python tools/bench_attn.py --reps 10 --dropout 0.1 > gpurun_out/bench_attn_$TAG.txt 2>&1; cat gpurun_out/bench_attn_$TAG.txt
timeout 900 python bench.py > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
