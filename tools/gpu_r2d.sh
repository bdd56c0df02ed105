#!/bin/bash
# Round-2 GPU call D: full GPU suite, reference arm, bench line without `also`.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_$TAG.log | cut -c1-220
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref_$TAG.json; tail -2 gpurun_out/bench_ref_$TAG.err
timeout 600 python bench.py --no-also > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_$TAG.log
