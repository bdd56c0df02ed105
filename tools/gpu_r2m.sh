#!/bin/bash
# Round-2 GPU call M: full GPU suite, smoke, the default bench line (with `also` and the CUDA-graph probe), attention micro-benchmark.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 1200 python bench.py > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
timeout 200 python tools/bench_attn.py --reps 10 --dropout 0.1 --only large > gpurun_out/microbench_attn_$TAG.txt 2>&1; cat gpurun_out/microbench_attn_$TAG.txt
