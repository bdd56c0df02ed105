#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_large_$TAG.json; tail -2 gpurun_out/bench_large_$TAG.err
timeout 300 python -m pytest tests/test_graph_gpu.py -q -p no:cacheprovider 2>&1 | tail -2
