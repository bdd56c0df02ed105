#!/bin/bash
# Round-2 GPU call H: full GPU suite, stream-K A/B on the weight-gradient shapes, ragged + default bench lines.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_$TAG.log | cut -c1-220
G=gpurun_out/microbench_gemm_large_$TAG.txt
echo "== stream-K (default)" > $G; timeout 300 python tools/bench_gemm.py --large >> $G 2>&1
echo "== B200S_WGRAD_STREAMK=0" >> $G; B200S_WGRAD_STREAMK=0 timeout 300 python tools/bench_gemm.py --large --only wgrad_qkv,wgrad_o,wgrad_fc1,wgrad_fc2 >> $G 2>&1
cat $G
timeout 600 python bench.py --ragged --no-also > gpurun_out/bench_ragged_$TAG.json 2> gpurun_out/bench_ragged_$TAG.err; echo "ragged exit $?"; cut -c1-400 gpurun_out/bench_ragged_$TAG.json; tail -2 gpurun_out/bench_ragged_$TAG.err
timeout 600 python bench.py --no-also > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
