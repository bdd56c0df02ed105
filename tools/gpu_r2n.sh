#!/bin/bash
# Round-2 GPU call N: tensor-core diagonal sums in the attention backward (tests + micro-benchmark), forward-determinism probe.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py tests/test_fullscale_gpu.py tests/test_ragged_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 200 python tools/bench_attn.py --reps 10 --dropout 0.1 > gpurun_out/microbench_attn_$TAG.txt 2>&1; cat gpurun_out/microbench_attn_$TAG.txt
timeout 200 python tools/debug_determinism.py > gpurun_out/determinism_$TAG.txt 2>&1; cat gpurun_out/determinism_$TAG.txt | cut -c1-400
timeout 200 python tools/debug_determinism.py --large >> gpurun_out/determinism_$TAG.txt 2>&1; tail -6 gpurun_out/determinism_$TAG.txt | cut -c1-400
