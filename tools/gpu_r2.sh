#!/bin/bash
# Round-2 GPU call: attention kernel tests + micro-benchmark first (new kernels), then the full-scale parity tests, the rest of
# the GPU suite, and a bench line.  Every step has its own timeout; logs land in gpurun_out/ with the given tag.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x -k "attn" -p no:cacheprovider > gpurun_out/attn_tests_$TAG.log 2>&1; echo "attn tests exit $?"; tail -6 gpurun_out/attn_tests_$TAG.log
timeout 200 python tools/bench_attn.py --reps 10 --dropout 0.1 > gpurun_out/bench_attn_$TAG.txt 2>&1; cat gpurun_out/bench_attn_$TAG.txt
timeout 900 python -m pytest tests/test_fullscale_gpu.py -q -s -p no:cacheprovider > gpurun_out/fullscale_$TAG.log 2>&1; echo "fullscale exit $?"; tail -8 gpurun_out/fullscale_$TAG.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_fullscale_gpu.py > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 500 python bench.py --no-also > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
