#!/bin/bash
# Round-2 GPU call Q: wav2vec 2.0 head tests, then the whole GPU suite.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_w2v_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_w2v_$TAG.log 2>&1; echo "w2v pytest exit $?"; tail -15 gpurun_out/pytest_w2v_$TAG.log | cut -c1-300
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
