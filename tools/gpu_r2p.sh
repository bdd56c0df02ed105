#!/bin/bash
# Round-2 GPU call P: ILS-HuBERT tests, graph test, compute-sanitizer passes over the small-shape tests.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ils_gpu.py tests/test_graph_gpu.py tests/test_pretrain_gpu.py tests/test_sat_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
bash tools/sanitize.sh $TAG
