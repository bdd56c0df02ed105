#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_w2v_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_w2v_$TAG.log 2>&1; echo "w2v exit $?"; grep -E "AssertionError|passed|failed" gpurun_out/pytest_w2v_$TAG.log | cut -c1-1500
