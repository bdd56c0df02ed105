#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards;
# synccheck: barrier misuse).  Tiny models keep each pass to a minute or two; the full-size kernels run the same code paths
# (tile loops are shape-generic).  Output: gpurun_out/sanitizer_<tool>_<tag>.log (summary lines are copied into profiles/).
TAG=${1:-x}
mkdir -p gpurun_out
SEL='tests/test_model_gpu.py::test_forward_vs_golden tests/test_model_gpu.py::test_gradients_vs_golden tests/test_kernels_gpu.py::test_ragged_row_kernels tests/test_gemm_pair_gpu.py'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest $SEL -x -q -p no:cacheprovider \
      -k "tiny or ragged or pair" > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
  echo "sanitizer $tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_$TAG.log | tail -4
done
