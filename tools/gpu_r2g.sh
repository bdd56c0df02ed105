#!/bin/bash
# Round-2 GPU call G: A/B of the LayerNorm kernel variants, GEMM epilogue exposure, kernel tests, default bench line.
TAG=${1:-x}
mkdir -p gpurun_out
O=gpurun_out/microbench_norms_$TAG.txt
echo "== wide, R=4" > $O; timeout 300 python tools/bench_rowops.py --norms --only large >> $O 2>&1
echo "== wide, R=2" >> $O; B200S_LN_R=2 timeout 300 python tools/bench_rowops.py --norms --only large >> $O 2>&1
echo "== warp-per-row kernels" >> $O; B200S_LN_WIDE=0 timeout 300 python tools/bench_rowops.py --norms --only large >> $O 2>&1
cat $O
G=gpurun_out/microbench_gemm_large_$TAG.txt
echo "== normal" > $G; timeout 300 python tools/bench_gemm.py --large >> $G 2>&1
echo "== no epilogue global traffic (B200S_GEMM_DEBUG=1)" >> $G; B200S_GEMM_DEBUG=1 timeout 300 python tools/bench_gemm.py --large >> $G 2>&1
echo "== no MMAs (B200S_GEMM_DEBUG=2)" >> $G; B200S_GEMM_DEBUG=2 timeout 300 python tools/bench_gemm.py --large >> $G 2>&1
echo "== no loads (B200S_GEMM_DEBUG=4)" >> $G; B200S_GEMM_DEBUG=4 timeout 300 python tools/bench_gemm.py --large >> $G 2>&1
for S in 2 3 4 6 8 12; do echo "== B200S_WGRAD_SPLITS=$S" >> $G; B200S_WGRAD_SPLITS=$S timeout 200 python tools/bench_gemm.py --large --only wgrad_qkv,wgrad_o,wgrad_fc1,wgrad_fc2 >> $G 2>&1; done
cat $G
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-220
timeout 600 python bench.py --no-also > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
