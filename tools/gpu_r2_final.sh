#!/bin/bash
# Round-2 evidence call: the driver's own commands (GPU suite with -x, smoke, default bench incl. the CUDA-graph probe, the reference
# arm), the launch list of one Large step, the DRAM traffic of its GEMM launches, ncu --set full of the weight-gradient GEMM, and
# the attention micro-benchmark with / without the bias path.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "reference exit $?"; cut -c1-300 gpurun_out/bench_reference_$TAG.json
timeout 200 python tools/bench_attn.py --reps 10 --dropout 0.1 --only large > gpurun_out/microbench_attn_$TAG.txt 2>&1; cat gpurun_out/microbench_attn_$TAG.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_large_$TAG.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1; echo "launch list exit $?"
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k regex:gemm_bf16 --csv --log-file gpurun_out/gemm_traffic_large_$TAG.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1; echo "gemm traffic exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k 'regex:gemm_bf16_pair_kernel<.bool.1, .bool.1' -s 40 -c 1 \
    -o gpurun_out/prof_gemm_wgrad_$TAG python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_gemm_wgrad_$TAG.log 2>&1; echo "ncu wgrad exit $?"
timeout 200 python tools/debug_determinism.py > gpurun_out/determinism_$TAG.txt 2>&1; timeout 200 python tools/debug_determinism.py --large >> gpurun_out/determinism_$TAG.txt 2>&1; cut -c1-300 gpurun_out/determinism_$TAG.txt
ls -la gpurun_out | tail -12
