#!/bin/bash
# One gpurun call: new-feature tests first, the whole GPU suite, smoke, bench lines, then (only if the new tests passed)
# micro-benchmarks and ncu captures of the new kernels.  Outputs under gpurun_out/.  Usage: bash tools/gpu_round.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 300 python -m pytest tests/test_dropout_gpu.py tests/test_optim_gpu.py tests/test_pretrain_gpu.py tests/test_ragged_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_new_$TAG.log 2>&1
NEW=$?
echo "new tests exit $NEW"; tail -5 gpurun_out/pytest_new_$TAG.log
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_dropout_gpu.py --deselect tests/test_optim_gpu.py --deselect tests/test_pretrain_gpu.py --deselect tests/test_ragged_gpu.py > gpurun_out/pytest_all_$TAG.log 2>&1
echo "suite exit $?"; tail -3 gpurun_out/pytest_all_$TAG.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
if [ $NEW -eq 0 ]; then
  timeout 420 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench large exit $?"
else
  timeout 420 python bench.py --steps 10 --warmup 3 --no-also > gpurun_out/bench_large_$TAG.json 2> gpurun_out/bench_large_$TAG.err; echo "bench large (no also) exit $?"
fi
cut -c1-300 gpurun_out/bench_large_$TAG.json; tail -3 gpurun_out/bench_large_$TAG.err
timeout 200 python bench.py --model base --steps 10 --warmup 3 --no-also > gpurun_out/bench_base_$TAG.json 2> gpurun_out/bench_base_$TAG.err; echo "bench base exit $?"
cut -c1-300 gpurun_out/bench_base_$TAG.json
if [ $NEW -eq 0 ]; then
  timeout 120 python tools/bench_attn.py --reps 10 --dropout 0.1 > gpurun_out/bench_attn_$TAG.txt 2>&1; cat gpurun_out/bench_attn_$TAG.txt
  timeout 120 python tools/bench_rowops.py --reps 20 > gpurun_out/bench_rowops_$TAG.txt 2>&1; cat gpurun_out/bench_rowops_$TAG.txt
  cap() {  # name regex skip command...
    local name=$1 rx=$2 skip=$3; shift 3
    timeout 240 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_${name}_$TAG "$@" > gpurun_out/ncu_${name}_$TAG.log 2>&1
    echo "ncu $name exit $?"
  }
  cap attn_fwd_drop attn_fwd_kernel 4 python tools/bench_attn.py --reps 1 --only base --dropout 0.1
  cap attn_bwd_drop attn_bwd_fused_kernel 4 python tools/bench_attn.py --reps 1 --only base --dropout 0.1
  cap dropout_rows dropout_rows_kernel 5 python tools/bench_rowops.py --reps 1 --only base
  cap adam_step adam_step_kernel 2 python tools/bench_rowops.py --reps 1 --only base
fi
# DRAM traffic of the GEMM launches of one WavLM-Large step (roofline.traffic of the headline line)
timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:gemm_bf16 --csv --log-file gpurun_out/gemm_traffic_large_$TAG.csv \
    python bench.py --ncu-step --warmup 3 > /dev/null 2>&1; echo "gemm traffic exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_large_$TAG.csv python bench.py --ncu-step --warmup 3 > /dev/null 2>&1; echo "launch list exit $?"
ls -la gpurun_out | tail -30
