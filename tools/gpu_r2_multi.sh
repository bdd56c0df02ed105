#!/bin/bash
# Multi-GPU call: `gpurun --gpus N -- bash tools/gpu_r2_multi.sh N tag`: NCCL gradient-equality test (N = 2), then weak-scaling
# lines for BASELINE configs[2] (WavLM-Large), configs[3] (UniSpeech-SAT pre-training step) and configs[4] (ragged).
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_ragged_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_multigpu_$TAG.log 2>&1; echo "multigpu pytest exit $?"; tail -4 gpurun_out/pytest_multigpu_$TAG.log
fi
timeout 600 bash -c "$(declare -f run); N=$N; run --no-cpu-baseline --no-also" > gpurun_out/bench_large_n${N}_$TAG.json 2> gpurun_out/bench_large_n${N}_$TAG.err; echo "large exit $?"; cut -c1-260 gpurun_out/bench_large_n${N}_$TAG.json; tail -2 gpurun_out/bench_large_n${N}_$TAG.err
timeout 600 bash -c "$(declare -f run); N=$N; run --sat --no-cpu-baseline --no-also --no-profile" > gpurun_out/bench_sat_n${N}_$TAG.json 2> gpurun_out/bench_sat_n${N}_$TAG.err; echo "sat exit $?"; cut -c1-260 gpurun_out/bench_sat_n${N}_$TAG.json; tail -2 gpurun_out/bench_sat_n${N}_$TAG.err
timeout 600 bash -c "$(declare -f run); N=$N; run --ragged --no-cpu-baseline --no-also --no-profile" > gpurun_out/bench_ragged_n${N}_$TAG.json 2> gpurun_out/bench_ragged_n${N}_$TAG.err; echo "ragged exit $?"; cut -c1-260 gpurun_out/bench_ragged_n${N}_$TAG.json; tail -2 gpurun_out/bench_ragged_n${N}_$TAG.err
