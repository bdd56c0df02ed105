#!/bin/bash
# Multi-GPU call: `gpurun --gpus N -- bash tools/gpu_r2_multi.sh N tag`: NCCL gradient-equality test (N = 2), then weak-scaling
# lines for BASELINE configs[2] (WavLM-Large), configs[3] (UniSpeech-SAT pre-training step) and configs[4] (ragged), each with the
# device time per phase (--phases); at N = 2 also the N = 1 lines of the same box for the pre-training step.
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
one() {  # name, extra bench flags...
  local name=$1; shift
  timeout 600 bash -c "$(declare -f run); N=$N; run --no-cpu-baseline --no-also --no-profile --phases $*" > gpurun_out/bench_${name}_n${N}_$TAG.out 2> gpurun_out/bench_${name}_n${N}_$TAG.err
  echo "$name exit $?"; grep '^{' gpurun_out/bench_${name}_n${N}_$TAG.out | tail -1 > gpurun_out/bench_${name}_n${N}_$TAG.json
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_${name}_n${N}_$TAG.json"))
print("${name} N=${N}: value %.1f  ms/step %.2f  e2e %.1f  phases %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("phases_ms")))
PY
  tail -2 gpurun_out/bench_${name}_n${N}_$TAG.err | cut -c1-300
}
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_graph_gpu.py -q -p no:cacheprovider > gpurun_out/pytest_multigpu_$TAG.log 2>&1; echo "multigpu pytest exit $?"; tail -5 gpurun_out/pytest_multigpu_$TAG.log | cut -c1-250
  for extra in "--sat" ""; do
    CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-cpu-baseline --no-also --no-profile --phases $extra > gpurun_out/bench_n1_same_box_${extra#--}_$TAG.out 2> gpurun_out/bench_n1_same_box_${extra#--}_$TAG.err
    grep '^{' gpurun_out/bench_n1_same_box_${extra#--}_$TAG.out | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 same box [$extra]: value %.1f ms/step %.2f phases %s' % (d['value'], d['ms_per_step'], d.get('phases_ms')))"
  done
fi
if [ "$N" = "4" ]; then timeout 300 python -m pytest tests/test_w2v_gpu.py -q -p no:cacheprovider 2>&1 | tail -3; fi
one large
one sat --sat
one ragged --ragged
if [ "$N" = "8" ]; then B200S_NCCL_CTAS=4 one large_ctas4; fi
