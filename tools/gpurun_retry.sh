#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit code 3: nothing charged)
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD"; fi
  rc=$?
  if [ $rc -eq 2 ] && /usr/local/graft/bin/gpurun --status | grep -q '"in_flight": 1'; then sleep 60; continue; fi
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
