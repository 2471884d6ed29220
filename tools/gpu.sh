#!/bin/bash
# usage: [GPUS=2] tools/gpu.sh <timeout_s> '<command>'  -- retries gpurun while the pod is busy (exit 3: nothing charged)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
