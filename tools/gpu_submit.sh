#!/bin/bash
# usage: tools/gpu_submit.sh <timeout_s> [--gpus N] -- '<command>'   -- retries while the pod answers "busy"
T=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun --timeout "$T" "$@" 2>&1); rc=$?
  echo "$out" | tail -6
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 150; continue; fi
  exit $rc
done
exit 3
