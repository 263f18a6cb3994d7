#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> '<command>'   -- retries while the pod reports no free slot (nothing is charged for those)
T=$1; shift
for i in $(seq 1 60); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -qE "status=transient|status=refused"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
