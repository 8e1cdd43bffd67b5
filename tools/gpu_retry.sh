#!/bin/bash
# usage: gpu_retry.sh <timeout> '<command>'   -- retries gpurun while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
