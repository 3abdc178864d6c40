#!/bin/bash
# usage: gpurun_retry.sh <gpus> <timeout> '<command>'  -- retries while the pod answers busy / transient (nothing is charged for those)
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3" 2>&1)
  if echo "$out" | grep -q "status=transient\|no box or slot\|retry in a few minutes"; then sleep 150; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
