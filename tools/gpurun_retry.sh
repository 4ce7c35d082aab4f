#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...>   — retries while the pod is busy
G=$1; T=$2; shift 2
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up: pod busy"; exit 3
