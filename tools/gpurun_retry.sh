#!/bin/bash
# Usage: tools/gpurun_retry.sh <timeout_s> <script> [gpus]  -- retries while the pod answers "busy" (nothing charged)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$OUT" | tail -60
  exit 0
done
echo "gave up"
