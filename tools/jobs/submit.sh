#!/bin/bash
# usage: tools/jobs/submit.sh <timeout_s> <script> [gpus]   -- retries while the pod answers "busy"
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then out=$(gpurun --timeout $T -- "bash $S" 2>&1); else out=$(gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  rc=$?
  echo "$out" | tail -150
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 120; continue; fi
  if [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
