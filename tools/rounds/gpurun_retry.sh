#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> <script> [gpus]   -- retries while the pod answers "busy" (exit code 3)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then gpurun --timeout $T -- "bash $S" > /tmp/gpurun_last.txt 2>&1; else gpurun --gpus $G --timeout $T -- "bash $S" > /tmp/gpurun_last.txt 2>&1; fi
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt; then sleep 60; continue; fi
  break
done
grep -v "^+" /tmp/gpurun_last.txt | tail -${TAIL:-25}
