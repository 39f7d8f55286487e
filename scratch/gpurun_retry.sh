#!/bin/bash
# usage: scratch/gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod has no free GPU slot (exit 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then tail -200 /tmp/gpurun_last.txt; exit $rc; fi
  sleep 120
done
echo "gave up: no GPU slot"; exit 3
