#!/bin/bash
# usage: scratch/gpu_retry.sh <timeout> <command...>   -- retries while the pod answers busy (exit 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.txt 2>&1; rc=$?
  if [ $rc -ne 3 ]; then cat /tmp/gpurun_last.txt; exit $rc; fi
  sleep 45
done
echo "gave up"; exit 3
