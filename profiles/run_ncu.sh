#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list + one full capture of the top kernel.
# usage (on the GPU box, via gpurun):  bash profiles/run_ncu.sh <tag> <kernel-regex> [bench args...]
set -u
TAG=$1; KRE=$2; shift 2
CMD="python bench.py --steps 1 --warmup 1 --fields 32 --pool 16 --chunk 16 --no-cpu-baseline --no-svm-extras $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 3 \
    -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log | cut -c1-400
ls -la gpurun_out | tail -8
