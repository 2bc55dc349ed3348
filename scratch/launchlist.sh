#!/bin/bash
# launch list of one small bench step (ncu, durations only)
TAG=$1; shift
CMD="python bench.py --steps 1 --warmup 1 --fields 32 --pool 16 --chunk 16 --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
ls -la gpurun_out/launches_$TAG.csv
