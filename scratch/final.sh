#!/bin/bash
# Final-state validation: GPU parity suite, default bench line, reference arm, launch list,
# full captures of the CAE kernels of one pass and of the crop kernel.
TAG=${TAG:-r1j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $?"
cut -c1-600 gpurun_out/bench_$TAG.json
CMD="python bench.py --steps 1 --warmup 1 --fields 32 --pool 16 --chunk 16 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv -s 7 -c 7 \
    -o gpurun_out/prof_cae_$TAG $CMD > gpurun_out/ncu_full_cae_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:crop_clahe -s 3 -c 3 \
    -o gpurun_out/prof_crop_$TAG $CMD > gpurun_out/ncu_full_crop_$TAG.log 2>&1
ls -la gpurun_out | tail -12
