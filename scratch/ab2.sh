#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r1k}
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python scratch/svm_time.py 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['e2e']['value'],d['stages_ms_per_step'],d['stage_rooflines'])"
CMD="python bench.py --steps 1 --warmup 1 --fields 64 --pool 32 --chunk 32 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'scaler_pca|svm_rbf|label_scan' -s 9 -c 4 \
    -o gpurun_out/prof_score_$TAG $CMD > gpurun_out/ncu_full_score_$TAG.log 2>&1
ls -la gpurun_out | tail -4
