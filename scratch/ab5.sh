#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python scratch/svm_time.py 2>&1 | tail -2
CIA_PCA_VECTOR=1 timeout 300 python scratch/svm_time.py 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['e2e']['value'],d['stages_ms_per_step'])"
