#!/bin/bash
# GPU parity suite, a short bench, a launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['e2e']['value'],d['stages_ms_per_step'])"
bash scratch/launchlist.sh ${TAG:-x} > /dev/null 2>&1
