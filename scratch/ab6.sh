#!/bin/bash
run() { timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1',d['value'],d['stages_ms_per_step']['cae'])"; }
run base
CIA_L3_KERNEL=1 run l3tma
CIA_L3_TAPS_PER_FLUSH=3 run l3taps3
CIA_L3_TAPS_PER_FLUSH=3 timeout 300 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
