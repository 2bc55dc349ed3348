#!/bin/bash
run() { timeout 400 python bench.py --no-cpu-baseline --steps 2 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$*',d['value'],d['e2e']['value'],d['stages_ms_per_step'])"; }
run --chunk 128 --pool 128
