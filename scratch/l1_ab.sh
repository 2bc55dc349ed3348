#!/bin/bash
# per-layer errors against the oracle, the GPU parity suite, a short bench, a launch list
mkdir -p gpurun_out
timeout 300 python tests/debug_tc.py 1 2>&1 | grep -E "A1 h\+l|A2 h\+l|feat|mse rel"
echo "== debias off"; CIA_L2_DEBIAS=0 CIA_L3_DEBIAS=0 timeout 300 python tests/debug_tc.py 1 2>&1 | grep -E "A2 h\+l|feat"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['e2e']['value'],d['stages_ms_per_step'])"
bash scratch/launchlist.sh ${TAG:-x} > /dev/null 2>&1
