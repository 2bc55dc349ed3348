#!/bin/bash
# calibration of the truncation de-bias factors and taps per flush (per-layer signed errors vs the oracle)
run() { echo "== $*"; env "$@" timeout 300 python tests/debug_tc.py 1 2>&1 | grep -E "A2 h\+l|feat"; }
run CIA_L2_TAPS_PER_FLUSH=9 CIA_L2_DEBIAS=0
run CIA_L2_TAPS_PER_FLUSH=9 CIA_L2_DEBIAS=6
for g in 3 9; do
CIA_L2_TAPS_PER_FLUSH=$g timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('G=$g',d['value'],d['e2e']['value'],d['stages_ms_per_step'])"
done
