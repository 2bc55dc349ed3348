#!/bin/bash
# A/B of the layer-1 kernels: per-layer errors against the oracle, the mode gates, a short bench
mkdir -p gpurun_out
for k in 1 0; do
  echo "== CIA_L1_KERNEL=$k"; CIA_L1_KERNEL=$k timeout 300 python tests/debug_tc.py 1 2>&1 | tail -16
done
echo "== debias off"; CIA_L1_DEBIAS=0 timeout 300 python tests/debug_tc.py 1 2>&1 | grep -E "A1 h|feat|max\|d dec"
timeout 600 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for k in 1 0; do
  CIA_L1_KERNEL=$k timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('L1_KERNEL=$k',d['value'],d['e2e']['value'],d['stages_ms_per_step'])"
done
