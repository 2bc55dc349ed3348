#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],{k:d['e2e'][k] for k in ('value','rle_fraction','host_threads','h2d_bytes_per_step')})"
done
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --rle-fraction 1.0 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],{k:d['e2e'][k] for k in ('value','rle_fraction','host_threads','h2d_bytes_per_step')})"
