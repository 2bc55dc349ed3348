#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; echo "bench rc $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_r1n.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'],d['stages_ms_per_step'],d['roofline']['frac'],d['roofline']['issued_frac'],d['cpu_baseline']['value'],d['gpu_launches'],d['clocks'])"
