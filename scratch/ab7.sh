#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/b7.err > gpurun_out/b7.json; tail -3 gpurun_out/b7.err
python -c "
import json;d=json.load(open('gpurun_out/b7.json'));print(d['value'],d['e2e']['value']);print(d['roofline']);print(d['cae_layers'])"
python __graft_entry__.py smoke 2>&1 | tail -2
