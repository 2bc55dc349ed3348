#!/bin/bash
N=${N:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_n$N.json
python -c "
import json;d=json.load(open('gpurun_out/bench_n$N.json'));print(d['n_gpus'],d['value'],d['e2e'],d['stages_ms_per_step'])"
nproc
