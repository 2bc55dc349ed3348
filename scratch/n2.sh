#!/bin/bash
N=${N:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l);print(d['n_gpus'],d['value'],{k:d['e2e'][k] for k in ('value','label_transport','rle_fraction','host_threads','h2d_bytes_per_step')})"; }
run
run --label-transport raw
