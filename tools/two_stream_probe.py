"""Experiment: alternate chunks between TWO engines (own workspaces) on two streams, so that the
CUDA-core stages of one chunk (scan, gates, crop, scoring epilogues) can overlap the tensor-core
autoencoder of the other.  Prints cells/s for one and two engines.  python tools/two_stream_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                       # noqa: E402
from cell_image_analysis_b200.artifacts import load_model_dir      # noqa: E402
from cell_image_analysis_b200.screening import Engine              # noqa: E402

Fc, P, NF = 64, 64, 1024
greens, labels = bench.make_pool(list(range(P)))
max_label = int(labels.max())
arts = load_model_dir(bench.MODEL_DIR)
engs = [Engine(device=0, precision=1) for _ in range(2)]
for e in engs:
    e.load_artifacts(arts)
dev = engs[0].tdev
g_dev = torch.from_numpy(greens.view(np.int16)).to(dev)
l_dev = torch.from_numpy(labels).to(dev)
cap = Fc * max_label
outs = [[e.alloc_outputs(cap, Fc) for _ in range(2)] for e in engs]
accs = [torch.zeros((4, 8), dtype=torch.float64, device=dev) for _ in engs]
streams = [torch.cuda.Stream(device=dev) for _ in engs]


def one_pass(n_eng):
    for i in range(NF // Fc):
        k = i % n_eng
        p0 = (i * Fc) % P
        with torch.cuda.stream(streams[k]):
            engs[k].screen_fields(g_dev[p0:p0 + Fc], l_dev[p0:p0 + Fc], max_label, outs[k][(i // n_eng) & 1], acc=accs[k])


for n_eng in (1, 2, 1, 2):
    for _ in range(2):
        one_pass(n_eng)
    torch.cuda.synchronize()
    for a in accs:
        a.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        one_pass(n_eng)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    cells = sum(float(a[:, 0].sum()) for a in accs)
    ms = e0.elapsed_time(e1)
    print(f"{n_eng} engine(s): {cells / (ms * 1e-3) / 1e6:.3f}M cells/s, {ms / 3:.1f} ms per {NF} fields")
