"""Experiment: two engine handles on two streams, each scoring every other chunk, so that the
ALU-bound crop kernel of one chunk can share the SMs with the tensor-bound CAE of the other."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cell_image_analysis_b200.artifacts import load_model_dir
from cell_image_analysis_b200.batch import BatchScreen
from cell_image_analysis_b200.screening import Engine

P, Fc, NF = 64, int(os.environ.get("FC", 32)), 512
greens, labels = bench.make_pool(list(range(P)))
max_label = int(labels.max())
dev = torch.device("cuda", 0)
g_dev = torch.from_numpy(greens.view(np.int16)).to(dev)
l_dev = torch.from_numpy(labels).to(dev)
arts = load_model_dir(bench.MODEL_DIR)
lanes = []
for k in range(2):
    eng = Engine(device=0, precision=1)
    eng.load_artifacts(arts)
    lanes.append(BatchScreen(eng, bench.H, bench.W, max_label, chunk_fields=Fc, n_strains=1))

def run(nl):
    for k in range(nl):
        with torch.cuda.stream(lanes[k].compute):
            lanes[k].acc.zero_()
    if nl == 1:
        lanes[0].run_device(g_dev, l_dev, NF)
    else:
        # lane k takes the pool halves: same total work
        half = P // 2
        for k in range(2):
            lanes[k].run_device(g_dev[k * half:(k + 1) * half], l_dev[k * half:(k + 1) * half], NF // 2)

for nl in (1, 2, 1, 2):
    for _ in range(2): run(nl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): run(nl)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    cells = sum(float(lanes[k].acc[:, 0].sum()) for k in range(nl))
    print(f"lanes={nl} chunk={Fc}: {dt*1e3:.1f} ms per {NF} fields, {cells/dt/1e6:.3f}M cells/s")
