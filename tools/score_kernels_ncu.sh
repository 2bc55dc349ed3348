#!/bin/bash
# per-kernel durations of the scoring stage (PCA + 2 detectors) at the three detector sizes, both kernel families
# usage (GPU box): bash tools/score_kernels_ncu.sh <tag> [n_cells]
TAG=$1; N=${2:-30400}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/score_launches_$TAG.csv \
    -k regex:'scaler_pca|svm_' python tools/svm_tc_probe.py $N > gpurun_out/score_ncu_$TAG.log 2>&1
python - <<P
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/score_launches_$TAG.csv", errors="ignore")) if len(r) > 14 and r[0].isdigit()]
seq=[(re.sub(r"^void |\(.*$|<unnamed>::|.*::", "", r[4]), r[8], float(r[14])/1e3) for r in rows]
# group consecutive identical (name, grid) launches
out=collections.OrderedDict()
for name, grid, us in seq:
    out.setdefault((name, grid), []).append(us)
for (name, grid), v in out.items():
    print(f"{name:36s} grid {grid:16s} n={len(v):3d}  min {min(v):9.1f} us  median {sorted(v)[len(v)//2]:9.1f} us")
P
