"""A few scoring calls (scaler/PCA + both detectors) on the golden artifacts: the smallest program that
launches the scoring kernels, for ncu captures.  python tools/score_once.py [n_cells] [calls]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cell_image_analysis_b200.artifacts import load_model_dir   # noqa: E402
from cell_image_analysis_b200.screening import Engine           # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30400
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
size = sys.argv[3] if len(sys.argv) > 3 else "golden"      # golden | config4 (2 x 20000 SVs x 256-d, synthetic)
eng = Engine(device=0, precision=1)
arts = load_model_dir(os.path.join(ROOT, "tests", "golden", "model_dir"))
if size == "config4":
    r4 = np.random.default_rng(1)
    q, _ = np.linalg.qr(r4.standard_normal((2048, 256)))
    arts["scaler_pca"] = dict(arts["scaler_pca"], C=256, center=None, scale=None, components=np.ascontiguousarray(q.T),
                              offset=np.zeros(256), f32_flow=True)
    for k in ("svm_conservative", "svm_moderate"):
        arts[k] = dict(sv=r4.standard_normal((20000, 256)) * 3.0, coef=r4.uniform(0, 1, 20000), gamma=1.0 / (256 * 9.0), rho=1.0)
eng.load_artifacts(arts)
g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_field.npz"))
rng = np.random.default_rng(0)
f0 = g["features"].astype(np.float32)
feat = np.concatenate([f0 * (1 + 0.05 * rng.standard_normal(f0.shape).astype(np.float32))
                       for _ in range((n + len(f0) - 1) // len(f0))])[:n]
if size == "config4":
    feat = (rng.standard_normal((n, 2048)) * 3.0).astype(np.float32)
feat = torch.from_numpy(feat).to(eng.tdev)
for _ in range(calls):
    out = eng.svm_decision(feat, n)
torch.cuda.synchronize()
eng.check_status()
print("ok", float(out[0][:n].sum()))
