"""A/B of the two SVM kernels (cia_set_option "svm_kernel": 1 = tcgen05 GEMM form, 0 = fp64 DMMA):
max |d decision| against each other and against real libsvm, and the time per call, at the golden
detector sizes, a realistic size (5k SV x 100-d) and BASELINE config 4 (20k SV x 256-d).
Run on the GPU box:  python tools/svm_tc_probe.py [n_cells]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cell_image_analysis_b200.artifacts import load_model_dir, svm_arrays   # noqa: E402
from cell_image_analysis_b200.screening import Engine                       # noqa: E402

MODEL_DIR = os.path.join(ROOT, "tests", "golden", "model_dir")


def timed(eng, feat, n, reps=10, rounds=4):
    for _ in range(3):
        eng.svm_decision(feat, n)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(rounds):                  # best of a few rounds: the box's clocks ramp and other tenants come and go
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = eng.svm_decision(feat, n, want_pca=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best, out


def case(tag, arts, feat_np, dets=None, n_ref=64):
    n = len(feat_np)
    eng = Engine(device=0, precision=1)
    eng.load_artifacts(arts)
    feat = torch.from_numpy(feat_np).to(eng.tdev)
    res = {}
    for k in (0, 1):
        eng.set_option("svm_kernel", k)
        eng.set_option("pca_kernel", k)
        ms, (dc, dm, pc, pm, z) = timed(eng, feat, n)
        eng.check_status()
        res[k] = (ms, dc[:n].cpu().numpy(), dm[:n].cpu().numpy(), pc[:n].cpu().numpy(), pm[:n].cpu().numpy(), z[:n].cpu().numpy())
    d01 = max(np.abs(res[0][1] - res[1][1]).max(), np.abs(res[0][2] - res[1][2]).max())
    flips = int((res[0][3] != res[1][3]).sum() + (res[0][4] != res[1][4]).sum())
    dz = np.abs(res[0][5] - res[1][5]).max() / np.abs(res[0][5]).max()
    line = f"{tag}: n={n} max|dz|/max|z| {dz:.2e}, dmma {res[0][0]:.3f} ms, tc {res[1][0]:.3f} ms ({res[0][0] / res[1][0]:.1f}x), max|tc-dmma| {d01:.3e}, sign flips {flips}"
    if dets is not None:
        zr = res[1][5][:n_ref]
        for i, det in enumerate(dets):
            ref = det.decision_function(zr)
            line += f", det{i} |tc-libsvm| {np.abs(res[1][1 + i][:n_ref] - ref).max():.3e} |dmma-libsvm| {np.abs(res[0][1 + i][:n_ref] - ref).max():.3e} (|dec| max {np.abs(ref).max():.1f})"
    print(line, flush=True)
    eng.close()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 15130
    arts = load_model_dir(MODEL_DIR)
    rng = np.random.default_rng(1)
    import pickle
    dets = [pickle.load(open(os.path.join(MODEL_DIR, f"detector_{k}.pkl"), "rb")) for k in ("conservative", "moderate")]
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_field.npz"))
    f0 = g["features"].astype(np.float32)
    reps = (n + len(f0) - 1) // len(f0)
    feat = np.concatenate([f0 * (1 + 0.05 * rng.standard_normal(f0.shape).astype(np.float32)) for _ in range(reps)])[:n]
    feat[: len(f0)] = f0
    feat[len(f0)] *= 30.0            # an outlier cell
    feat[len(f0) + 1] = 0.0
    case("golden detectors (167 / 296 SV x 100-d)", arts, feat, dets, n_ref=min(n, 2000))

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_svm_stress as st
    for tag, nsv, dim in (("realistic 5k SV x 100-d", 5000, 100), ("config 4: 20k SV x 256-d", 20000, 256)):
        rng = np.random.default_rng(1234)
        q, _ = np.linalg.qr(rng.standard_normal((2048, dim)))
        comp = np.ascontiguousarray(q.T.astype(np.float32))
        mean = rng.standard_normal(2048).astype(np.float32) * 0.1
        a = dict(arts)
        a["scaler_pca"] = dict(arts["scaler_pca"], C=dim, center=None, scale=None, components=comp.astype(np.float64),
                               offset=(mean.reshape(1, -1) @ comp.T)[0].astype(np.float64), f32_flow=True)
        st.N_SV, st.DIM = nsv, dim
        z_train = rng.standard_normal((4000, dim)) * 3.0
        t0 = time.time()
        d = [st._synthetic_detector(rng, z_train, nu) for nu in st.NU]
        a["svm_conservative"], a["svm_moderate"] = svm_arrays(d[0]), svm_arrays(d[1])
        fx = (rng.standard_normal((n, 2048)) * 3.0).astype(np.float32)
        case(tag, a, fx, d, n_ref=96)


if __name__ == "__main__":
    main()
