"""GPU parity on BASELINE.json config 4: a synthetic detector with 20,000 support vectors x
256-dim PCA output, scored by the REAL libsvm through an overwritten fitted
OneClassSVM, against cia_svm_decision."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_SV, DIM, NU, N_TRAIN = 20000, 256, (0.05, 0.10), 200_000


def _synthetic_detector(rng, z_train, nu):
    from sklearn.svm import OneClassSVM
    det = OneClassSVM(kernel="rbf", gamma="scale", nu=0.5).fit(z_train[:64])
    sv = (z_train[rng.integers(0, len(z_train), N_SV)] +
          0.05 * rng.standard_normal((N_SV, DIM))).astype(np.float64)
    coef = rng.uniform(1e-6, 1.0, N_SV)
    coef *= nu * N_TRAIN / coef.sum()
    gamma = 1.0 / (DIM * z_train.var())
    det.support_vectors_ = np.ascontiguousarray(sv)
    det._dual_coef_ = det.dual_coef_ = np.ascontiguousarray(coef[None, :])
    det.support_ = np.arange(N_SV, dtype=np.int32)
    det._n_support = np.array([N_SV, N_SV], dtype=np.int32)      # libsvm one-class: nr_class = 2
    det.shape_fit_ = (N_TRAIN, DIM)
    det._gamma = gamma
    # put rho in the middle of the decision sums so that both signs occur
    d2 = ((z_train[:200, None, :] - sv[None, :2000, :]) ** 2).sum(-1)
    approx = (coef[:2000] * np.exp(-gamma * d2)).sum(1) * (N_SV / 2000)
    rho = float(np.median(approx))
    det._intercept_ = det.intercept_ = np.array([-rho])
    det.offset_ = np.array([rho])
    return det


@pytest.mark.parametrize("kernel", [0, 1])
def test_config4_svm_20k_sv_256d(model_dir, artifacts, golden_config1, field_config1, kernel):
    """kernel 0: fp64 DMMA anchor, 1e-7 of the largest decision.  kernel 1 (default): tcgen05 GEMM form,
    north_star's gate |d dec| <= 1e-4 absolute on decisions of magnitude ~500, identical signs."""
    from cell_image_analysis_b200.artifacts import svm_arrays
    from cell_image_analysis_b200.screening import Engine
    rng = np.random.default_rng(1234)
    # features of real cells through the fp32 anchor
    eng = Engine(device=0, precision=0)
    arts = dict(artifacts)
    q, _ = np.linalg.qr(rng.standard_normal((2048, DIM)))
    comp = np.ascontiguousarray(q.T.astype(np.float32))                     # [256, 2048] orthonormal
    base = artifacts["scaler_pca"]
    mean = rng.standard_normal(2048).astype(np.float32) * 0.1
    # no scaler here: with orthonormal components z = f @ W.T - offset keeps the N(0, 3^2) scale the
    # synthetic detectors were sized for (the scaler's own arithmetic is covered by the other tests)
    sp = dict(base, C=DIM, center=None, scale=None, components=comp.astype(np.float64),
              offset=(mean.reshape(1, -1) @ comp.T)[0].astype(np.float64), f32_flow=True)
    arts["scaler_pca"] = sp
    # z of a training-like cloud to size the synthetic detectors
    z_train = rng.standard_normal((4000, DIM)) * 3.0
    dets = [_synthetic_detector(rng, z_train, nu) for nu in NU]
    arts["svm_conservative"], arts["svm_moderate"] = svm_arrays(dets[0]), svm_arrays(dets[1])
    eng.load_artifacts(arts)
    eng.set_option("svm_kernel", kernel)
    n = 96
    feat = torch.from_numpy((rng.standard_normal((n, 2048)) * 3.0).astype(np.float32)).to(eng.tdev)
    dc, dm, pc, pm, z = eng.svm_decision(feat, n, want_pca=True, precision=1)   # the options decide the kernels
    eng.check_status()
    z = z[:n].cpu().numpy()
    for det, d_gpu, p_gpu in ((dets[0], dc, pc), (dets[1], dm, pm)):
        ref = det.decision_function(z)                       # real libsvm, 20k SVs x 256-d
        got = d_gpu[:n].cpu().numpy()
        assert np.abs(ref).max() > 1.0 and (ref > 0).any() and (ref < 0).any()
        print(f"config 4, svm_kernel {kernel}: max |d dec| {np.abs(got - ref).max():.3e} (|dec| max {np.abs(ref).max():.1f})")
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-7 * max(1.0, np.abs(ref).max()) if kernel == 0 else 1e-4)
        far = np.abs(ref) > 1e-6
        assert np.array_equal(p_gpu[:n].cpu().numpy().astype(np.intp)[far], det.predict(z)[far])
    eng.close()
