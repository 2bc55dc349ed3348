"""Shared builders for the GPU parity tests (test infrastructure only)."""
from __future__ import annotations

import numpy as np

# (Cin, Cout) of the seven Conv2D layers, CAE_improved_modeltrain.py:191-216
CONV_SHAPES = [(1, 32), (32, 64), (64, 32), (32, 32), (32, 64), (64, 32), (32, 1)]


def synth_cae_weights(seed: int = 7, scale: float = 1.0, n_negative_gamma: int = 3):
    """He-normal conv kernels (x ``scale``), BN gamma ~ U(0.5, 1.5) with a few NEGATIVE gammas,
    beta / mean ~ N(0, 0.1), var ~ U(0.5, 1.5) (SURVEY 8d "synthetic artifacts").  seed 7, scale 1
    reproduces tests/golden/make_golden.py's model."""
    rng = np.random.default_rng(seed)
    kernels, biases, bns = [], [], []
    for i, (cin, cout) in enumerate(CONV_SHAPES):
        std = np.sqrt(2.0 / (9 * cin))
        kernels.append((rng.standard_normal((3, 3, cin, cout)) * std * scale).astype(np.float32))
        biases.append((rng.standard_normal(cout) * 0.05).astype(np.float32))
        if i < 6:
            gamma = rng.uniform(0.5, 1.5, cout).astype(np.float32)
            gamma[rng.choice(cout, n_negative_gamma, replace=False)] *= -1
            bns.append((gamma, (rng.standard_normal(cout) * 0.1).astype(np.float32),
                        (rng.standard_normal(cout) * 0.1).astype(np.float32),
                        rng.uniform(0.5, 1.5, cout).astype(np.float32)))
    return {"kernels": kernels, "biases": biases, "bns": bns}


def fit_detectors(features: np.ndarray):
    """RobustScaler -> PCA(min(100, ...)) -> two OneClassSVMs, exactly as
    CAE_improved_modeltrain.py:407-427 fits them (real scikit-learn)."""
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import RobustScaler
    from sklearn.svm import OneClassSVM
    scaler = RobustScaler()                                              # train:408
    fs = scaler.fit_transform(features)
    pca = PCA(n_components=min(100, fs.shape[1], fs.shape[0] - 1))       # train:412-413
    fr = pca.fit_transform(fs)
    cons = OneClassSVM(kernel="rbf", gamma="scale", nu=0.05).fit(fr)     # train:420-423
    mod = OneClassSVM(kernel="rbf", gamma="scale", nu=0.10).fit(fr)
    return scaler, pca, cons, mod


def artifacts_from(weights: dict, scaler, pca, det_cons, det_mod, bn_eps: float = 1e-3):
    """The dict ``Engine.load_artifacts`` takes, from in-memory objects (what
    ``artifacts.load_model_dir`` builds from the six files of det:28-41)."""
    from cell_image_analysis_b200 import artifacts as A
    ae = dict(n_conv=7, kernels=weights["kernels"], biases=weights["biases"], bns=weights["bns"],
              bn_eps=bn_eps)
    return dict(autoencoder=ae, encoder_same=True, scaler_pca=A.scaler_pca_arrays(scaler, pca),
                svm_conservative=A.svm_arrays(det_cons), svm_moderate=A.svm_arrays(det_mod))
