"""Oracle for ``compute_anomaly_scores`` (improved_detection.py:117-153).

The scaler / PCA / OneClassSVM calls are the REAL scikit-learn 1.9 + vendored libsvm
the reference calls (pinned); the CAE forward is oracle.cae (unpinned restatement).
Test infrastructure only.
"""
from __future__ import annotations

import numpy as np

from . import cae


def compute_anomaly_scores(cell_images, ae_w, enc_w, scaler, pca, det_cons, det_mod, exact=True):
    if len(cell_images) == 0:                                                # det:119
        return {}
    X = np.expand_dims(np.array(cell_images), axis=-1).astype("float32")     # det:122
    recon, enc_ae = cae.forward(X, ae_w, exact=exact)                                     # det:125
    mse, mae = cae.recon_errors(X, recon)                                    # det:126-127
    if enc_w is ae_w:
        enc = enc_ae
    else:
        _, enc = cae.forward(X, enc_w, n_layers=3, exact=exact)                           # det:130
    flat = enc.reshape(len(enc), -1)                                         # det:131 (HWC)
    z = pca.transform(scaler.transform(flat))                                # det:134-135
    cp, mp = det_cons.predict(z), det_mod.predict(z)                         # det:138-139
    cs, ms = det_cons.decision_function(z), det_mod.decision_function(z)     # det:141-142
    return {
        "reconstruction_mse": mse,
        "reconstruction_mae": mae,
        "conservative_predictions": cp,
        "moderate_predictions": mp,
        "conservative_scores": -cs,
        "moderate_scores": -ms,
        "conservative_anomaly_rate": np.sum(cp == -1) / len(cp),
        "moderate_anomaly_rate": np.sum(mp == -1) / len(mp),
        "_features": flat, "_pca": z,      # extra taps for stage-wise parity tests
    }
