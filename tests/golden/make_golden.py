"""Generates the committed fixtures under tests/golden/ with the CPU oracle.

Run here (CPU container):  python tests/golden/make_golden.py
Outputs
  model_dir/   best_autoencoder.keras, encoder.keras (same weights), scaler.pkl, pca.pkl,
               detector_conservative.pkl, detector_moderate.pkl -- the six files
               load_trained_models reads (improved_detection.py:28-41).  CAE weights are
               synthetic (seed 7); scaler / PCA / SVMs are really *fit* with
               scikit-learn on oracle encoder features exactly as
               CAE_improved_modeltrain.py:407-427 does.
  tiny_field.npz      a 256x256 field + every intermediate the oracle produces for it.
  config1_seed0.npz   oracle results for the 2048x2048 config-1 field (seed 0): region
                      table, kept labels, stats, scores -- the field itself is regenerated
                      from the seed (a checksum guards the generator).
The reference itself ships no golden vectors, and skimage / Keras cannot be
imported here, so these pin the *oracle* (regression) -- see oracle/__init__.py.
"""
import hashlib
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cae, clahe, extraction, h5write, scoring  # noqa: E402
from cell_image_analysis_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def synth_cae_weights(seed=7):
    rng = np.random.default_rng(seed)
    kernels, biases, bns = [], [], []
    for i, (cin, cout) in enumerate(cae.CONV_SHAPES):
        std = np.sqrt(2.0 / (9 * cin))
        kernels.append((rng.standard_normal((3, 3, cin, cout)) * std).astype(np.float32))
        biases.append((rng.standard_normal(cout) * 0.05).astype(np.float32))
        if i < 6:
            gamma = rng.uniform(0.5, 1.5, cout).astype(np.float32)
            gamma[rng.choice(cout, 3, replace=False)] *= -1          # a few negative gammas
            bns.append((gamma, (rng.standard_normal(cout) * 0.1).astype(np.float32),
                        (rng.standard_normal(cout) * 0.1).astype(np.float32),
                        rng.uniform(0.5, 1.5, cout).astype(np.float32)))
    return {"kernels": kernels, "biases": biases, "bns": bns}


def main():
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import RobustScaler
    from sklearn.svm import OneClassSVM

    md = os.path.join(HERE, "model_dir")
    os.makedirs(md, exist_ok=True)
    w = synth_cae_weights()
    h5write.write_keras(os.path.join(md, "best_autoencoder.keras"), w, encoder_only=False)
    h5write.write_keras(os.path.join(md, "encoder.keras"), w, encoder_only=True, name_offset=0)

    # training cells: config-1 style fields, seeds 100.. (train:113-157)
    cells = []
    for seed in range(100, 106):
        g, l = synth.make_field(seed)
        c, _s, _k, _t = extraction.extract_quality_cells_from_labels(g, l)
        cells.extend(c)
        print(f"seed {seed}: {len(c)} cells", flush=True)
    X = np.expand_dims(np.array(cells), -1).astype("float32")           # train:398
    _, enc = cae.forward(X, w, n_layers=3)
    feats = enc.reshape(len(enc), -1)                                    # train:402
    scaler = RobustScaler()                                              # train:408
    fs = scaler.fit_transform(feats)
    ncomp = min(100, fs.shape[1], fs.shape[0] - 1)                       # train:412
    pca = PCA(n_components=ncomp)
    fr = pca.fit_transform(fs)
    dets = {"conservative": OneClassSVM(kernel="rbf", gamma="scale", nu=0.05),
            "moderate": OneClassSVM(kernel="rbf", gamma="scale", nu=0.10)}   # train:420-423
    for name, d in dets.items():
        d.fit(fr)
        print(name, "nSV", d.support_vectors_.shape, "gamma", d._gamma,
              "rate", float(np.mean(d.predict(fr) == -1)), flush=True)
    for name, obj in (("scaler.pkl", scaler), ("pca.pkl", pca),
                      ("detector_conservative.pkl", dets["conservative"]),
                      ("detector_moderate.pkl", dets["moderate"])):
        with open(os.path.join(md, name), "wb") as f:
            pickle.dump(obj, f)                                          # train:437-444

    def score(cells_):
        return scoring.compute_anomaly_scores(cells_, w, w, scaler, pca, dets["conservative"],
                                              dets["moderate"])

    # ---- tiny field: every intermediate ----
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    g, l = synth.make_field(3, H, W, n, lo, hi, lu)
    c, stats, kept, tab = extraction.extract_quality_cells_from_labels(g, l)
    s = score(c)
    levels = [clahe.clahe_levels(g[k["bbox"][0]:k["bbox"][2], k["bbox"][1]:k["bbox"][3]]) for k in kept]
    np.savez_compressed(
        os.path.join(HERE, "tiny_field.npz"), green=g, labels=l, table=tab,
        kept_labels=np.array([k["label"] for k in kept]),
        kept_bbox=np.array([k["bbox"] for k in kept]),
        ecc=np.array([k["eccentricity"] for k in kept]),
        mean=np.array([k["mean_intensity"] for k in kept]),
        std=np.array([k["std_intensity"] for k in kept]),
        levels=np.concatenate([lv.ravel() for lv in levels]),
        crops=np.array(c), mse=s["reconstruction_mse"], mae=s["reconstruction_mae"],
        features=s["_features"], pca=s["_pca"],
        dec_cons=-s["conservative_scores"], dec_mod=-s["moderate_scores"],
        pred_cons=s["conservative_predictions"], pred_mod=s["moderate_predictions"])
    print("tiny:", len(c), "cells")

    # ---- config 1, seed 0 ----
    g, l = synth.make_field(0)
    c, stats, kept, tab = extraction.extract_quality_cells_from_labels(g, l)
    s = score(c)
    sha = hashlib.sha256(g.tobytes() + l.tobytes()).hexdigest()
    sub = np.arange(0, len(c), 16)
    np.savez_compressed(
        os.path.join(HERE, "config1_seed0.npz"), field_sha256=np.array(sha), table=tab,
        kept_labels=np.array([k["label"] for k in kept]),
        ecc=np.array([k["eccentricity"] for k in kept]),
        mean=np.array([k["mean_intensity"] for k in kept]),
        std=np.array([k["std_intensity"] for k in kept]),
        crop_subset_idx=sub, crop_subset=np.array(c)[sub].astype(np.float64),
        mse=s["reconstruction_mse"], mae=s["reconstruction_mae"],
        dec_cons=-s["conservative_scores"], dec_mod=-s["moderate_scores"],
        pred_cons=s["conservative_predictions"], pred_mod=s["moderate_predictions"])
    print("config1 seed0:", len(c), "cells; anomaly rates",
          s["conservative_anomaly_rate"], s["moderate_anomaly_rate"])


if __name__ == "__main__":
    main()
