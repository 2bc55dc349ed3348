"""Turns "parity unpinned" into a checked fact: golden vectors from the REAL reference.

The build image has neither scikit-image nor TensorFlow/Keras, so the committed
fixtures (``tiny_field.npz`` / ``config1_seed0.npz``) come from the restated oracle.  This script
is the other half: run it ONCE in any environment that has the reference's own dependencies

    pip install numpy scipy scikit-image tensorflow scikit-learn tifffile pandas matplotlib seaborn
    python tests/golden/make_reference_golden.py --reference /path/to/cell-image-analysis

and it imports the UNMODIFIED ``improved_detection.ProductionMutantScreening`` and calls its own
``extract_quality_cells`` (improved_detection.py:48-115) and ``compute_anomaly_scores``
(:117-153) on the same seeded synthetic fields the tests use.  Nothing of this repository's
oracle or CUDA path is on that path.  Only the two steps upstream of the hot path are replaced,
because they need a network download / are outside SURVEY 8's scope:

  * ``StarDist2D.from_pretrained`` (det:44) is never called (the instance is created without
    ``__init__``), and ``stardist_model.predict_instances`` (det:63) returns the synthetic label
    field; ``csbdeep.utils.normalize`` (det:62) is the identity.  ``stardist`` / ``csbdeep`` are
    stubbed in ``sys.modules`` when they are not installed, so the reference module imports.
  * ``tiff.imread`` (det:51) returns the synthetic image for the path ``"synthetic"``.

Outputs (commit them; tests/test_reference_golden.py picks them up automatically):

  tests/golden/reference_model_dir/   best_autoencoder.keras + encoder.keras written by the real
        Keras (``Model.save``) from the train:188-216 architecture carrying the same synthetic
        weights as tests/golden/model_dir, plus copies of the four scikit-learn pickles
  tests/golden/reference_tiny_field.npz, reference_config1_seed0.npz
        crops, stats (incl. solidity), mse / mae, scores and predictions as the reference returns them
  tests/golden/reference_stardist_model/, reference_stardist_tiny.npz   (only where stardist + csbdeep are installed)
        a StarDist2D model folder written by the real package and its normalize / predict / predict_instances
        outputs for the tiny field: pins oracle/stardist.py, oracle/stardist_post.c and csrc/segment.cu
  tests/golden/reference_versions.json   library versions of the generating environment
"""
import argparse
import importlib
import json
import os
import pickle
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _stub_missing(names):
    """``import stardist.models`` etc. must succeed for the reference module to import."""
    for name in names:
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            sub = ".".join(parts[:i])
            if sub not in sys.modules:
                sys.modules[sub] = types.ModuleType(sub)
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[sub])
    sd = sys.modules["stardist.models"]
    if not hasattr(sd, "StarDist2D"):
        sd.StarDist2D = type("StarDist2D", (), {"from_pretrained": staticmethod(lambda *_a, **_k: None)})
    cu = sys.modules["csbdeep.utils"]
    if not hasattr(cu, "normalize"):
        cu.normalize = lambda x, *a, **k: x


def build_keras_models(weights):
    """train:188-216 with the real Keras; weights in layer order (Conv kernel, bias; BN gamma,
    beta, moving_mean, moving_variance)."""
    from tensorflow.keras.layers import BatchNormalization, Conv2D, Input, MaxPooling2D, UpSampling2D
    from tensorflow.keras.models import Model
    inp = Input(shape=(64, 64, 1))
    x = inp
    convs, bns = [], []
    for filters in (32, 64, 32):                                  # train:190-200
        c = Conv2D(filters, (3, 3), activation="relu", padding="same"); x = c(x); convs.append(c)
        b = BatchNormalization(); x = b(x); bns.append(b)
        x = MaxPooling2D((2, 2), padding="same")(x)
    encoded = x
    for filters in (32, 64, 32):                                  # train:203-213
        c = Conv2D(filters, (3, 3), activation="relu", padding="same"); x = c(x); convs.append(c)
        b = BatchNormalization(); x = b(x); bns.append(b)
        x = UpSampling2D((2, 2))(x)
    c = Conv2D(1, (3, 3), activation="sigmoid", padding="same"); decoded = c(x); convs.append(c)   # train:215
    autoencoder, encoder = Model(inp, decoded), Model(inp, encoded)
    for layer, k, b in zip(convs, weights["kernels"], weights["biases"]):
        layer.set_weights([k, b])
    for layer, (gamma, beta, mean, var) in zip(bns, weights["bns"]):
        layer.set_weights([gamma, beta, mean, var])
    return autoencoder, encoder


def stardist_golden(out, synth):
    """Segmentation (det:44, 62-63) from the REAL csbdeep / stardist when they are installed: a StarDist2D model of
    the 2D_versatile_fluo configuration with freshly initialised weights (the pretrained ones need a download) is
    saved as a model folder, and ``normalize`` + ``predict`` + ``predict_instances`` of the seeded tiny field are
    dumped.  tests/test_reference_golden.py then holds the oracle (oracle/stardist.py, stardist_post.c) and the CUDA
    path to them -- including the one stated deviation of the restatement (Clipper's integer-snapped polygon
    intersection): a mismatch there shows up as differing labels."""
    try:
        from csbdeep.utils import normalize
        from stardist.models import Config2D, StarDist2D
    except Exception as e:                                   # noqa: BLE001
        print(f"stardist / csbdeep not importable ({type(e).__name__}): segmentation golden skipped")
        return
    if not hasattr(StarDist2D, "predict_instances") or not callable(normalize):
        print("stardist / csbdeep are stubs: segmentation golden skipped")
        return
    conf = Config2D(n_rays=32, grid=(2, 2), n_channel_in=1)   # the published 2-D configuration
    model = StarDist2D(conf, name="reference_stardist_model", basedir=out)
    folder = os.path.join(out, "reference_stardist_model")
    model.keras_model.save_weights(os.path.join(folder, "weights_best.h5"))
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    green, _ = synth.make_field(3, H, W, n, lo, hi, lu)
    x = normalize(green)                                                        # det:62
    prob, dist = model.predict(x)
    thr = float(np.quantile(prob, 0.7))                       # an untrained network: a threshold that yields objects
    with open(os.path.join(folder, "thresholds.json"), "w") as f:
        json.dump({"prob": thr, "nms": 0.3}, f)
    labels, details = model.predict_instances(x, prob_thresh=thr, nms_thresh=0.3)   # det:63
    np.savez_compressed(os.path.join(out, "reference_stardist_tiny.npz"), normalized=x, prob=prob, dist=dist,
                        labels=labels.astype(np.int32), points=np.asarray(details["points"]),
                        prob_kept=np.asarray(details["prob"]), coord=np.asarray(details["coord"]),
                        prob_thresh=np.array(thr), nms_thresh=np.array(0.3))
    print("stardist tiny:", int(labels.max()), "instances")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference", help="checkout of Kmatsuo57/cell-image-analysis")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()

    _stub_missing(["stardist.models", "csbdeep.utils"])
    sys.path.insert(0, args.reference)
    try:
        import matplotlib
        matplotlib.use("Agg")
        ref = importlib.import_module("improved_detection")      # the unmodified reference module
    except ImportError as e:
        sys.exit(f"make_reference_golden: {e}.\nThis recipe runs the reference's own code: it needs the reference's "
                 "imports (scikit-image, tensorflow / keras, tifffile, matplotlib, seaborn, pandas, scikit-learn; "
                 "stardist / csbdeep are optional) and a checkout at --reference. Nothing was written; the tests "
                 "keep reporting 'parity unpinned' for the stages that depend on these files.")

    from helpers import synth_cae_weights
    from cell_image_analysis_b200 import synth

    md = os.path.join(args.out, "reference_model_dir")
    os.makedirs(md, exist_ok=True)
    autoencoder, encoder = build_keras_models(synth_cae_weights(7))
    autoencoder.save(os.path.join(md, "best_autoencoder.keras"))  # train:262 (ModelCheckpoint) / :431
    encoder.save(os.path.join(md, "encoder.keras"))               # train:432
    for name in ("scaler.pkl", "pca.pkl", "detector_conservative.pkl", "detector_moderate.pkl"):
        shutil.copy(os.path.join(HERE, "model_dir", name), os.path.join(md, name))

    # the reference class without det:44's network download
    scr = ref.ProductionMutantScreening.__new__(ref.ProductionMutantScreening)
    scr.model_dir = md
    scr.autoencoder = ref.load_model(os.path.join(md, "best_autoencoder.keras"))     # det:28
    scr.encoder = ref.load_model(os.path.join(md, "encoder.keras"))                  # det:29
    for attr, name in (("scaler", "scaler.pkl"), ("pca", "pca.pkl"),
                       ("detector_conservative", "detector_conservative.pkl"),
                       ("detector_moderate", "detector_moderate.pkl")):
        with open(os.path.join(md, name), "rb") as f:
            setattr(scr, attr, pickle.load(f))                                        # det:32-41

    current = {}
    ref.tiff.imread = lambda path: current["green"]                                   # det:51
    ref.normalize = lambda x, *a, **k: x                                              # det:62
    scr.stardist_model = types.SimpleNamespace(
        predict_instances=lambda img: (current["labels"], {}))                        # det:63

    def run(green, labels):
        current["green"], current["labels"] = green, labels
        cells, stats = scr.extract_quality_cells("synthetic")                         # det:48-115
        scores = scr.compute_anomaly_scores(cells)                                    # det:117-153
        return cells, stats, scores

    def dump(path, green, labels, cells, stats, s, subset=None, extra=None):
        idx = np.arange(len(cells)) if subset is None else subset
        np.savez_compressed(
            path, n_cells=np.array(len(cells)),
            area=np.array([d["area"] for d in stats]),
            eccentricity=np.array([d["eccentricity"] for d in stats]),
            solidity=np.array([d["solidity"] for d in stats]),
            mean=np.array([d["mean_intensity"] for d in stats]),
            std=np.array([d["std_intensity"] for d in stats]),
            crop_idx=idx, crops=np.array(cells)[idx].astype(np.float64),
            mse=s["reconstruction_mse"], mae=s["reconstruction_mae"],
            dec_cons=-s["conservative_scores"], dec_mod=-s["moderate_scores"],
            pred_cons=s["conservative_predictions"], pred_mod=s["moderate_predictions"],
            **(extra or {}))

    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    g, l = synth.make_field(3, H, W, n, lo, hi, lu)
    cells, stats, s = run(g, l)
    dump(os.path.join(args.out, "reference_tiny_field.npz"), g, l, cells, stats, s)
    print("tiny:", len(cells), "cells")

    g, l = synth.make_field(0)
    cells, stats, s = run(g, l)
    dump(os.path.join(args.out, "reference_config1_seed0.npz"), g, l, cells, stats, s,
         subset=np.arange(0, len(cells), 16))
    print("config 1 seed 0:", len(cells), "cells; anomaly rates",
          s["conservative_anomaly_rate"], s["moderate_anomaly_rate"])

    stardist_golden(args.out, synth)

    versions = {}
    for mod in ("numpy", "scipy", "skimage", "tensorflow", "keras", "sklearn", "stardist", "csbdeep"):
        try:
            versions[mod] = importlib.import_module(mod).__version__
        except Exception as e:                      # noqa: BLE001
            versions[mod] = f"unavailable ({type(e).__name__})"
    with open(os.path.join(args.out, "reference_versions.json"), "w") as f:
        json.dump(versions, f, indent=1)
    print(versions)


if __name__ == "__main__":
    main()
