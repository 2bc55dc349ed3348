"""Artifact loading: the six files ``load_trained_models`` reads
(improved_detection.py:23-41), unchanged, into plain NumPy.

* ``best_autoencoder.keras`` / ``encoder.keras``: Keras-3 zip (config.json +
  model.weights.h5).  Parsed with the in-tree minimal HDF5 reader; the layer graph in
  config.json is checked against the topology of CAE_improved_modeltrain.py:188-216
  and anything else is refused.
* ``scaler.pkl`` / ``pca.pkl`` / ``detector_*.pkl``: scikit-learn pickles
  (CAE_improved_modeltrain.py:437-444) -- unpickled with scikit-learn, then only their
  fitted arrays are used; all arithmetic happens in libcia.
"""
from __future__ import annotations

import json
import os
import pickle
import re
import zipfile

import numpy as np

from .hdf5_min import H5File, H5FormatError

FILTERS = [32, 64, 32, 32, 64, 32, 1]
MODEL_FILES = ("best_autoencoder.keras", "encoder.keras", "scaler.pkl", "pca.pkl",
               "detector_conservative.pkl", "detector_moderate.pkl")


class ArtifactError(ValueError):
    pass


def _snake(name: str) -> str:
    s = re.sub(r"\W+", "", name)
    s = re.sub(r"(.)([A-Z][a-z]+)", r"\1_\2", s)
    return re.sub(r"([a-z])([A-Z])", r"\1_\2", s).lower()


def _expect(cond, msg):
    if not cond:
        raise ArtifactError("unsupported CAE topology (expected CAE_improved_modeltrain.py:188-216): " + msg)


def _check_topology(layers, n_conv):
    """layers: list of (class_name, config) in model order, InputLayer first."""
    expect = ["InputLayer"]
    for i in range(n_conv):
        expect.append("Conv2D")
        if i < 6:
            expect.append("BatchNormalization")
            expect.append("MaxPooling2D" if i < 3 else "UpSampling2D")
    got = [c for c, _ in layers]
    _expect(got == expect, f"layer sequence {got}")
    eps = None
    ci = 0
    for cls, cfg in layers:
        if cls == "InputLayer":
            shape = cfg.get("batch_shape") or cfg.get("batch_input_shape")
            _expect(shape is None or list(shape)[1:] == [64, 64, 1], f"input shape {shape}")
        elif cls == "Conv2D":
            _expect(cfg.get("filters") == FILTERS[ci], f"conv {ci} filters {cfg.get('filters')}")
            _expect(list(cfg.get("kernel_size", [])) == [3, 3], "kernel_size")
            _expect(list(cfg.get("strides", [1, 1])) == [1, 1], "strides")
            _expect(cfg.get("padding") == "same", "padding")
            _expect(list(cfg.get("dilation_rate", [1, 1])) == [1, 1], "dilation")
            _expect(cfg.get("groups", 1) == 1 and cfg.get("use_bias", True), "groups/use_bias")
            _expect(cfg.get("data_format", "channels_last") == "channels_last", "data_format")
            act = cfg.get("activation")
            _expect(act == ("sigmoid" if ci == 6 else "relu"), f"conv {ci} activation {act}")
            ci += 1
        elif cls == "BatchNormalization":
            ax = cfg.get("axis", -1)
            ax = ax[0] if isinstance(ax, (list, tuple)) else ax
            _expect(ax in (-1, 3), f"BN axis {ax}")
            _expect(cfg.get("center", True) and cfg.get("scale", True), "BN center/scale")
            e = float(cfg.get("epsilon", 1e-3))
            _expect(eps is None or e == eps, "mixed BN epsilons")
            eps = e
        elif cls == "MaxPooling2D":
            _expect(list(cfg.get("pool_size", [2, 2])) == [2, 2], "pool_size")
            st = cfg.get("strides")
            _expect(st is None or list(st) == [2, 2], "pool strides")
        elif cls == "UpSampling2D":
            _expect(list(cfg.get("size", [2, 2])) == [2, 2], "upsampling size")
            _expect(cfg.get("interpolation", "nearest") == "nearest", "upsampling interpolation")
    return eps if eps is not None else 1e-3


def load_keras_cae(path: str) -> dict:
    """Read a ``.keras`` archive written by train:270-275 / 299-300.

    Returns dict(kernels=[...], biases=[...], bns=[(gamma, beta, mean, var), ...],
    bn_eps=float, n_conv=7 or 3), float32 arrays, kernels HWIO.
    """
    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        if "config.json" not in names or "model.weights.h5" not in names:
            raise ArtifactError(f"{path}: not a Keras v3 archive (members: {names})")
        cfg = json.loads(z.read("config.json"))
        h5 = z.read("model.weights.h5")
    _expect(cfg.get("class_name") in ("Functional", "Model"), f"model class {cfg.get('class_name')}")
    layers = [(l["class_name"], l.get("config", {})) for l in cfg["config"]["layers"]]
    n_conv = sum(1 for c, _ in layers if c == "Conv2D")
    _expect(n_conv in (3, 7), f"{n_conv} Conv2D layers")
    eps = _check_topology(layers, n_conv)

    try:
        ds = H5File(h5).datasets()
    except H5FormatError as e:
        raise ArtifactError(f"{path}: model.weights.h5: {e}") from e
    norm = {k.replace("\\", "/"): v for k, v in ds.items()}

    def var(layer, idx):
        suffix = f"/{layer}/vars/{idx}"
        hits = [k for k in norm if ("/" + k).endswith(suffix) and "optimizer" not in k.split("/")]
        if len(hits) != 1:
            raise ArtifactError(f"{path}: weight {suffix!r} found {len(hits)} times in model.weights.h5")
        return np.ascontiguousarray(norm[hits[0]], dtype=np.float32)

    # saving_lib names each layer's group snake_case(class) with a per-class counter in
    # model-layer order, independent of the user-visible layer names in config.json
    counters = {}
    out = dict(kernels=[], biases=[], bns=[], bn_eps=float(eps), n_conv=n_conv)
    ci = 0
    for cls, _cfg in layers:
        base = _snake(cls)
        k = counters.get(base, 0)
        counters[base] = k + 1
        lname = base if k == 0 else f"{base}_{k}"
        if cls == "Conv2D":
            kern, bias = var(lname, 0), var(lname, 1)
            cin = 1 if ci == 0 else FILTERS[ci - 1]
            if kern.shape != (3, 3, cin, FILTERS[ci]) or bias.shape != (FILTERS[ci],):
                raise ArtifactError(f"{path}: {lname} kernel {kern.shape} / bias {bias.shape}")
            out["kernels"].append(kern)
            out["biases"].append(bias)
            ci += 1
        elif cls == "BatchNormalization":
            bn = tuple(var(lname, j) for j in range(4))
            out["bns"].append(bn)
    return out


def same_encoder(ae: dict, enc: dict) -> bool:
    """D8: does encoder.keras hold the same weights as the autoencoder's encoder half?"""
    for i in range(3):
        if not np.array_equal(ae["kernels"][i], enc["kernels"][i]) or \
                not np.array_equal(ae["biases"][i], enc["biases"][i]):
            return False
        if any(not np.array_equal(a, b) for a, b in zip(ae["bns"][i], enc["bns"][i])):
            return False
    return ae["bn_eps"] == enc["bn_eps"]


def _unpickle(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def scaler_pca_arrays(scaler, pca) -> dict:
    """Fitted arrays of RobustScaler (det:134) and PCA (det:135)."""
    if getattr(pca, "whiten", False):
        raise ArtifactError("pca.whiten=True is not the reference configuration (train:413)")
    comp = np.asarray(pca.components_)
    mean = getattr(pca, "mean_", None)
    C, F = comp.shape
    center = getattr(scaler, "center_", None) if getattr(scaler, "with_centering", True) else None
    scale = getattr(scaler, "scale_", None) if getattr(scaler, "with_scaling", True) else None
    f32_flow = comp.dtype == np.float32
    if mean is None:
        offset = np.zeros(C, comp.dtype)
    else:
        # exactly sklearn/decomposition/_base.py: reshape(mean_, (1, -1)) @ components_.T
        offset = (np.reshape(np.asarray(mean), (1, -1)) @ comp.T)[0]
    return dict(F=F, C=C,
                center=None if center is None else np.ascontiguousarray(center, np.float64),
                center_is_f32=bool(center is not None and np.asarray(center).dtype == np.float32),
                scale=None if scale is None else np.ascontiguousarray(scale, np.float64),
                components=np.ascontiguousarray(comp, np.float64),
                offset=np.ascontiguousarray(offset, np.float64), f32_flow=bool(f32_flow))


def svm_arrays(det) -> dict:
    """Fitted arrays of a OneClassSVM (det:138-142)."""
    if det.kernel != "rbf":
        raise ArtifactError(f"detector kernel {det.kernel!r}: only 'rbf' (train:421-422) is supported")
    sv = det.support_vectors_
    if hasattr(sv, "toarray"):
        sv = sv.toarray()
    return dict(sv=np.ascontiguousarray(sv, np.float64),
                coef=np.ascontiguousarray(np.asarray(det.dual_coef_).reshape(-1), np.float64),
                gamma=float(det._gamma), rho=float(-np.asarray(det.intercept_).reshape(-1)[0]))


def load_model_dir(model_dir: str) -> dict:
    """Everything ``load_trained_models`` (det:23-41) reads except StarDist."""
    ae = load_keras_cae(os.path.join(model_dir, "best_autoencoder.keras"))
    if ae["n_conv"] != 7:
        raise ArtifactError("best_autoencoder.keras is not the 7-conv autoencoder")
    enc = load_keras_cae(os.path.join(model_dir, "encoder.keras"))
    if enc["n_conv"] != 3:
        raise ArtifactError("encoder.keras is not the 3-conv encoder")
    scaler = _unpickle(os.path.join(model_dir, "scaler.pkl"))
    pca = _unpickle(os.path.join(model_dir, "pca.pkl"))
    dc = _unpickle(os.path.join(model_dir, "detector_conservative.pkl"))
    dm = _unpickle(os.path.join(model_dir, "detector_moderate.pkl"))
    return dict(autoencoder=ae, encoder=enc, encoder_same=same_encoder(ae, enc),
                scaler_pca=scaler_pca_arrays(scaler, pca),
                svm_conservative=svm_arrays(dc), svm_moderate=svm_arrays(dm),
                sklearn=dict(scaler=scaler, pca=pca, detector_conservative=dc, detector_moderate=dm))
