"""Parity against golden vectors produced by the REAL reference (skimage + Keras + sklearn).

``tests/golden/make_reference_golden.py`` runs the unmodified
``improved_detection.ProductionMutantScreening.extract_quality_cells`` /
``compute_anomaly_scores`` where the reference's dependencies are installed and writes
``tests/golden/reference_*.npz`` + ``reference_model_dir/``.  The build image cannot run it
(scikit-image / TensorFlow are absent), so until someone commits those files these
tests SKIP with the reason "parity unpinned" -- and from then on they are the pinned gates:
  kept-cell count / area ........ exact
  eccentricity, solidity ......... 1e-9 / 1e-12
  crops .......................... |d| <= 1e-5 * max(|ref|, 1e-3)
  reconstruction MSE / MAE ....... <= 1e-3 relative
  SVM decision ................... identical signs, |d| <= 1e-4
The same gates are applied to the restated oracle (CPU) and to the CUDA path (gpu).
"""
import os
import runpy

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_MD = os.path.join(GOLDEN, "reference_model_dir")
CASES = {"tiny": "reference_tiny_field.npz", "config1": "reference_config1_seed0.npz"}
UNPINNED = ("parity unpinned: tests/golden/{} not committed -- generate it with "
            "tests/golden/make_reference_golden.py where scikit-image + TensorFlow are installed")


def _load(case):
    p = os.path.join(GOLDEN, CASES[case])
    if not os.path.exists(p):
        pytest.skip(UNPINNED.format(CASES[case]))
    return dict(np.load(p))


def _field(case):
    from cell_image_analysis_b200 import synth
    if case == "tiny":
        H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
        return synth.make_field(3, H, W, n, lo, hi, lu)
    return synth.make_field(0)


def _check(ref, cells, stats, scores):
    assert len(cells) == int(ref["n_cells"])
    assert np.array_equal(np.array([s["area"] for s in stats]), ref["area"])
    np.testing.assert_allclose([s["eccentricity"] for s in stats], ref["eccentricity"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([s["solidity"] for s in stats], ref["solidity"], rtol=1e-12)
    np.testing.assert_array_equal([s["mean_intensity"] for s in stats], ref["mean"])
    np.testing.assert_allclose([s["std_intensity"] for s in stats], ref["std"], rtol=1e-12)
    got = np.array(cells, dtype=np.float64)[ref["crop_idx"]]
    assert np.all(np.abs(got - ref["crops"]) <= 1e-5 * np.maximum(np.abs(ref["crops"]), 1e-3))
    np.testing.assert_allclose(scores["reconstruction_mse"], ref["mse"], rtol=1e-3)
    np.testing.assert_allclose(scores["reconstruction_mae"], ref["mae"], rtol=1e-3)
    for name, key in (("conservative", "cons"), ("moderate", "mod")):
        assert np.array_equal(scores[f"{name}_predictions"], ref[f"pred_{key}"])
        assert np.max(np.abs(-scores[f"{name}_scores"] - ref[f"dec_{key}"])) <= 1e-4


def test_generator_script_is_loadable():
    """The committed recipe parses and its dependency stubs make ``import stardist.models`` /
    ``csbdeep.utils`` succeed (what lets the unmodified reference module import without StarDist)."""
    mod = runpy.run_path(os.path.join(GOLDEN, "make_reference_golden.py"), run_name="not_main")
    import sys
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("stardist", "csbdeep")}
    try:
        mod["_stub_missing"](["stardist.models", "csbdeep.utils"])
        from csbdeep.utils import normalize
        from stardist.models import StarDist2D
        assert callable(normalize) and hasattr(StarDist2D, "from_pretrained")
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("stardist", "csbdeep")]:
            if k not in saved:
                del sys.modules[k]


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_against_reference_golden(case, oracle_weights, artifacts):
    """The restated oracle (numpy CLAHE / regionprops / Keras math) against the real libraries."""
    ref = _load(case)
    import pickle
    from oracle import extraction, scoring
    green, labels = _field(case)
    cells, stats, _kept, _tab = extraction.extract_quality_cells_from_labels(green, labels)
    objs = {}
    for name in ("scaler", "pca", "detector_conservative", "detector_moderate"):
        with open(os.path.join(GOLDEN, "model_dir", name + ".pkl"), "rb") as f:
            objs[name] = pickle.load(f)
    s = scoring.compute_anomaly_scores(cells, oracle_weights, oracle_weights, objs["scaler"], objs["pca"],
                                       objs["detector_conservative"], objs["detector_moderate"])
    _check(ref, cells, stats, s)


def test_keras_written_archive_loads():
    """hdf5_min / artifacts.load_model_dir on a .keras file written by the real Keras (h5py)."""
    if not os.path.exists(os.path.join(REF_MD, "best_autoencoder.keras")):
        pytest.skip(UNPINNED.format("reference_model_dir/best_autoencoder.keras"))
    from cell_image_analysis_b200.artifacts import load_model_dir
    from helpers import synth_cae_weights
    art = load_model_dir(REF_MD)
    w = synth_cae_weights(7)
    for a, b in zip(art["autoencoder"]["kernels"], w["kernels"]):
        assert np.array_equal(a, b)
    for a, b in zip(art["autoencoder"]["bns"], w["bns"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("case", list(CASES))
def test_cuda_path_against_reference_golden(case, precision, model_dir):
    """The product path (libcia.so through the drop-in class) against the real reference."""
    ref = _load(case)
    from cell_image_analysis_b200.screening import ProductionMutantScreening
    md = REF_MD if os.path.exists(os.path.join(REF_MD, "best_autoencoder.keras")) else model_dir
    s = ProductionMutantScreening(md, segmenter=lambda ch: None, device=0, precision=precision)
    green, labels = _field(case)
    cells, stats = s.extract_quality_cells_from_labels(green, labels)
    _check(ref, cells, stats, s.compute_anomaly_scores(cells))


# ---- segmentation (det:44, 62-63): golden from the real csbdeep / stardist, when the recipe could make it ----
SD_NPZ = os.path.join(GOLDEN, "reference_stardist_tiny.npz")
SD_DIR = os.path.join(GOLDEN, "reference_stardist_model")
needs_sd = pytest.mark.skipif(not (os.path.exists(SD_NPZ) and os.path.isdir(SD_DIR)),
                              reason="parity unpinned: tests/golden/reference_stardist_tiny.npz has not been generated "
                                     "(run tests/golden/make_reference_golden.py where stardist + csbdeep are installed)")


def _sd_case():
    import json
    from cell_image_analysis_b200 import synth
    from cell_image_analysis_b200.stardist import load_weights_h5
    ref = np.load(SD_NPZ)
    with open(os.path.join(SD_DIR, "config.json")) as f:
        cfg = json.load(f)
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    green, _ = synth.make_field(3, H, W, n, lo, hi, lu)
    return ref, cfg, load_weights_h5(os.path.join(SD_DIR, "weights_best.h5")), green


@needs_sd
def test_stardist_oracle_against_reference_golden():
    from oracle import stardist as sd
    ref, cfg, weights, green = _sd_case()
    assert np.array_equal(sd.normalize(green), ref["normalized"])
    prob, dist = sd.unet_forward(cfg, weights, ref["normalized"])
    assert np.abs(prob - ref["prob"]).max() <= 1e-4 and np.abs(dist - ref["dist"]).max() <= 1e-3 * np.abs(ref["dist"]).max()
    labels, det = sd.instances_from_prediction(ref["prob"], ref["dist"], int(cfg["grid"][0]), green.shape,
                                               float(ref["prob_thresh"]), float(ref["nms_thresh"]))
    assert np.array_equal(det["points"], ref["points"]), "kept set differs (Clipper's integer-snapped intersection?)"
    assert np.array_equal(labels, ref["labels"])


@needs_sd
@pytest.mark.gpu
def test_stardist_cuda_path_against_reference_golden():
    from cell_image_analysis_b200.stardist import StarDist2D
    ref, cfg, _weights, green = _sd_case()
    m = StarDist2D(None, name="reference_stardist_model", basedir=GOLDEN)
    assert np.array_equal(m.normalize_device(green).cpu().numpy(), ref["normalized"])
    prob, dist = m.predict(ref["normalized"])
    assert np.abs(prob.cpu().numpy() - ref["prob"]).max() <= 2e-2
    labels, n = m.instances_from_prediction(green.shape, torch_from(ref["prob"]), torch_from(ref["dist"]),
                                            float(ref["prob_thresh"]), float(ref["nms_thresh"]))
    assert np.array_equal(labels.cpu().numpy(), ref["labels"])


def torch_from(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a))

