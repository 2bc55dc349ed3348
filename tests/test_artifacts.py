"""CPU tests: .keras / HDF5 / pickle loading (improved_detection.py:23-41 contract),
the C-ABI surface, and the host-side helpers."""
import ctypes
import io
import json
import os
import re
import zipfile

import numpy as np
import pytest

from cell_image_analysis_b200 import _lib, artifacts, hdf5_min
from oracle import h5write

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hdf5_round_trip_nested_groups():
    rng = np.random.default_rng(0)
    tree = {"layers": {f"conv2d_{i}": {"vars": {"0": rng.standard_normal((3, 3, 4, 5)).astype(np.float32),
                                                "1": rng.standard_normal(5).astype(np.float32)}}
                       for i in range(23)},           # > 8 entries: several SNODs
            "optimizer": {"vars": {"0": np.array(7, np.int64), "1": rng.standard_normal(3)}},
            "vars": {}}
    ds = hdf5_min.H5File(h5write.write_h5(tree)).datasets()
    assert len(ds) == 23 * 2 + 2
    for i in range(23):
        assert np.array_equal(ds[f"layers/conv2d_{i}/vars/0"], tree["layers"][f"conv2d_{i}"]["vars"]["0"])
    assert ds["optimizer/vars/0"] == 7 and ds["optimizer/vars/1"].dtype == np.float64


def test_hdf5_rejects_garbage():
    with pytest.raises(hdf5_min.H5FormatError):
        hdf5_min.H5File(b"not an hdf5 file" * 10)


def test_load_keras_autoencoder_and_encoder(model_dir, oracle_weights):
    ae = artifacts.load_keras_cae(os.path.join(model_dir, "best_autoencoder.keras"))
    enc = artifacts.load_keras_cae(os.path.join(model_dir, "encoder.keras"))
    assert ae["n_conv"] == 7 and enc["n_conv"] == 3 and ae["bn_eps"] == 1e-3
    assert [k.shape for k in ae["kernels"]] == [(3, 3, 1, 32), (3, 3, 32, 64), (3, 3, 64, 32),
                                                (3, 3, 32, 32), (3, 3, 32, 64), (3, 3, 64, 32), (3, 3, 32, 1)]
    assert len(ae["bns"]) == 6 and len(enc["bns"]) == 3
    assert artifacts.same_encoder(ae, enc)
    assert all(k.dtype == np.float32 for k in ae["kernels"])


def test_layer_names_in_config_do_not_matter(tmp_path, oracle_weights):
    # Keras numbers layer *names* per session (conv2d_7 ...); weight paths use per-class counters
    p = tmp_path / "enc.keras"
    h5write.write_keras(str(p), oracle_weights, encoder_only=True, name_offset=14)
    enc = artifacts.load_keras_cae(str(p))
    assert np.array_equal(enc["kernels"][2], oracle_weights["kernels"][2])


def _rewrite_config(src, dst, edit):
    with zipfile.ZipFile(src) as z:
        members = {n: z.read(n) for n in z.namelist()}
    cfg = json.loads(members["config.json"])
    edit(cfg)
    members["config.json"] = json.dumps(cfg).encode()
    with zipfile.ZipFile(dst, "w") as z:
        for n, b in members.items():
            z.writestr(n, b)


def test_foreign_topology_is_refused(tmp_path, model_dir):
    src = os.path.join(model_dir, "best_autoencoder.keras")

    def relu6(cfg):
        cfg["config"]["layers"][1]["config"]["activation"] = "tanh"
    _rewrite_config(src, tmp_path / "a.keras", relu6)
    with pytest.raises(artifacts.ArtifactError):
        artifacts.load_keras_cae(str(tmp_path / "a.keras"))

    def drop_bn(cfg):
        del cfg["config"]["layers"][2]
    _rewrite_config(src, tmp_path / "b.keras", drop_bn)
    with pytest.raises(artifacts.ArtifactError):
        artifacts.load_keras_cae(str(tmp_path / "b.keras"))

    def filters(cfg):
        cfg["config"]["layers"][1]["config"]["filters"] = 16
    _rewrite_config(src, tmp_path / "c.keras", filters)
    with pytest.raises(artifacts.ArtifactError):
        artifacts.load_keras_cae(str(tmp_path / "c.keras"))


def test_sklearn_artifacts(artifacts):
    sp = artifacts["scaler_pca"]
    assert sp["F"] == 2048 and sp["C"] == 100 and sp["f32_flow"] and sp["center_is_f32"]
    pca = artifacts["sklearn"]["pca"]
    assert np.allclose(sp["offset"], pca.mean_ @ pca.components_.T)
    for key, nu in (("svm_conservative", 0.05), ("svm_moderate", 0.10)):
        m = artifacts[key]
        assert m["sv"].shape[1] == 100 and m["coef"].shape == (m["sv"].shape[0],)
        assert m["gamma"] > 0
        # libsvm one-class: sum(alpha) = nu * n_train
        n_train = artifacts["sklearn"]["detector_" + key.split("_")[1]].shape_fit_[0]
        assert abs(m["coef"].sum() - nu * n_train) < 1e-6 * n_train


def test_abi_exports_every_declared_symbol():
    """include/cia.h <-> libcia.so <-> the ctypes table agree; no compute is invoked."""
    hdr = open(os.path.join(ROOT, "include", "cia.h")).read()
    declared = set(re.findall(r"\b(cia_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cia_ctx"}
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cia.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.cia_version() >= 100
    p = _lib.default_params()
    assert (p.border_margin, p.area_min, p.area_max) == (10, 200, 8000)
    assert (p.ecc_max, p.mean_min, p.std_min, p.clip_limit) == (0.95, 0.5, 0.1, 0.02)
    assert ctypes.sizeof(_lib.Region) == 64 and ctypes.sizeof(_lib.Cell) == 56


def test_null_handle_is_an_error_not_a_crash():
    lib = _lib.load()
    assert lib.cia_destroy(None) < 0
    assert lib.cia_check_status(None, None) < 0
    assert lib.cia_launch_count(None) == 0
    assert lib.cia_last_error(None) == b"null handle"


def test_engine_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cell_image_analysis_b200.screening import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine()


def test_summary_helpers_match_reference_keys():
    from cell_image_analysis_b200.screening import detail_rows, summarize_sample
    s = {"reconstruction_mse": np.array([0.1, 0.3], np.float32),
         "reconstruction_mae": np.array([0.2, 0.4], np.float32),
         "conservative_predictions": np.array([1, -1]), "moderate_predictions": np.array([-1, -1]),
         "conservative_scores": np.array([-1.0, 2.0]), "moderate_scores": np.array([0.5, 3.0]),
         "conservative_anomaly_rate": 0.5, "moderate_anomaly_rate": 1.0}
    r = summarize_sample("wt", 3, s)
    assert list(r) == ["sample_name", "total_cells", "files_processed", "conservative_anomaly_rate",
                       "moderate_anomaly_rate", "mean_mse", "std_mse", "mean_mae", "std_mae"]
    rows = detail_rows("wt", s)
    assert list(rows[0]) == ["sample_name", "cell_id", "mse", "mae", "conservative_anomaly",
                             "moderate_anomaly", "conservative_score", "moderate_score"]
    assert rows[1]["conservative_anomaly"] and rows[1]["cell_id"] == 1
