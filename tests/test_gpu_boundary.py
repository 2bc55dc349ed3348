"""GPU tests of the boundary pieces around the hot path: the ``solidity`` key of the stats dict
(det:106), 8-bit fields, negative labels, the training-side twin (CAE_improved_modeltrain.py:39-111,
328-339, 394-446)."""
import os

import numpy as np
import pytest

from oracle import extraction as oext
from oracle import regions as oreg
from oracle import scoring as oscoring

pytestmark = pytest.mark.gpu


def test_solidity_matches_qhull_oracle(screener, field_config1):
    green, labels = field_config1
    cells, stats, rec = screener.extract_quality_cells_from_labels(green, labels, return_regions=True)
    assert list(stats[0]) == ["area", "eccentricity", "solidity", "mean_intensity", "std_intensity"]   # det:103-109
    worst = 0.0
    for s, r in zip(stats[::3], rec[::3]):
        mask = labels[r["minr"]:r["maxr"], r["minc"]:r["maxc"]] == r["label"]
        ref = oreg.solidity(mask)
        worst = max(worst, abs(s["solidity"] - ref))
        assert s["solidity"] == ref, (r["label"], s["solidity"], ref)        # integer counts: exact
    assert 0.5 < np.mean([s["solidity"] for s in stats]) <= 1.0


def test_solidity_of_awkward_shapes(screener):
    """Concave, disconnected and one-pixel-wide regions (a label need not be connected)."""
    rng = np.random.default_rng(9)
    H = W = 160
    labels = np.zeros((H, W), np.int32)
    labels[20:50, 20:24] = 1; labels[46:50, 20:60] = 1                  # an L
    labels[70:100, 30:60][rng.random((30, 60 - 30)) > 0.55] = 2          # salt: many holes, ragged rows
    labels[110:140, 100:130][np.add.outer(np.arange(30), np.arange(30)) % 7 == 0] = 3   # diagonal stripes
    labels[20:45, 100:101] = 4; labels[20:21, 100:140] = 4               # thin hook
    green = (rng.integers(200, 4000, (H, W))).astype(np.uint16)
    eng = screener.engine
    old = (eng.params.area_min, eng.params.ecc_max)
    eng.params.area_min, eng.params.ecc_max = 20, 1.0
    try:
        cells, stats, rec = screener.extract_quality_cells_from_labels(green, labels, return_regions=True)
    finally:
        eng.params.area_min, eng.params.ecc_max = old
    assert rec["label"].tolist() == [1, 2, 3, 4]
    for s, r in zip(stats, rec):
        mask = labels[r["minr"]:r["maxr"], r["minc"]:r["maxc"]] == r["label"]
        assert s["solidity"] == oreg.solidity(mask), r["label"]


def test_uint8_field_and_negative_labels(screener):
    """8-bit fields take img_as_float's 1/255 inside equalize_adapthist; negative labels are
    background (scipy.ndimage.find_objects ignores them)."""
    from cell_image_analysis_b200 import synth
    from cell_image_analysis_b200.screening import UnsupportedImageError
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    green16, labels = synth.make_field(5, H, W, n, lo, hi, lu)
    green8 = np.minimum(green16 // 16, 255).astype(np.uint8)
    lab = labels.copy()
    lab[0:5, 0:7] = -3
    cells, stats = screener.extract_quality_cells_from_labels(green8, lab)
    ref_cells, ref_stats, kept, tab = oext.extract_quality_cells_from_labels(green8, lab)
    assert len(cells) == len(ref_cells) > 0
    for s, k in zip(stats, ref_stats):
        assert s["area"] == k["area"] and s["mean_intensity"] == k["mean_intensity"]
    d = np.abs(np.array(cells) - np.array(ref_cells))
    assert (d <= 1e-5 * np.maximum(np.abs(np.array(ref_cells)), 1e-3)).all(), d.max()
    with pytest.raises(UnsupportedImageError):
        screener.extract_quality_cells_from_labels(green16.astype(np.float32), labels)


class _FakeStarDist:
    """Stands in for StarDist2D: the segmentation channel of the fixture carries the label ids."""

    def __init__(self, labels_by_shape):
        self.labels = labels_by_shape

    def predict_instances(self, normalized):
        return self.labels[normalized.shape], {}


def test_training_twin(tmp_path, model_dir, artifacts, oracle_weights):
    from cell_image_analysis_b200 import synth, tiff_min
    from cell_image_analysis_b200.training import ImprovedAnomalyDetectionTraining, normalize_percentile
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    folder = tmp_path / "train"
    folder.mkdir()
    fields = {}
    for seed in (31, 32, 33):
        green, labels = synth.make_field(seed, H, W, n, lo, hi, lu)
        tiff_min.write_tiff(str(folder / f"f{seed}.tif"), green, compression=8)
        fields[f"f{seed}.tif"] = (green, labels)
    out = tmp_path / "out"

    class Seg:
        def __init__(self):
            self.i = 0
            self.order = sorted(fields)

        def predict_instances(self, normalized):
            assert normalized.dtype == np.float32                       # csbdeep normalize's output dtype
            lab = fields[self.order[self.i]][1]
            self.i += 1
            return lab, {}

    t = ImprovedAnomalyDetectionTraining(str(out))
    cells, stats_df = t.create_training_dataset(str(folder), stardist_model=Seg())
    ref_cells, ref_stats = [], []
    for name in sorted(fields):
        c, st, _k, _t = oext.extract_quality_cells_from_labels(*fields[name])
        ref_cells.extend(c)
        ref_stats.extend(dict(s, file=name) for s in st)
    assert cells.shape == (len(ref_cells), 64, 64)
    assert list(stats_df.columns) == ["area", "eccentricity", "solidity", "mean_intensity", "std_intensity", "file"]
    assert stats_df["file"].tolist() == [s["file"] for s in ref_stats]
    assert stats_df["solidity"].tolist() == [s["solidity"] for s in ref_stats]
    assert np.abs(cells - np.array(ref_cells)).max() <= 1e-5
    for f in ("cell_statistics.csv", "file_summary.csv", "data_quality_report.txt"):
        assert os.path.exists(out / f)
    assert "Solidity:" in (out / "data_quality_report.txt").read_text()
    # train:328-339 and train:394-446 on the committed .keras files
    mse, mae = t.evaluate_reconstruction_quality(os.path.join(model_dir, "best_autoencoder.keras"), cells)
    sk = artifacts["sklearn"]
    ref = oscoring.compute_anomaly_scores(list(cells), oracle_weights, oracle_weights, sk["scaler"], sk["pca"],
                                          sk["detector_conservative"], sk["detector_moderate"])
    np.testing.assert_allclose(mse, ref["reconstruction_mse"], rtol=1e-3)
    np.testing.assert_allclose(mae, ref["reconstruction_mae"], rtol=1e-3)
    detectors, scaler, pca = t.create_anomaly_detector(os.path.join(model_dir, "encoder.keras"), cells)
    assert set(detectors) == {"Conservative", "Moderate"} and pca.n_components_ == min(100, 2048, len(cells) - 1)
    for f in ("scaler.pkl", "pca.pkl", "detector_conservative.pkl", "detector_moderate.pkl"):
        assert os.path.exists(out / f)
    feats = t.encode_features(os.path.join(model_dir, "encoder.keras"), cells)
    assert np.abs(feats - ref["_features"]).max() <= 1e-5 * max(1.0, np.abs(ref["_features"]).max())
    x = np.arange(1000, dtype=np.uint16).reshape(10, 100)
    nrm = normalize_percentile(x)
    assert nrm.dtype == np.float32 and abs(float(np.percentile(nrm, 3))) < 1e-6 and abs(float(np.percentile(nrm, 99.8)) - 1) < 1e-6
    t.engine.close()
