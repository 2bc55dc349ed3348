"""GPU test of the whole drop-in: folders of 16-bit multi-channel TIFF fields ->
``screen_mutant_samples`` -> per-strain results, per-cell rows, CSV files and report, compared
with the oracle run of the same reference flow (improved_detection.py:155-244)."""
import os

import numpy as np
import pytest

from cell_image_analysis_b200 import synth, tiff_min
from oracle import extraction, scoring

pytestmark = pytest.mark.gpu


def _write_field(path, seed):
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    green, labels = synth.make_field(seed, H, W, n, lo, hi, lu)
    rgb = np.zeros((H, W, 3), np.uint16)
    rgb[..., 1] = green                      # det:56  analysis channel
    rgb[..., 2] = labels.astype(np.uint16)   # det:55  segmentation channel (carries the label ids here)
    tiff_min.write_tiff(path, rgb)
    return green, labels


def test_screen_mutant_samples_end_to_end(tmp_path, model_dir, artifacts, oracle_weights, capsys):
    from cell_image_analysis_b200.screening import ProductionMutantScreening
    folders, truth = {}, {}
    for strain, seeds in (("wt", (21, 22)), ("mutA", (23,))):
        d = tmp_path / strain
        d.mkdir()
        folders[strain] = str(d)
        truth[strain] = [_write_field(str(d / f"field_{s}.tif"), s) for s in seeds]
    (tmp_path / "wt" / "broken.tif").write_bytes(b"II*\0garbage")        # det:113-115 path
    empty = tmp_path / "empty"
    empty.mkdir()
    folders["empty"] = str(empty)                                        # det:168-170 path

    s = ProductionMutantScreening(model_dir, segmenter=lambda ch: ch.astype(np.int32))
    out = tmp_path / "out"
    results, rows = s.screen_mutant_samples(folders, str(out))
    printed = capsys.readouterr().out
    assert "Error processing" in printed and "broken.tif" in printed
    assert "No .tif files found" in printed
    assert set(results) == {"wt", "mutA"}

    sk = artifacts["sklearn"]
    for strain, fields in truth.items():
        cells = []
        for green, labels in fields:
            c, _st, _k, _t = extraction.extract_quality_cells_from_labels(green, labels)
            cells.extend(c)
        ref = scoring.compute_anomaly_scores(cells, oracle_weights, oracle_weights, sk["scaler"], sk["pca"],
                                             sk["detector_conservative"], sk["detector_moderate"])
        r = results[strain]
        assert r["total_cells"] == len(cells)
        assert r["files_processed"] == len(fields) + (1 if strain == "wt" else 0)   # glob counts the broken file
        assert r["conservative_anomaly_rate"] == ref["conservative_anomaly_rate"]
        assert r["moderate_anomaly_rate"] == ref["moderate_anomaly_rate"]
        assert abs(r["mean_mse"] / np.mean(ref["reconstruction_mse"]) - 1) < 1e-3
        mine = [x for x in rows if x["sample_name"] == strain]
        assert [x["cell_id"] for x in mine] == list(range(len(cells)))
        d = np.abs(np.array([x["conservative_score"] for x in mine]) - ref["conservative_scores"])
        assert d.max() <= 1e-4
    for name in ("screening_summary.csv", "detailed_cell_results.csv", "mutant_screening_report.txt"):
        assert os.path.exists(out / name)
    # the sharded twin (one rank here; two ranks in tests/test_gpu_sharded.py): same strains, counts and rows
    results2, rows2 = s.screen_mutant_samples_sharded(folders, str(tmp_path / "out2"), chunk_fields=2)
    assert list(results2) == list(results)
    for strain in results:
        a, b = results[strain], results2[strain]
        assert list(a) == list(b)                                              # key order of det:202-212
        for k in ("total_cells", "files_processed", "conservative_anomaly_rate", "moderate_anomaly_rate"):
            assert a[k] == b[k], (strain, k)
        for k in ("mean_mse", "std_mse", "mean_mae", "std_mae"):
            assert abs(float(a[k]) - b[k]) <= 1e-6 * abs(b[k]), (strain, k)    # float32 np.mean vs fp64 accumulator
    assert len(rows2) == len(rows)
    for x, y in zip(rows, rows2):
        assert x["sample_name"] == y["sample_name"] and x["cell_id"] == y["cell_id"]
        assert x["mse"] == y["mse"] and x["conservative_score"] == y["conservative_score"]
        assert x["moderate_anomaly"] == y["moderate_anomaly"]
    assert os.path.exists(tmp_path / "out2" / "screening_summary.csv")
    feats = s.encode_features(cells)
    assert feats.shape == (len(cells), 2048) and feats.dtype == np.float32
    assert np.abs(feats - ref["_features"]).max() <= 1e-5 * max(1.0, np.abs(ref["_features"]).max())
    s.engine.close()
