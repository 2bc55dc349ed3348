"""CPU tests of the pieces either side of the hot path: the
baseline-TIFF reader behind ``extract_quality_cells(path)`` and the result writers."""
import numpy as np
import pandas as pd
import pytest

from cell_image_analysis_b200 import reporting, tiff_min


@pytest.mark.parametrize("shape,dtype", [((37, 53), np.uint16), ((64, 40, 3), np.uint16), ((20, 20), np.uint8),
                                         ((130, 17, 4), np.uint16)])
def test_tiff_round_trip(tmp_path, shape, dtype):
    rng = np.random.default_rng(0)
    img = rng.integers(0, np.iinfo(dtype).max, shape).astype(dtype)
    p = tmp_path / "f.tif"
    tiff_min.write_tiff(str(p), img, rows_per_strip=16)
    back = tiff_min.read_tiff(str(p))
    assert back.dtype == dtype and np.array_equal(back, img)


def test_tiff_matches_opencv_decoder(tmp_path):
    cv2 = pytest.importorskip("cv2")
    img = np.random.default_rng(1).integers(0, 65535, (48, 64)).astype(np.uint16)
    p = tmp_path / "g.tif"
    tiff_min.write_tiff(str(p), img)
    assert np.array_equal(cv2.imread(str(p), cv2.IMREAD_UNCHANGED), img)     # an independent reader agrees
    cv2.imwrite(str(tmp_path / "cv.tif"), img, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    assert np.array_equal(tiff_min.read_tiff(str(tmp_path / "cv.tif")), img)  # and we read its files


def test_tiff_rejects_what_it_cannot_read(tmp_path):
    p = tmp_path / "bad.tif"
    p.write_bytes(b"II*\0" + b"\0" * 3)
    with pytest.raises(tiff_min.TiffError):
        tiff_min.read_tiff(str(p))
    p.write_bytes(b"not a tiff at all")
    with pytest.raises(tiff_min.TiffError):
        tiff_min.read_tiff(str(p))


def _results():
    res = {
        "wt": dict(sample_name="wt", total_cells=950, files_processed=2, conservative_anomaly_rate=0.05,
                   moderate_anomaly_rate=0.11, mean_mse=0.0123, std_mse=0.004, mean_mae=0.08, std_mae=0.01),
        "mutA": dict(sample_name="mutA", total_cells=400, files_processed=1, conservative_anomaly_rate=0.31,
                     moderate_anomaly_rate=0.42, mean_mse=0.0456, std_mse=0.01, mean_mae=0.15, std_mae=0.03),
    }
    rows = [dict(sample_name="wt", cell_id=0, mse=0.01, mae=0.08, conservative_anomaly=False, moderate_anomaly=True,
                 conservative_score=-1.5, moderate_score=0.2)]
    return res, rows


def test_writers_produce_reference_files(tmp_path):
    res, rows = _results()
    summary = reporting.save_and_report(res, rows, str(tmp_path))
    s = pd.read_csv(tmp_path / reporting.SUMMARY_CSV, index_col=0)
    assert list(s.index) == ["wt", "mutA"]
    assert list(s.columns) == ["sample_name", "total_cells", "files_processed", "conservative_anomaly_rate",
                               "moderate_anomaly_rate", "mean_mse", "std_mse", "mean_mae", "std_mae"]
    d = pd.read_csv(tmp_path / reporting.DETAILED_CSV)
    assert list(d.columns) == ["sample_name", "cell_id", "mse", "mae", "conservative_anomaly", "moderate_anomaly",
                               "conservative_score", "moderate_score"]
    txt = (tmp_path / reporting.REPORT_TXT).read_text()
    assert txt.startswith("=== MUTANT SCREENING REPORT (IMPROVED MODEL) ===")
    assert "HIGH ANOMALY CANDIDATES (Conservative >15%):\n- mutA: 31.0%" in txt
    assert "HIGH ANOMALY CANDIDATES (Moderate >25%):\n- mutA: 42.0%" in txt
    assert "NORMAL-LEVEL SAMPLES (Conservative ≤10%):\n- wt: 5.0%" in txt
    assert f"{'wt':<20} {950:<8} {5.0:>8.1f}% {11.0:>10.1f}% {0.0123:>10.6f}" in txt
    assert len(summary) == 2


def test_report_omits_empty_groups():
    res, _ = _results()
    only_wt = pd.DataFrame.from_dict({"wt": res["wt"]}, orient="index")
    lines = reporting.screening_report_lines(only_wt)
    assert not any(l.startswith("HIGH ANOMALY") for l in lines)
    assert any(l.startswith("NORMAL-LEVEL") for l in lines)
