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


def _field(shape=(300, 411), seed=0):
    rng = np.random.default_rng(seed)
    ramp = np.linspace(0, 30000, shape[1])[None, :]
    return (rng.integers(0, 4000, shape) + ramp).astype(np.uint16)


@pytest.mark.parametrize("compression", [5, 8, 32946, 32773])
def test_tiff_reads_compressed_files_written_by_libtiff(tmp_path, compression):
    """LZW (with libtiff's horizontal predictor), Deflate (both tag values) and PackBits strips
    as OpenCV / libtiff writes them -- an independent encoder."""
    cv2 = pytest.importorskip("cv2")
    img = _field()
    p = str(tmp_path / "c.tif")
    assert cv2.imwrite(p, img, [cv2.IMWRITE_TIFF_COMPRESSION, compression])
    assert np.array_equal(tiff_min.read_tiff(p), img)
    rgb = np.random.default_rng(2).integers(0, 65535, (120, 90, 3)).astype(np.uint16)
    assert cv2.imwrite(p, rgb, [cv2.IMWRITE_TIFF_COMPRESSION, compression])
    assert np.array_equal(tiff_min.read_tiff(p), rgb[..., ::-1])            # OpenCV stores BGR as RGB


def test_tiff_reads_pillow_files_and_stacks_pages(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    img = _field()
    for comp in ("tiff_lzw", "tiff_adobe_deflate", "packbits", None):
        p = str(tmp_path / "p.tif")
        Image.fromarray(img).save(p, compression=comp)
        assert np.array_equal(tiff_min.read_tiff(p), img), comp
    pages = [Image.fromarray((img + i).astype(np.uint16)) for i in range(3)]
    p = str(tmp_path / "stack.tif")
    pages[0].save(p, save_all=True, append_images=pages[1:], compression="tiff_lzw")
    back = tiff_min.read_tiff(p)                                             # tifffile.imread returns the stack
    assert back.shape == (3,) + img.shape and np.array_equal(back[2], img + 2)


@pytest.mark.parametrize("kw", [dict(tile=(64, 48 + 16)), dict(tile=(128, 128), compression=8),
                                dict(compression=8, rows_per_strip=7), dict(bigtiff=True),
                                dict(bigtiff=True, tile=(32, 32), compression=8)])
def test_tiff_tiles_deflate_bigtiff_round_trip(tmp_path, kw):
    for img in (_field((150, 203)), np.random.default_rng(3).integers(0, 65535, (70, 50, 3)).astype(np.uint16)):
        p = str(tmp_path / "t.tif")
        tiff_min.write_tiff(p, img, **kw)
        assert np.array_equal(tiff_min.read_tiff(p), img)


def test_tiff_tiled_file_matches_opencv_decoder(tmp_path):
    cv2 = pytest.importorskip("cv2")
    img = _field((150, 203))
    p = str(tmp_path / "t.tif")
    tiff_min.write_tiff(p, img, tile=(64, 64), compression=8)
    assert np.array_equal(cv2.imread(p, cv2.IMREAD_UNCHANGED), img)


def test_native_and_python_lzw_agree(tmp_path):
    """The interpreter fallback decodes the same stream to the same bytes as csrc/host_tiff.cpp."""
    cv2 = pytest.importorskip("cv2")
    import struct
    img = _field((64, 200))
    p = str(tmp_path / "l.tif")
    cv2.imwrite(p, img, [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    buf = open(p, "rb").read()
    # first strip of the file, located through the reader's own IFD parser
    (ifd,) = struct.unpack("<I", buf[4:8])
    (n,) = struct.unpack("<H", buf[ifd:ifd + 2])
    tags = {}
    for i in range(n):
        e = buf[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ, cnt = struct.unpack("<HHI", e[:8])
        tags[tag] = tiff_min._ifd_values(buf, "<", typ, cnt, e[8:12], struct.unpack("<I", e[8:12])[0])
    strip = buf[tags[273][0]: tags[273][0] + tags[279][0]]
    cap = tags[278][0] * 200 * 2
    native = tiff_min._decompress(strip, 5, cap)
    assert tiff_min._lzw_python(strip, cap) == native and len(native) == cap


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


def test_training_data_quality_report_layout(tmp_path):
    """data_quality_report.txt as CAE_improved_modeltrain.py:159-182 lays it out (sample std, ddof = 1); checked
    byte for byte against the reference's own function when this file was written"""
    import re
    pd = pytest.importorskip("pandas")
    from cell_image_analysis_b200.training import ImprovedAnomalyDetectionTraining
    t = ImprovedAnomalyDetectionTraining.__new__(ImprovedAnomalyDetectionTraining)     # no device needed for the report
    t.output_dir = str(tmp_path)
    stats = pd.DataFrame({"area": [400.0, 600.0, 800.0], "eccentricity": [0.5, 0.6, 0.7], "solidity": [0.9, 0.95, 1.0],
                          "mean_intensity": [1.0, 2.0, 3.0], "std_intensity": [0.25, 0.5, 0.75]})
    files = pd.DataFrame({"filename": ["a.tif", "b.tif"], "cells_extracted": [3, 0], "mean_cell_intensity": [2.0, 0]})
    t.generate_data_quality_report(stats, files)
    text = (tmp_path / "data_quality_report.txt").read_text()
    text = re.sub(r"Generated: \d{4}-\d\d-\d\d \d\d:\d\d:\d\d", "Generated: <now>", text)
    assert text == ("=== TRAINING DATA QUALITY REPORT ===\n\nGenerated: <now>\n\n"
                    "OVERALL STATISTICS:\nTotal files processed: 2\nTotal cells extracted: 3\nAverage cells per file: 1.5\n\n"
                    "CELL MORPHOLOGY STATISTICS:\nArea: 600.0 ± 200.0\nEccentricity: 0.600 ± 0.100\nSolidity: 0.950 ± 0.050\n\n"
                    "INTENSITY STATISTICS:\nMean intensity: 2.000 ± 1.000\nStd intensity: 0.500 ± 0.250\n\n"
                    "FILE-WISE SUMMARY:\na.tif: 3 cells, avg intensity: 2.000\nb.tif: 0 cells, avg intensity: 0.000\n")
