"""GPU parity on BASELINE.json config 3: a dense 4096x4096 field (~5k labels, log-uniform
sizes -> bboxes from ~16 to ~100+ px, overlapping/clipped shapes) stressing the scan, the
gates and every size class of the crop + CLAHE + resize kernel (incl. the gaussian
anti-aliasing branch for sides > 64)."""
import numpy as np
import pytest

from oracle import clahe as oclahe
from oracle import regions as oreg
from oracle import resize as oresize

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def field3():
    from cell_image_analysis_b200 import synth
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["config3"]
    return synth.make_field(0, H, W, n, lo, hi, lu)


def test_config3_regions_gates_crops(screener, field3):
    green, labels = field3
    cells, stats, rec = screener.extract_quality_cells_from_labels(green, labels, return_regions=True)
    kept, tab = oreg.quality_regions(green, labels)
    # scan + gates: identical kept set, in order, exact bbox / area
    assert [k["label"] for k in kept] == rec["label"].tolist()
    assert [k["bbox"] for k in kept] == [tuple(int(v) for v in (r["minr"], r["minc"], r["maxr"], r["maxc"])) for r in rec]
    assert [k["area"] for k in kept] == rec["area"].tolist()
    np.testing.assert_allclose(rec["eccentricity"], [k["eccentricity"] for k in kept], atol=1e-9)
    np.testing.assert_array_equal(rec["mean_intensity"], [k["mean_intensity"] for k in kept])
    assert len(kept) > 2000
    sides = np.maximum(rec["maxr"] - rec["minr"], rec["maxc"] - rec["minc"])
    assert sides.max() > 64, "config 3 should contain cells that need anti-aliasing"
    # crops: a spread of sizes, always including the largest bboxes
    order = np.argsort(sides)
    pick = sorted(set(order[::40].tolist() + order[-12:].tolist()))
    worst = 0.0
    for i in pick:
        r = rec[i]
        crop = green[r["minr"]:r["maxr"], r["minc"]:r["maxc"]]
        ref = oresize.resize(oclahe.equalize_adapthist(crop, clip_limit=0.02), (64, 64))
        d = np.abs(cells[i] - ref)
        tol = 1e-5 * np.maximum(np.abs(ref), 1e-3)
        assert (d <= tol).all(), f"cell {i} bbox {crop.shape}: max |d| {d.max():.3e}"
        worst = max(worst, d.max())
    print(f"config 3: {len(kept)} cells, max side {sides.max()}, {len(pick)} crops checked, worst |d| {worst:.2e}")


def test_config3_scores_match_fp32_anchor(screener, field3):
    """tensor-core scores (default mode) vs the exact-fp32 anchor on ~4k cells."""
    green, labels = field3
    cells, _ = screener.extract_quality_cells_from_labels(green, labels)
    eng = screener.engine
    r0 = screener.compute_anomaly_scores(cells)
    eng.precision = 1
    try:
        r1 = screener.compute_anomaly_scores(cells)
    finally:
        eng.precision = 0
    np.testing.assert_allclose(r1["reconstruction_mse"], r0["reconstruction_mse"], rtol=1e-3)
    for key in ("conservative", "moderate"):
        d = np.abs(r1[f"{key}_scores"] - r0[f"{key}_scores"])
        print(f"config 3 {key}: max |d dec| tc vs fp32 {d.max():.3e}")
        assert d.max() <= 2e-4          # two implementations, each within 1e-4 of the oracle
        far = np.abs(r0[f"{key}_scores"]) > 2e-4
        assert np.array_equal(r1[f"{key}_predictions"][far], r0[f"{key}_predictions"][far])
