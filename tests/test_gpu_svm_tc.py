"""The tcgen05 GEMM-form SVM kernel (csrc/score_tc.cu, cia_set_option "svm_kernel" = 1, the default)
against the fp64 DMMA anchor (= libsvm to 1e-9, tests/test_gpu_parity.py) and real libsvm:
work distribution over CTAs, ragged sizes, device-side counts, outlier / zero / non-finite rows,
and the fp64 re-evaluation of decisions near zero."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(artifacts):
    from cell_image_analysis_b200.screening import Engine
    e = Engine(device=0, precision=1)
    e.load_artifacts(artifacts)
    yield e
    e.close()


def _features(golden_tiny, n, seed=0):
    rng = np.random.default_rng(seed)
    f0 = golden_tiny["features"].astype(np.float32)
    reps = (n + len(f0) - 1) // len(f0)
    return np.concatenate([f0 * (1 + 0.05 * rng.standard_normal(f0.shape).astype(np.float32))
                           for _ in range(reps)])[:n]


def _both(eng, feat_np, n=None, n_dev=None):
    n = len(feat_np) if n is None else n
    feat = torch.from_numpy(feat_np).to(eng.tdev)
    out = {}
    for k in (0, 1):
        eng.set_option("svm_kernel", k)
        dc, dm, pc, pm, _ = eng.svm_decision(feat, n, n_dev=n_dev)
        eng.check_status()
        out[k] = [t.cpu().numpy() for t in (dc, dm, pc, pm)]
    eng.set_option("svm_kernel", 1)
    return out


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000, 40000])
def test_tc_equals_dmma_across_sizes(eng, golden_tiny, n):
    """1 cell ... 313 cell tiles (more (cell tile, SV tile) pairs than CTAs: several tiles per CTA)."""
    out = _both(eng, _features(golden_tiny, n, seed=n))
    for i in (0, 1):
        d = np.abs(out[1][i][:n] - out[0][i][:n]).max()
        assert d <= 1e-5, f"n={n} det{i}: max |tc - dmma| {d:.3e}"
        assert np.array_equal(out[1][2 + i][:n], out[0][2 + i][:n])


def test_device_side_count_smaller_than_capacity(eng, golden_tiny):
    n_cap, n_real = 3000, 1234
    f = _features(golden_tiny, n_cap, seed=3)
    n_dev = torch.tensor([n_real], dtype=torch.int32, device=eng.tdev)
    a = _both(eng, f, n=n_cap, n_dev=n_dev)
    b = _both(eng, f[:n_real])
    for i in range(4):
        assert np.array_equal(a[1][i][:n_real], b[1][i][:n_real])


def test_outlier_zero_and_nonfinite_rows(eng, golden_tiny):
    f = _features(golden_tiny, 300, seed=5)
    f[3] *= 30.0            # gamma ||z||^2 ~ 1e3: the row exponent leaves the fp32 range (slow path)
    f[40] *= 300.0
    f[77] = 0.0
    f[130] *= 1e-3
    out = _both(eng, f)
    for i in (0, 1):
        d = np.abs(out[1][i] - out[0][i])
        assert np.isfinite(out[1][i]).all()
        assert d.max() <= 1e-5, f"det{i}: max |tc - dmma| {d.max():.3e} at row {d.argmax()}"
        assert np.array_equal(out[1][2 + i], out[0][2 + i])
    f[9, 17] = np.nan
    out = _both(eng, f)
    for i in (0, 1):
        assert np.isnan(out[1][i][9]) and np.isnan(out[0][i][9])
        ok = np.arange(300) != 9
        assert np.abs(out[1][i][ok] - out[0][i][ok]).max() <= 1e-5


def test_decisions_near_zero_are_recomputed_in_fp64(artifacts, golden_tiny):
    """Move rho onto a cell's kernel sum: that decision is ~0 and must come out of the fp64 refine pass
    (bit-identical to what the direct fp64 kernel gives), with the sign libsvm would give."""
    from cell_image_analysis_b200.screening import Engine
    f = _features(golden_tiny, 600, seed=9)
    e = Engine(device=0, precision=1)
    e.load_artifacts(artifacts)
    e.set_option("svm_kernel", 0)
    dc0 = e.svm_decision(torch.from_numpy(f).to(e.tdev), 600)[0].cpu().numpy()
    arts = dict(artifacts)
    cons = dict(artifacts["svm_conservative"])
    target = 123
    cons["rho"] = float(cons["rho"] + dc0[target] - 3e-7)       # new decision of `target`: +3e-7
    arts["svm_conservative"] = cons
    e.load_artifacts(arts)
    res = {}
    for refine in (0, 1):
        e.set_option("svm_kernel", 1)
        e.set_option("svm_refine", refine)
        dc, _dm, pc, _pm, _ = e.svm_decision(torch.from_numpy(f).to(e.tdev), 600)
        res[refine] = (dc.cpu().numpy(), pc.cpu().numpy())
    e.set_option("svm_kernel", 0)
    dc_exact = e.svm_decision(torch.from_numpy(f).to(e.tdev), 600)[0].cpu().numpy()
    e.close()
    assert abs(dc_exact[target] - 3e-7) < 1e-9
    assert abs(res[1][0][target] - dc_exact[target]) < 1e-12 and res[1][1][target] == 1
    untouched = np.abs(dc_exact) > 1e-3
    assert np.array_equal(res[1][0][untouched], res[0][0][untouched])      # the refine pass leaves the rest alone
    assert np.abs(res[1][0] - dc_exact).max() <= 1e-5


@pytest.mark.parametrize("n", [1, 130, 5000, 40000])
def test_pca_tc_equals_dmma_projection(eng, golden_tiny, n):
    """The tcgen05 projection (fp16 hi/lo x 3, stage partials added in fp32) against the fp64 DMMA one, which
    is the exactly rounded float32 flow: differences of a few float32 ulps of the row scale at most, and the
    scaler arithmetic (float32 subtraction, correctly rounded division) is the same code."""
    f = _features(golden_tiny, n, seed=100 + n)
    if n > 200:
        f[7] *= 1e4          # an outlier row: its per-stage power of two differs from its neighbours'
        f[8] *= 1e-4
        f[9] = 0.0
    feat = torch.from_numpy(f).to(eng.tdev)
    z = {}
    for k in (0, 1):
        eng.set_option("pca_kernel", k)
        z[k] = eng.svm_decision(feat, n, want_pca=True)[4][:n].cpu().numpy()
        eng.check_status()
    eng.set_option("pca_kernel", 1)
    scale = np.abs(z[0]).max(axis=1, keepdims=True) + 1e-30
    rel = np.abs(z[1] - z[0]) / scale
    assert np.isfinite(z[1]).all()
    assert rel.max() <= 2e-6, f"n={n}: max |dz| / row max {rel.max():.3e} (row {rel.max(axis=1).argmax()})"
    assert np.median(rel) <= 1e-7


def test_pca_tc_wide_projection_256_components(artifacts, golden_tiny):
    """C = 256 components: two component blocks (blockIdx.y) and no scaler."""
    from cell_image_analysis_b200.screening import Engine
    rng = np.random.default_rng(4)
    q, _ = np.linalg.qr(rng.standard_normal((2048, 256)))
    comp = np.ascontiguousarray(q.T.astype(np.float32)).astype(np.float64)
    arts = dict(artifacts)
    arts["scaler_pca"] = dict(artifacts["scaler_pca"], C=256, center=None, scale=None, components=comp,
                              offset=rng.standard_normal(256) * 0.1, f32_flow=True)
    for k in ("svm_conservative", "svm_moderate"):
        arts[k] = dict(sv=rng.standard_normal((300, 256)) * 3.0, coef=rng.uniform(0, 1, 300), gamma=1.0 / (256 * 9.0), rho=1.0)
    e = Engine(device=0, precision=1)
    e.load_artifacts(arts)
    f = (rng.standard_normal((700, 2048)) * 3.0).astype(np.float32)
    feat = torch.from_numpy(f).to(e.tdev)
    z = {}
    for k in (0, 1):
        e.set_option("pca_kernel", k)
        z[k] = e.svm_decision(feat, 700, want_pca=True)[4][:700].cpu().numpy()
    e.close()
    ref = (f.astype(np.float64) @ comp.T).astype(np.float32) - arts["scaler_pca"]["offset"].astype(np.float32)
    assert np.abs(z[0] - ref).max() <= 1e-5
    rel = np.abs(z[1] - z[0]) / (np.abs(z[0]).max(axis=1, keepdims=True))
    assert rel.max() <= 1e-6, f"{rel.max():.3e}"


def test_models_the_tensor_core_kernels_do_not_serve_take_the_fp64_kernels(artifacts, golden_tiny):
    """Negative dual coefficients, more than 256 PCA dimensions, a feature count that is not a multiple of
    32 and a float64 scaler centre are outside the tcgen05 kernels' contract: the fp64 DMMA kernels run
    instead, silently and exactly (same bits as with the options set to 0)."""
    from cell_image_analysis_b200.screening import Engine
    rng = np.random.default_rng(8)
    f = _features(golden_tiny, 500, seed=21)

    def run(arts, feat, pca_kernels=(1, 0)):
        e = Engine(device=0, precision=1)
        e.load_artifacts(arts)
        x = torch.from_numpy(feat).to(e.tdev)
        out = {}
        for k, pk in zip((1, 0), pca_kernels):
            e.set_option("svm_kernel", k)
            e.set_option("pca_kernel", pk)
            dc, dm, pc, pm, z = e.svm_decision(x, len(feat), want_pca=True)
            e.check_status()
            out[k] = [t[:len(feat)].cpu().numpy() for t in (dc, dm, z)]
        e.close()
        return out

    # (a) a negative dual coefficient in one detector: that detector goes to the fp64 kernel, the other stays
    a = dict(artifacts)
    cons = dict(artifacts["svm_conservative"])
    coef = np.array(cons["coef"], dtype=np.float64).copy()
    coef[3] = -abs(coef[3])
    cons["coef"] = coef
    a["svm_conservative"] = cons
    o = run(a, f, pca_kernels=(0, 0))                          # same z in both runs
    assert np.array_equal(o[1][0], o[0][0])                    # conservative: the fp64 kernel either way
    assert not np.array_equal(o[1][1], o[0][1]) and np.abs(o[1][1] - o[0][1]).max() <= 1e-5   # moderate: tcgen05 vs fp64
    # (b) float64 scaler centre: the projection takes the fp64 kernel, so z is bit-identical under both options
    b = dict(artifacts)
    sp = dict(artifacts["scaler_pca"])
    sp["center"] = np.asarray(sp["center"], dtype=np.float64) + 1e-9
    sp["center_is_f32"] = False
    b["scaler_pca"] = sp
    o = run(b, f)
    assert np.array_equal(o[1][2], o[0][2])
    # (c) 300 PCA dimensions (> 256): both stages on the fp64 kernels
    q, _ = np.linalg.qr(rng.standard_normal((2048, 300)))
    c = dict(artifacts)
    c["scaler_pca"] = dict(artifacts["scaler_pca"], C=300, center=None, scale=None, components=np.ascontiguousarray(q.T),
                           offset=np.zeros(300), f32_flow=True)
    for k in ("svm_conservative", "svm_moderate"):
        c[k] = dict(sv=rng.standard_normal((200, 300)) * 3.0, coef=rng.uniform(0, 1, 200), gamma=1.0 / (300 * 9.0), rho=1.0)
    o = run(c, (rng.standard_normal((300, 2048)) * 3.0).astype(np.float32))
    assert np.abs(o[1][2] - o[0][2]).max() <= 2e-6 * np.abs(o[0][2]).max()      # z: tcgen05 projection (3 column blocks)
    assert np.abs(o[1][0] - o[0][0]).max() <= 1e-5                               # decisions: fp64 SVM on nearly equal z
