"""The autoencoder's multi-pass path: one ``cia_cae_forward`` / ``cia_screen_fields`` call walks
its cells in passes of ``cae_pass_cells`` (default 18944 = 128 per SM; cae_tc.cu re-bases the
activation buffers and the TMA cell coordinate per pass).  bench.py's default workload -- 64-field
fused calls, ~30k cells -- always takes two passes, so that path is tested here at the
north_star gates and for invariance under the pass size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tc_screener(model_dir):
    from cell_image_analysis_b200.screening import ProductionMutantScreening
    s = ProductionMutantScreening(model_dir, segmenter=lambda ch: None, device=0, precision=1)
    yield s
    s.engine.close()


def _forward(eng, x, n, pass_cells):
    eng.set_option("cae_pass_cells", pass_cells)
    try:
        mse, mae, feat = eng.cae_forward(x, n, precision=1)
        dc, dm, pc, pm, _ = eng.svm_decision(feat, n)
        eng.check_status()
        return [t[:n].cpu().numpy() for t in (mse, mae, feat, dc, dm, pc, pm)]
    finally:
        eng.set_option("cae_pass_cells", 18944)


def test_pass_size_invariance(tc_screener, field_config1):
    """The same cells through 1 pass and through 4 / 2 ragged passes: bit-identical outputs."""
    green, labels = field_config1
    cells, _ = tc_screener.extract_quality_cells_from_labels(green, labels)
    eng = tc_screener.engine
    n = len(cells)
    x = torch.from_numpy(np.array(cells).astype(np.float32)).to(eng.tdev)
    one = _forward(eng, x, n, 18944)
    for pc in (148, 300):
        assert n > pc
        many = _forward(eng, x, n, pc)
        for a, b, name in zip(one, many, ("mse", "mae", "features", "dec_cons", "dec_mod", "pred_cons", "pred_mod")):
            assert np.array_equal(a, b), f"pass size {pc}: {name} differs from the single pass"


def test_second_pass_of_a_bench_sized_call(tc_screener, golden_config1):
    """88 field visits (8 seeded config-1 fields x 11) = ~41k cells through ONE fused call at the
    default precision and pass size: three passes.  Every visit of the seed-0 field -- the last ones
    sit in the second and third pass -- must meet the gates against the oracle's golden vectors,
    and all visits of one field must agree bit for bit."""
    from cell_image_analysis_b200 import _lib, synth
    eng = tc_screener.engine
    NV = 88
    fields = [synth.make_field(s) for s in range(8)]
    H, W = fields[0][0].shape
    max_label = max(int(l.max()) for _, l in fields)
    pool_g = torch.from_numpy(np.stack([g for g, _ in fields]).view(np.int16)).to(eng.tdev)
    pool_l = torch.from_numpy(np.stack([l for _, l in fields])).to(eng.tdev)
    visits = torch.arange(NV, device=eng.tdev) % 8
    g, l = pool_g[visits], pool_l[visits]
    out = eng.alloc_outputs(NV * max_label, NV, keep_features=False)
    eng.screen_fields(g, l, max_label, out, precision=1)
    eng.check_status()
    counts = out["counts"].cpu().numpy()
    n = int(counts[0])
    assert n > 2 * 18944, n                              # the call really took more than two passes
    per_field = counts[1:]
    starts = np.concatenate([[0], np.cumsum(per_field)])
    cells = out["cells"][:n].cpu().numpy().view(_lib.CELL_DTYPE).reshape(-1)
    res = {k: out[k][:n].cpu().numpy() for k in ("mse", "mae", "dec_cons", "dec_mod", "pred_cons", "pred_mod")}
    gold = golden_config1
    worst = 0.0
    for v in range(NV):
        s0, s1 = starts[v], starts[v + 1]
        f = v % 8
        first = slice(starts[f], starts[f + 1])
        assert np.array_equal(cells["label"][s0:s1], cells["label"][first])
        for k in res:
            assert np.array_equal(res[k][s0:s1], res[k][first]), f"visit {v} of field {f}: {k} differs from its first visit"
        if f == 0:
            assert np.array_equal(cells["label"][s0:s1], gold["kept_labels"])
            np.testing.assert_allclose(res["mse"][s0:s1], gold["mse"], rtol=1e-3)
            np.testing.assert_allclose(res["mae"][s0:s1], gold["mae"], rtol=1e-3)
            for dec, pred in (("dec_cons", "pred_cons"), ("dec_mod", "pred_mod")):
                d = np.abs(res[dec][s0:s1] - gold[dec])
                worst = max(worst, d.max())
                assert d.max() <= 1e-4, f"visit {v} (cells {s0}..{s1}): max |d dec| {d.max():.3e}"
                far = np.abs(gold[dec]) > 1e-4
                assert np.array_equal(res[pred][s0:s1][far], gold[pred][far])
    in_late_pass = [v for v in range(0, NV, 8) if starts[v] >= 18944]
    assert len(in_late_pass) >= 4
    print(f"{n} cells in one call ({(n + 18943) // 18944} passes); seed-0 visits starting in pass >= 2: "
          f"{in_late_pass}; worst |d dec| vs oracle {worst:.3e}")
