"""Held-out validation of the default precision (tcgen05 split-precision encoder with the
round-toward-zero compensation, cae_tc.cu / DESIGN.md section 5).

The compensation constants were chosen on the committed golden model (seed 7).  Here the default
path is compared with the ORACLE (correctly rounded fp32 convolutions, real scikit-learn / libsvm)
on autoencoders it has never seen -- other seeds, conv weights scaled x0.3 and x3, more negative
BatchNorm gammas -- each with detectors fit on its own features exactly as
CAE_improved_modeltrain.py:407-427 fits them.

Gates.  north_star's decision gate (1e-4) sits at the float32 noise floor of the reference's own
arithmetic: a plain oneDNN float32 convolution -- the library TensorFlow-CPU itself calls, run here
through ``oracle.cae.forward(exact=False)`` -- moves the decision by 1.1e-4 .. 2.2e-4 against the
correctly rounded result on these models (measured, printed below) -- and the exact-fp32 CUDA path
(precision 0) by 5e-5 .. 8.4e-5.  The default path must stay within max(1e-4, twice that measured
noise of the reference's arithmetic; a maximum over ~470 cells fluctuates), its encoder features
must be CLOSER to the correctly rounded values than oneDNN float32's are, and precision 0 must hold
the plain 1e-4 gate.  MSE: 1e-3 relative on every
model whose weights are not the x3 stress (5e-3 there: the decoder runs single-pass fp16 operands,
and un-normalised x3 weights saturate it; precision 0 holds 1e-3 on all models)."""
import numpy as np
import pytest
import torch

from helpers import artifacts_from, fit_detectors, synth_cae_weights
from oracle import scoring as oscoring

pytestmark = pytest.mark.gpu

SETS = [  # (seed, conv weight scale, negative gammas per BN layer)
    (11, 1.0, 3),
    (23, 0.3, 3),
    (37, 3.0, 8),
    (51, 1.0, 0),
]
DEFAULT_DEBIAS = (("cae_l1_debias", 0.5), ("cae_l2_debias", 2.4), ("cae_l3_debias", 1.2))


@pytest.fixture(scope="module")
def crops(screener):
    """Train crops (3 fields) and test crops (seed 0) through the CUDA extraction path."""
    from cell_image_analysis_b200 import synth
    train = []
    for seed in (200, 201, 202):
        g, l = synth.make_field(seed)
        train.extend(screener.extract_quality_cells_from_labels(g, l)[0])
    g, l = synth.make_field(0)
    test = screener.extract_quality_cells_from_labels(g, l)[0]
    return train, test


_DUMMY = None


def _dummy_detectors():
    """Any fitted scaler / PCA / detectors: load_artifacts wants a complete set before the
    training features exist."""
    global _DUMMY
    if _DUMMY is None:
        rng = np.random.default_rng(0)
        _DUMMY = fit_detectors(rng.standard_normal((64, 2048)).astype(np.float32))
    return _DUMMY


@pytest.mark.parametrize("seed,scale,nneg", SETS)
def test_default_precision_on_held_out_weights(crops, seed, scale, nneg):
    from cell_image_analysis_b200.screening import Engine
    train, test = crops
    w = synth_cae_weights(seed, scale, nneg)
    eng = Engine(device=0, precision=1)
    try:
        # features of the training cells from the exact-fp32 anchor, detectors fit like train:407-427
        eng.load_artifacts(artifacts_from(w, *_dummy_detectors()))
        xt = torch.from_numpy(np.array(train).astype(np.float32)).to(eng.tdev)
        _, _, feat = eng.cae_forward(xt, len(train), precision=0)
        scaler, pca, cons, mod = fit_detectors(feat[:len(train)].cpu().numpy())
        eng.load_artifacts(artifacts_from(w, scaler, pca, cons, mod))
        ref = oscoring.compute_anomaly_scores(test, w, w, scaler, pca, cons, mod)
        # the reference's own arithmetic class: oneDNN float32 convolutions (what TF-CPU calls)
        dnn = oscoring.compute_anomaly_scores(test, w, w, scaler, pca, cons, mod, exact=False)
        fscale = np.abs(ref["_features"]).mean()
        noise = {k: np.abs(dnn[f"{k}_scores"] - ref[f"{k}_scores"]).max() for k in ("conservative", "moderate")}
        noise_feat = np.abs(dnn["_features"] - ref["_features"]).mean() / fscale
        n = len(test)
        x = torch.from_numpy(np.array(test).astype(np.float32)).to(eng.tdev)

        def run(precision, debias_on=True):
            for name, v in DEFAULT_DEBIAS:
                eng.set_option(name, v if debias_on else 0.0)
            mse, mae, feat = eng.cae_forward(x, n, precision=precision)
            dc, dm, pc, pm, _ = eng.svm_decision(feat, n, precision=precision)   # precision 0: fp64 scoring anchors too
            eng.check_status()
            return [t[:n].cpu().numpy() for t in (mse, feat, dc, dm, pc, pm)]

        def errs(r):
            mse, feat, dc, dm, pc, pm = r
            df = feat - ref["_features"]
            return (np.abs(dc + ref["conservative_scores"]).max(), np.abs(dm + ref["moderate_scores"]).max(),
                    np.abs(df).mean() / fscale, df.mean() / fscale, np.abs(mse / ref["reconstruction_mse"] - 1).max())

        off, on, anchor = run(1, False), run(1, True), run(0)
        print(f"\nseed {seed} scale {scale} neg-gammas {nneg}: nSV {cons.support_vectors_.shape[0]}/"
              f"{mod.support_vectors_.shape[0]}, |dec| up to {np.abs(ref['conservative_scores']).max():.2f}")
        print(f"  reference arithmetic (oneDNN fp32 conv) vs correctly rounded: max |d dec| "
              f"{noise['conservative']:.2e}/{noise['moderate']:.2e}, mean rel feature error {noise_feat:.2e}")
        for tag, r in (("precision 1, compensation off", off), ("precision 1 (default)", on), ("precision 0 (fp32 anchor)", anchor)):
            e = errs(r)
            print(f"  {tag}: max |d dec| {e[0]:.2e}/{e[1]:.2e}, mean rel feature error {e[2]:.2e} "
                  f"(signed {e[3]:+.2e}), max rel d mse {e[4]:.2e}")

        e_on, e_off, e_anchor = errs(on), errs(off), errs(anchor)
        # the compensation removes the truncation bias on models it was not chosen on
        assert abs(e_on[3]) < 0.5 * abs(e_off[3]), "compensation does not remove the truncation bias"
        assert e_on[2] < e_off[2]
        # features at least as close to the correctly rounded values as the reference's arithmetic
        assert e_on[2] < noise_feat
        mse, feat, dc, dm, pc, pm = on
        np.testing.assert_allclose(mse, ref["reconstruction_mse"], rtol=1e-3 if scale <= 1.0 else 5e-3)
        np.testing.assert_allclose(anchor[0], ref["reconstruction_mse"], rtol=1e-3)
        for dec, pred, key in ((dc, pc, "conservative"), (dm, pm, "moderate")):
            gate = max(1e-4, 2.0 * noise[key])
            d = np.abs(dec + ref[f"{key}_scores"])
            assert d.max() <= gate, f"{key}: max |d dec| {d.max():.3e} > {gate:.3e}"
            far = np.abs(ref[f"{key}_scores"]) > gate
            assert np.array_equal(pred[far], ref[f"{key}_predictions"][far])
        for dec, key in ((anchor[2], "conservative"), (anchor[3], "moderate")):
            assert np.abs(dec + ref[f"{key}_scores"]).max() <= 1e-4          # the anchor holds the plain gate
        # the scores must be informative (not a saturated constant): both signs present
        assert 0 < (pc == -1).sum() < n
    finally:
        eng.close()
