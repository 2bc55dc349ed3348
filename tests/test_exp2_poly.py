"""The 2^f polynomial of the tensor-core SVM epilogue (csrc/score_tc.cu, exp2x2): coefficients read
from the source, evaluated here in float32 Horner form over the whole reduced range [-0.5, 0.5] --
the header's claim (max relative error <= 1.0e-7 in fp32, mean ~4e-10, i.e. no systematic bias over
the thousands of kernel values a decision sums) is checked, and so is the magic-number reduction."""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "cell-image-analysis_b200", "csrc", "score_tc.cu")


def _coefficients():
    text = open(SRC).read()
    body = text[text.index("void exp2x2("):text.index("r0f = __uint_as_float")]
    vals = [float(m) for m in re.findall(r"pack2\(([0-9.eE+-]+)f, [0-9.eE+-]+f\)", body)]
    # [magic, -magic, -1 (f = t - n)], then c6 .. c0 in Horner order
    assert vals[0] == 12582912.0 and vals[1] == -12582912.0 and vals[2] == -1.0
    return [np.float32(v) for v in vals[3:]]


def test_polynomial_error_in_float32():
    c = _coefficients()
    assert len(c) == 7 and c[-1] == np.float32(1.0)
    f = np.linspace(-0.5, 0.5, 400001).astype(np.float32)
    p = np.full_like(f, c[0])
    for ck in c[1:]:
        # fma: exact product + add, rounded once (float64 holds the float32 product exactly)
        p = (p.astype(np.float64) * f.astype(np.float64) + np.float64(ck)).astype(np.float32)
    ref = np.exp2(f.astype(np.float64))
    rel = (p.astype(np.float64) - ref) / ref
    assert np.abs(rel).max() <= 1.0e-7, np.abs(rel).max()
    assert abs(rel.mean()) <= 2e-9, rel.mean()


def test_magic_number_reduction_and_exponent_add():
    """n = rint(t) from the low mantissa bits of t + 1.5 * 2^23, f = t - n exact, 2^n by an integer add to
    the exponent field -- the same integer arithmetic as the kernel, against numpy's exp2."""
    c = _coefficients()
    rng = np.random.default_rng(0)
    t = np.concatenate([rng.uniform(-126, 2, 200000), [-126.0, -0.5, 0.5, 0.0, 1.5, -125.5]]).astype(np.float32)
    m = (t + np.float32(12582912.0)).astype(np.float32)
    nf = (m - np.float32(12582912.0)).astype(np.float32)
    f = (t - nf).astype(np.float32)
    assert np.all(np.abs(f) <= 0.5) and np.array_equal(nf, np.rint(t))          # ties to even like the hardware add
    p = np.full_like(f, c[0])
    for ck in c[1:]:
        p = (p.astype(np.float64) * f.astype(np.float64) + np.float64(ck)).astype(np.float32)
    bits = (p.view(np.uint32).astype(np.uint64) + ((m.view(np.uint32).astype(np.uint64) << np.uint64(23)) & np.uint64(0xFFFFFFFF))) \
        & np.uint64(0xFFFFFFFF)
    r = bits.astype(np.uint32).view(np.float32)
    ref = np.exp2(t.astype(np.float64))
    ok = t > -125.6            # at the clamp the result leaves the normal range (a term < 1e-37 of a sum of O(1) terms)
    assert np.abs(r[ok].astype(np.float64) / ref[ok] - 1).max() <= 1.0e-7
