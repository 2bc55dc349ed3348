"""The resize stage normalises CLAHE levels with (r - rmin) / (rmax - rmin); crop.cu computes it
as reciprocal + product + one FMA correction instead of an IEEE division.  tests/aux/exact_division.c
checks that sequence against the division for every operand pair the kernel can see
(0 <= a <= den <= 16383): the two must agree bit for bit."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_reciprocal_fma_division_is_exact(tmp_path):
    exe = str(tmp_path / "exact_division")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "aux", "exact_division.c"), "-lm"],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "134225919 quotients, 0 mismatches" in r.stdout


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_scaler_division_by_reciprocal_and_fma_is_exact(tmp_path):
    """score.cu computes RobustScaler's v / scale_ as q = v*r, rem = fma(-q, scale, v), fma(rem, r, q)
    with r = 1/scale rounded on the host; tests/aux/markstein_division.c checks the sequence against
    the IEEE division bit for bit (random float32 numerators x random / all-ones / power-of-two /
    float32-valued double scales)."""
    exe = str(tmp_path / "markstein_division")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "aux", "markstein_division.c"), "-lm"],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "32000000 quotients, 0 mismatches" in r.stdout

