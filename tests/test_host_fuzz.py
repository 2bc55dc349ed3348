"""The host-only C++ of libcia parses untrusted bytes (TIFF strips: LZW / PackBits) and writes into caller-sized
slots (run-length label encoder).  This builds tests/host_fuzz/fuzz_host.cpp together with those two sources under
AddressSanitizer + UndefinedBehaviorSanitizer and runs it: exact round trips of valid streams, clipping at short
capacities, no out-of-bounds access on corrupted / truncated / random input, AVX2 and scalar encoder paths."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cell-image-analysis_b200", "csrc")


@pytest.fixture(scope="module")
def fuzz_binary(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("fuzz") / "fuzz_host")
    cmd = [gxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           os.path.join(ROOT, "tests", "host_fuzz", "fuzz_host.cpp"), os.path.join(CSRC, "host_tiff.cpp"),
           os.path.join(CSRC, "host_rle.cpp"), "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and ("asan" in r.stdout or "ubsan" in r.stdout):
        pytest.skip("sanitizer runtimes not installed")
    assert r.returncode == 0, r.stdout
    return exe


@pytest.mark.parametrize("scalar", [False, True])
def test_host_decoders_and_encoder_under_sanitizers(fuzz_binary, scalar):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    if scalar:
        env["CIA_HOST_RLE_SCALAR"] = "1"
    r = subprocess.run([fuzz_binary, "1500", "600"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok:"), r.stdout[-3000:]
    assert int(r.stdout.split()[1]) == 2 * 1500 + 600


def test_threaded_host_encoders_under_sanitizers(tmp_path):
    """cia_rle_encode_fields / cia_rle_encode_pack_fields (csrc/transport.cu, host functions only)"""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    exe = str(tmp_path / "fuzz_transport")
    cmd = [nvcc, "-O1", "-g", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=all",
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_fuzz", "fuzz_transport.cpp"),
           os.path.join(CSRC, "transport.cu"), os.path.join(CSRC, "host_rle.cpp"), "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and ("asan" in r.stdout or "ubsan" in r.stdout):
        pytest.skip("sanitizer runtimes not installed")
    assert r.returncode == 0, r.stdout[-3000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0")
    r = subprocess.run([exe, "600"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok:"), r.stdout[-3000:]
    encoded, refused = int(r.stdout.split()[1]), int(r.stdout.split()[3])
    assert encoded + refused == 600 and encoded > 100 and refused > 50
