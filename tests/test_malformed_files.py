"""Corrupted input files: the TIFF reader (stand-in for tiff.imread, improved_detection.py:51) and the HDF5 reader
behind the .keras artifacts (improved_detection.py:28-29) must answer a damaged file with their own error type
(the reference's catch-all at det:113-115 then reports the file and moves on) -- never with a hang, an unbounded
allocation or a stray IndexError.  Seeded byte flips, truncations and wild 32 / 64-bit values; no GPU."""
import glob
import os
import random
import signal
import zipfile

import numpy as np
import pytest

from cell_image_analysis_b200 import tiff_min
from cell_image_analysis_b200.hdf5_min import H5File, H5FormatError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Deadline:
    def __init__(self, seconds):
        self.seconds = seconds

    def __enter__(self):
        def fire(*_):
            raise TimeoutError("parser did not return")
        self.old = signal.signal(signal.SIGALRM, fire)
        signal.alarm(self.seconds)

    def __exit__(self, *exc):
        signal.alarm(0)
        signal.signal(signal.SIGALRM, self.old)


def _mutate(rng, src, meta_bytes):
    b = bytearray(src)
    mode = rng.randrange(3)
    if mode == 0:
        for _ in range(rng.randrange(1, 7)):
            b[rng.randrange(min(len(b), meta_bytes))] = rng.randrange(256)
    elif mode == 1:
        b = b[:rng.randrange(len(b))]
    else:
        k = rng.randrange(min(len(b), meta_bytes) - 8)
        wild = rng.choice([0xFFFFFFFFFFFFFFFF, rng.randrange(1 << 40), len(b) - 3, 0])
        n = rng.choice([4, 8])
        b[k:k + n] = (wild & ((1 << (8 * n)) - 1)).to_bytes(n, "little")
    return bytes(b)


@pytest.mark.parametrize("compression,tile", [(1, None), (5, None), (8, (32, 48)), (32773, None)])
def test_damaged_tiff_files_raise_tifferror(tmp_path, compression, tile):
    rng = random.Random(compression)
    img = np.random.default_rng(1).integers(0, 65535, (75, 101)).astype(np.uint16)
    img[10:50, 20:80] = 777
    p = str(tmp_path / "good.tif")
    if compression in (1, 8):
        tiff_min.write_tiff(p, img, rows_per_strip=16, compression=compression, tile=tile)
    else:                                        # LZW / PackBits files come from libtiff (OpenCV)
        cv2 = pytest.importorskip("cv2")
        assert cv2.imwrite(p, img, [cv2.IMWRITE_TIFF_COMPRESSION, compression])
    good = open(p, "rb").read()
    assert np.array_equal(tiff_min.read_tiff(p), img)
    q = str(tmp_path / "bad.tif")
    survived = 0
    for _ in range(150):
        with open(q, "wb") as f:
            f.write(_mutate(rng, good, len(good)))
        with _Deadline(20):
            try:
                a = tiff_min.read_tiff(q)
                survived += a.size > 0
            except tiff_min.TiffError:
                pass
    assert survived < 150            # the mutations do bite


def test_damaged_keras_weight_files_raise_h5formaterror():
    rng = random.Random(11)
    sources = []
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "model_dir", "*.keras"))):
        with zipfile.ZipFile(p) as z:
            sources += [z.read(n) for n in z.namelist() if n.endswith(".h5")]
    assert sources
    rejected = 0
    for _ in range(300):
        data = _mutate(rng, rng.choice(sources), 4096)
        with _Deadline(20):
            try:
                H5File(data).datasets()
            except H5FormatError:
                rejected += 1
    assert rejected > 30


def test_object_header_continuation_cycle_is_refused():
    """a version-2 object header whose continuation message points back at its own block"""
    import struct
    sb = bytearray(48)
    sb[:8] = b"\x89HDF\r\n\x1a\n"
    sb[8], sb[9], sb[10] = 2, 8, 8
    struct.pack_into("<QQQQ", sb, 12, 0, 0xFFFFFFFFFFFFFFFF, 0, 48)        # base, extension, eof (unused), root header
    ochk = 48 + 4 + 2 + 1 + 20                                             # address of the continuation block
    cont = struct.pack("<BHB", 0x10, 16, 0) + struct.pack("<QQ", ochk, 4 + 20 + 4)
    hdr = b"OHDR" + bytes([2, 0]) + bytes([len(cont)]) + cont
    blk = b"OCHK" + cont + b"\0\0\0\0"
    data = bytes(sb) + hdr + blk
    assert len(bytes(sb) + hdr) == ochk
    with _Deadline(20), pytest.raises(H5FormatError, match="continuation"):
        H5File(data).datasets()
