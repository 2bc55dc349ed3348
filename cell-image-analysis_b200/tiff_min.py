"""Minimal baseline-TIFF reader/writer for 16-bit microscopy fields.

The reference reads fields with ``tiff.imread`` (improved_detection.py:51); tifffile is not
installable here.  This covers what an acquisition system writes for such fields: classic
TIFF (not BigTIFF), little or big endian, uncompressed, strips, 8/16/32-bit unsigned or
signed integer samples, 1..N samples per pixel in chunky (interleaved) or planar layout.
Anything else (compression, tiles, BigTIFF, multi-page stacks) raises ``TiffError`` --
``_default_imread`` then falls back to tifffile / OpenCV when they are present.
"""
from __future__ import annotations

import struct

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1),
          8: ("h", 2), 9: ("i", 4), 16: ("Q", 8)}


class TiffError(ValueError):
    pass


def _ifd_values(buf, bo, typ, count, value_field, pos):
    if typ not in _TYPES:
        return None
    fmt, size = _TYPES[typ]
    n = count * size
    data = value_field[:n] if n <= 4 else buf[pos:pos + n]
    if typ == 5:
        v = struct.unpack(bo + "II" * count, data)
        return [v[2 * i] / max(v[2 * i + 1], 1) for i in range(count)]
    if typ == 2:
        return [bytes(data)]
    return list(struct.unpack(bo + fmt * count, data))


def read_tiff(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 8:
        raise TiffError("not a TIFF file")
    if buf[:2] == b"II":
        bo = "<"
    elif buf[:2] == b"MM":
        bo = ">"
    else:
        raise TiffError("not a TIFF file")
    magic, ifd = struct.unpack(bo + "HI", buf[2:8])
    if magic == 43:
        raise TiffError("BigTIFF is not supported")
    if magic != 42:
        raise TiffError("not a TIFF file")
    (n_entries,) = struct.unpack(bo + "H", buf[ifd:ifd + 2])
    tags = {}
    for i in range(n_entries):
        e = buf[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ, count = struct.unpack(bo + "HHI", e[:8])
        (off,) = struct.unpack(bo + "I", e[8:12])
        tags[tag] = _ifd_values(buf, bo, typ, count, e[8:12], off)
    (next_ifd,) = struct.unpack(bo + "I", buf[ifd + 2 + 12 * n_entries: ifd + 6 + 12 * n_entries])

    def tag1(t, default=None):
        v = tags.get(t)
        return default if not v else v[0]
    W, H = tag1(256), tag1(257)
    if not W or not H:
        raise TiffError("missing image dimensions")
    if tag1(259, 1) != 1:
        raise TiffError("compressed TIFF is not supported")
    if 322 in tags or 324 in tags:
        raise TiffError("tiled TIFF is not supported")
    spp = tag1(277, 1)
    bits = tags.get(258, [1])
    if len(set(bits)) != 1 or bits[0] not in (8, 16, 32):
        raise TiffError(f"unsupported BitsPerSample {bits}")
    fmt = tag1(339, 1)
    if fmt not in (1, 2):
        raise TiffError("only integer sample formats are supported")
    dt = np.dtype(f"{bo}{'u' if fmt == 1 else 'i'}{bits[0] // 8}")
    planar = tag1(284, 1)
    offsets, counts = tags.get(273), tags.get(279)
    if not offsets:
        raise TiffError("missing strip offsets")
    if not counts:
        counts = [H * W * spp * dt.itemsize] if len(offsets) == 1 else None
    if counts is None or len(counts) != len(offsets):
        raise TiffError("inconsistent strip tables")
    raw = b"".join(buf[o:o + c] for o, c in zip(offsets, counts))
    need = H * W * spp * dt.itemsize
    if len(raw) < need:
        raise TiffError("truncated strip data")
    a = np.frombuffer(raw[:need], dt)
    if spp == 1:
        img = a.reshape(H, W)
    elif planar == 2:
        img = a.reshape(spp, H, W).transpose(1, 2, 0)
    else:
        img = a.reshape(H, W, spp)
    _ = next_ifd   # further pages (z / time series) are ignored like a single-plane read
    return np.ascontiguousarray(img.astype(dt.newbyteorder("=")))


def write_tiff(path: str, img: np.ndarray, rows_per_strip: int = 64) -> None:
    """Uncompressed little-endian TIFF (fixture writer for tests and examples)."""
    img = np.ascontiguousarray(img)
    if img.dtype not in (np.uint8, np.uint16, np.uint32):
        raise TiffError("write_tiff: uint8/16/32 only")
    H, W = img.shape[:2]
    spp = 1 if img.ndim == 2 else img.shape[2]
    data = img.astype(img.dtype.newbyteorder("<")).tobytes()
    row_bytes = W * spp * img.itemsize
    strips = [(r * row_bytes, min(rows_per_strip, H - r) * row_bytes) for r in range(0, H, rows_per_strip)]
    bits = img.itemsize * 8
    header = 8
    data_off = header
    ifd_off = data_off + len(data)
    extra = bytearray()
    entries = []

    def put(tag, typ, values):
        fmt, size = _TYPES[typ]
        payload = struct.pack("<" + fmt * len(values), *values)
        if len(payload) <= 4:
            field = payload.ljust(4, b"\0")
        else:
            field = struct.pack("<I", ifd_off + 2 + 12 * N_ENTRIES + 4 + len(extra))
            extra.extend(payload + (b"\0" if len(payload) % 2 else b""))
        entries.append(struct.pack("<HHI", tag, typ, len(values)) + field)
    N_ENTRIES = 11
    put(256, 4, [W]); put(257, 4, [H]); put(258, 3, [bits] * spp); put(259, 3, [1])
    put(262, 3, [1 if spp == 1 else 2]); put(273, 4, [data_off + o for o, _ in strips])
    put(277, 3, [spp]); put(278, 4, [rows_per_strip]); put(279, 4, [c for _, c in strips])
    put(284, 3, [1]); put(339, 3, [1] * spp)
    assert len(entries) == N_ENTRIES
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, ifd_off))
        f.write(data)
        f.write(struct.pack("<H", N_ENTRIES) + b"".join(entries) + struct.pack("<I", 0) + bytes(extra))
