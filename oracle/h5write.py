"""Minimal HDF5 *writer* + ``.keras`` archive writer -- FIXTURE TOOLING ONLY.

The reference saves its CAE with Keras (CAE_improved_modeltrain.py:270-275,299-300);
neither Keras nor h5py exists in this image, so test fixtures are synthesised here
in the on-disk shape h5py's default (``libver='earliest'``) produces: superblock v0,
v1 object headers, symbol-table groups (B-tree v1 + SNOD with leaf K = 4 + local
heap), contiguous little-endian IEEE datasets.  Never imported by the product.
"""
from __future__ import annotations

import io
import json
import struct
import zipfile

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, n, align=8):
        while len(self.b) % align:
            self.b.append(0)
        off = len(self.b)
        self.b.extend(b"\0" * n)
        return off

    def put(self, off, data):
        self.b[off:off + len(data)] = data


def _msg(mtype, data, flags=0):
    pad = (-len(data)) % 8
    return struct.pack("<HHB3x", mtype, len(data) + pad, flags) + data + b"\0" * pad


def _header(buf, msgs):
    body = b"".join(msgs)
    off = buf.alloc(16 + len(body))
    buf.put(off, struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)
    return off


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        exp_bits, mant = {4: (8, 23), 8: (11, 52), 2: (5, 10)}[size]
        bias = (1 << (exp_bits - 1)) - 1
        return struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size) + \
            struct.pack("<HHBBBBI", 0, size * 8, mant, exp_bits, 0, mant, bias)
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, dt.itemsize * 8)
    raise TypeError(dt)


def _write_dataset(buf, arr):
    arr = np.ascontiguousarray(arr)
    le = arr.astype(arr.dtype.newbyteorder("<"), copy=False)
    raw = le.tobytes()
    daddr = buf.alloc(max(len(raw), 1))
    buf.put(daddr, raw)
    space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    fill = struct.pack("<BBBB", 2, 2, 2, 0)
    layout = struct.pack("<BBQQ", 3, 1, daddr, len(raw))
    return _header(buf, [_msg(1, space), _msg(3, _dtype_msg(arr.dtype), 1), _msg(5, fill),
                         _msg(8, layout)])


def _write_group(buf, tree):
    """tree: dict name -> ndarray | dict.  Returns object header address."""
    children = {}
    for name, node in tree.items():
        children[name] = _write_group(buf, node) if isinstance(node, dict) else _write_dataset(buf, node)
    names = sorted(children, key=lambda s: s.encode())
    # local heap: offset 0 is the empty string
    seg = bytearray(b"\0" * 8)
    offs = {}
    for n in names:
        offs[n] = len(seg)
        seg += n.encode() + b"\0"
        while len(seg) % 8:
            seg.append(0)
    seg_len = max(len(seg) + 16, 88)
    heap = buf.alloc(32)
    segaddr = buf.alloc(seg_len)
    free_off = len(seg)
    seg = seg + struct.pack("<QQ", 1, seg_len - len(seg)) + b"\0" * (seg_len - len(seg) - 16)
    buf.put(segaddr, bytes(seg))
    buf.put(heap, b"HEAP" + struct.pack("<B3xQQQ", 0, seg_len, free_off, segaddr))
    # SNODs
    snods, keys = [], [0]
    per = 2 * LEAF_K
    for i in range(0, max(len(names), 1), per):
        part = names[i:i + per]
        sn = buf.alloc(8 + per * 40)
        body = b"SNOD" + struct.pack("<BxH", 1, len(part))
        for n in part:
            body += struct.pack("<QQII16x", offs[n], children[n], 0, 0)
        buf.put(sn, body)
        snods.append(sn)
        keys.append(offs[part[-1]] if part else 0)
    assert len(snods) <= 2 * INTERNAL_K, "group too large for a single-level B-tree"
    bt = buf.alloc(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8)
    body = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
    for i, sn in enumerate(snods):
        body += struct.pack("<QQ", keys[i], sn)
    body += struct.pack("<Q", keys[len(snods)])
    buf.put(bt, body)
    return _header(buf, [_msg(0x11, struct.pack("<QQ", bt, heap))])


def write_h5(tree: dict) -> bytes:
    buf = _Buf()
    sb = buf.alloc(96)
    root = _write_group(buf, tree)
    # root group's btree/heap for the superblock scratch pad
    # (readers that ignore the cache read the symbol-table message instead)
    eof = len(buf.b)
    sbd = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBxHHI", 0, 0, 0, 0, 0, 8, 8, LEAF_K, INTERNAL_K, 0)
    sbd += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sbd += struct.pack("<QQII16x", 0, root, 0, 0)
    buf.put(sb, sbd)
    return bytes(buf.b)


# ---------------------------------------------------------------------------
# .keras archive (Keras 3 saving_lib layout: metadata.json, config.json, model.weights.h5)
# ---------------------------------------------------------------------------

def _layer_cfg(cls, name, **cfg):
    return {"module": "keras.layers", "class_name": cls,
            "config": dict(name=name, trainable=True, dtype="float32", **cfg),
            "registered_name": None, "name": name}


def cae_config(encoder_only=False, name_offset=0):
    """config.json of the Functional model of train:188-219 (layer *names* carry the
    session-global counters Keras assigns; weights paths use per-class counters)."""
    def nm(base, k):
        k += name_offset
        return base if k == 0 else f"{base}_{k}"
    layers = [_layer_cfg("InputLayer", nm("input_layer", 0), batch_shape=[None, 64, 64, 1])]
    filt = [32, 64, 32, 32, 64, 32, 1]
    n = 3 if encoder_only else 7
    pool = up = 0
    for i in range(n):
        layers.append(_layer_cfg("Conv2D", nm("conv2d", i), filters=filt[i], kernel_size=[3, 3],
                                 strides=[1, 1], padding="same", data_format="channels_last",
                                 dilation_rate=[1, 1], groups=1,
                                 activation="sigmoid" if i == 6 else "relu", use_bias=True))
        if i < 6:
            layers.append(_layer_cfg("BatchNormalization", nm("batch_normalization", i), axis=-1,
                                     momentum=0.99, epsilon=0.001, center=True, scale=True))
            if i < 3:
                layers.append(_layer_cfg("MaxPooling2D", nm("max_pooling2d", pool), pool_size=[2, 2],
                                         padding="same", strides=[2, 2], data_format="channels_last"))
                pool += 1
            else:
                layers.append(_layer_cfg("UpSampling2D", nm("up_sampling2d", up), size=[2, 2],
                                         data_format="channels_last", interpolation="nearest"))
                up += 1
    return {"module": "keras", "class_name": "Functional",
            "config": {"name": "functional", "trainable": True, "layers": layers,
                       "input_layers": [[layers[0]["name"], 0, 0]],
                       "output_layers": [[layers[-1]["name"], 0, 0]]},
            "registered_name": "Functional",
            "compile_config": None if encoder_only else {"optimizer": "adam", "loss": "mse"}}


def write_keras(path, w, encoder_only=False, name_offset=0, with_optimizer=True):
    """Write ``w`` (kernels/biases/bns as in oracle.cae) as a ``.keras`` zip."""
    n = 3 if encoder_only else 7
    layers = {}
    for i in range(n):
        key = "conv2d" if i == 0 else f"conv2d_{i}"
        layers[key] = {"vars": {"0": np.asarray(w["kernels"][i], np.float32),
                                "1": np.asarray(w["biases"][i], np.float32)}}
        if i < 6:
            key = "batch_normalization" if i == 0 else f"batch_normalization_{i}"
            g, b, m, v = w["bns"][i]
            layers[key] = {"vars": {"0": np.asarray(g, np.float32), "1": np.asarray(b, np.float32),
                                    "2": np.asarray(m, np.float32), "3": np.asarray(v, np.float32)}}
    layers["input_layer"] = {"vars": {}}
    for i in range(min(n, 3)):
        layers["max_pooling2d" if i == 0 else f"max_pooling2d_{i}"] = {"vars": {}}
    for i in range(max(0, min(n, 6) - 3)):
        layers["up_sampling2d" if i == 0 else f"up_sampling2d_{i}"] = {"vars": {}}
    tree = {"layers": layers, "vars": {}}
    if with_optimizer and not encoder_only:
        tree["optimizer"] = {"vars": {"0": np.array(1234, np.int64),
                                      "1": np.array(1e-3, np.float32),
                                      "2": np.zeros((3, 3, 1, 32), np.float32)}}
    h5 = write_h5(tree)
    meta = {"keras_version": "3.3.3", "date_saved": "2025-01-01@00:00:00"}
    bio = io.BytesIO()
    with zipfile.ZipFile(bio, "w", zipfile.ZIP_STORED) as z:
        for name, data in (("metadata.json", json.dumps(meta)),
                           ("config.json", json.dumps(cae_config(encoder_only, name_offset))),
                           ("model.weights.h5", h5)):
            z.writestr(zipfile.ZipInfo(name, date_time=(2025, 1, 1, 0, 0, 0)), data)  # reproducible bytes
    with open(path, "wb") as f:
        f.write(bio.getvalue())
