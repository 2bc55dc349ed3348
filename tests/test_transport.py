"""Run-length label transport (csrc/transport.cu): the host encoder is CPU code in the C-ABI
library and is checked here without a GPU; the device expansion is a GPU test."""
import ctypes as C

import numpy as np
import pytest

from cell_image_analysis_b200 import _lib
from cell_image_analysis_b200.synth import make_field


def _decode(slot, H, W):
    r0 = (H + 2) & ~1
    out = np.empty((H, W), np.int32)
    for y in range(H):
        lo, hi = int(slot[y]), int(slot[y + 1])
        runs = slot[r0 + 2 * lo: r0 + 2 * hi].reshape(-1, 2)
        assert runs[0, 0] == 0
        xs = list(runs[:, 0]) + [W]
        for (x0, lab), x1 in zip(runs, xs[1:]):
            out[y, x0:x1] = np.uint32(lab).astype(np.int32)
    assert int(slot[H]) == hi
    return out


def _fields(H, W):
    rng = np.random.default_rng(5)
    cells = make_field(3, H, W, 30, 6.0, 14.0)[1]
    stripes = (np.arange(H * W, dtype=np.int32).reshape(H, W) // 9) % 1000        # runs crossing the 8-wide skip
    short = np.repeat(rng.integers(0, 70000, (H, W // 2 + 1), dtype=np.int32), 2, axis=1)[:, :W]
    big = np.full((H, W), 2**31 - 1, np.int32)
    return np.ascontiguousarray(np.stack([cells, np.zeros_like(cells), stripes, short, big]))


@pytest.mark.parametrize("shape", [(64, 64), (37, 101), (5, 7), (128, 1030)])
@pytest.mark.parametrize("threads", [1, 3])
def test_rle_encode_roundtrip_cpu(shape, threads):
    lib = _lib.load()
    H, W = shape
    labs = _fields(H, W)
    F = labs.shape[0]
    sw = 2 * H * W + H + 4            # generous slot: even one run per pixel fits
    slots = np.zeros((F, sw), np.uint32)
    words = np.zeros(F, np.uint32)
    mx = C.c_int32(-1)
    rc = lib.cia_rle_encode_fields(labs.ctypes.data, F, H, W, slots.ctypes.data, sw, words.ctypes.data,
                                   C.byref(mx), threads)
    assert rc == 0
    assert mx.value == labs.max()
    for f in range(F):
        assert np.array_equal(_decode(slots[f], H, W), labs[f])
        n_runs = int(slots[f][H])
        assert words[f] == ((H + 2) & ~1) + 2 * n_runs
        assert n_runs == int((np.diff(labs[f], axis=1) != 0).sum()) + H


def test_rle_encode_capacity_cpu():
    lib = _lib.load()
    H, W = 32, 64
    noise = np.random.default_rng(0).integers(0, 1000, (1, H, W), dtype=np.int32)
    sw = lib.cia_rle_slot_words(H, W)
    slots = np.zeros((1, sw), np.uint32)
    words = np.zeros(1, np.uint32)
    rc = lib.cia_rle_encode_fields(noise.ctypes.data, 1, H, W, slots.ctypes.data, sw, words.ctypes.data, None, 1)
    assert rc == _lib.CIA_E_CAPACITY and words[0] == 0
    assert lib.cia_rle_encode_fields(None, 1, H, W, slots.ctypes.data, sw, words.ctypes.data, None, 1) == _lib.CIA_E_ARG


def test_run_closed_forms_equal_pixel_sums_cpu():
    """The algebra of label_scan_rle_kernel (scan.cu) restated in integers: per run (row r,
    columns [x0, x1)) area += x1-x0, m01 += (x0+x1-1)(x1-x0)/2, m02 by the cubic formula, m10/m20/m11
    from r -- summed over the encoded runs they equal the per-pixel sums and bboxes of the field."""
    H, W = 96, 200
    lab = np.ascontiguousarray(make_field(8, H, W, 12, 5.0, 12.0)[1][None])
    lib = _lib.load()
    sw = 2 * H * W + H + 4
    slots = np.zeros((1, sw), np.uint32); words = np.zeros(1, np.uint32)
    assert lib.cia_rle_encode_fields(lab.ctypes.data, 1, H, W, slots.ctypes.data, sw, words.ctypes.data, None, 1) == 0
    slot, r0 = slots[0].astype(np.int64), (H + 2) & ~1
    acc = {}
    for r in range(H):
        lo, hi = int(slot[r]), int(slot[r + 1])
        for j in range(lo, hi):
            x0, l = int(slot[r0 + 2 * j]), int(slot[r0 + 2 * j + 1])
            x1 = int(slot[r0 + 2 * j + 2]) if j + 1 < hi else W
            if l == 0:
                continue
            a = acc.setdefault(l, dict(area=0, m10=0, m01=0, m20=0, m02=0, m11=0, minr=H, minc=W, maxr=0, maxc=0))
            cnt = x1 - x0
            sc = (x0 + x1 - 1) * cnt // 2
            sc2 = ((x1 - 1) * x1 * (2 * x1 - 1) - (x0 - 1) * x0 * (2 * x0 - 1)) // 6
            a["area"] += cnt; a["m10"] += r * cnt; a["m01"] += sc
            a["m20"] += r * r * cnt; a["m02"] += sc2; a["m11"] += r * sc
            a["minr"] = min(a["minr"], r); a["maxr"] = max(a["maxr"], r + 1)
            a["minc"] = min(a["minc"], x0); a["maxc"] = max(a["maxc"], x1)
    present = [int(v) for v in np.unique(lab) if v != 0]
    assert sorted(acc) == present and len(present) >= 5
    for l in present:
        rr, cc = np.nonzero(lab[0] == l)
        rr, cc = rr.astype(np.int64), cc.astype(np.int64)
        want = dict(area=len(rr), m10=rr.sum(), m01=cc.sum(), m20=(rr * rr).sum(), m02=(cc * cc).sum(),
                    m11=(rr * cc).sum(), minr=rr.min(), minc=cc.min(), maxr=rr.max() + 1, maxc=cc.max() + 1)
        assert {k: int(v) for k, v in want.items()} == acc[l], l


_DIGEST_SNIPPET = """
import hashlib, sys
import numpy as np
sys.path.insert(0, {root!r})
from cell_image_analysis_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(11)
h = hashlib.sha256()
for H, W in ((3, 31), (5, 32), (4, 33), (7, 64), (6, 97), (9, 1030), (64, 2048)):
    # runs of random lengths, boundaries on and around the 8- and 32-label steps, repeated labels
    lens = rng.integers(1, 70, size=H * W)
    vals = rng.integers(0, 5, size=H * W, dtype=np.int32) * rng.integers(1, 2**30, size=H * W, dtype=np.int32)
    lab = np.ascontiguousarray(np.repeat(vals, lens)[:H * W].reshape(1, H, W))
    sw = 2 * H * W + H + 4
    slots = np.zeros((1, sw), np.uint32); words = np.zeros(1, np.uint32)
    assert lib.cia_rle_encode_fields(lab.ctypes.data, 1, H, W, slots.ctypes.data, sw, words.ctypes.data, None, 1) == 0
    h.update(slots[0, :words[0]].tobytes())
print(h.hexdigest())
"""


def test_rle_encoder_avx2_and_scalar_paths_agree_cpu():
    """The streaming AVX2 encoder and the portable run-by-run encoder emit identical slots."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = _DIGEST_SNIPPET.format(root=root)
    outs = []
    for scalar in (False, True):
        env = dict(os.environ)
        env.pop("CIA_HOST_RLE_SCALAR", None)
        if scalar:
            env["CIA_HOST_RLE_SCALAR"] = "1"
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert len(outs[0]) == 64 and outs[0] == outs[1]


def _pack(lib, labs, imgs, label_cap, cap_px, threads=2):
    F, H, W = labs.shape
    sw = int(lib.cia_rle_slot_words(H, W))
    slots = np.zeros((F, sw), np.uint32)
    words = np.zeros(F, np.uint32)
    patches = np.full((F, cap_px), 0xABCD, np.uint16)
    px = np.zeros(F, np.uint32)
    rc = lib.cia_rle_encode_pack_fields(labs.ctypes.data, imgs.ctypes.data, F, H, W, slots.ctypes.data, sw,
                                        words.ctypes.data, None, label_cap, patches.ctypes.data, cap_px,
                                        px.ctypes.data, threads)
    return rc, slots, words, patches, px


@pytest.mark.parametrize("shape", [(64, 64), (37, 101), (128, 1030)])
def test_patch_pack_cpu(shape):
    """The image's patch transport (host side): bbox rectangles of labels 1..label_cap, label ascending, rows
    contiguous -- against scipy.ndimage.find_objects (what regionprops' bbox is, det:67); the runs are the same
    words cia_rle_encode_fields writes."""
    import scipy.ndimage as ndi
    lib = _lib.load()
    H, W = shape
    labs = _fields(H, W)[:3].copy()
    labs[2] %= 40                                                        # stripes: 40 labels whose bboxes span the field
    labs[1, 3:9, 5:20] = -7                                              # negative labels are background
    rng = np.random.default_rng(1)
    imgs = rng.integers(0, 65536, labs.shape, dtype=np.uint16)
    cap = int(labs.max())
    rc, slots, words, patches, px = _pack(lib, labs, imgs, cap, 40 * H * W)
    assert rc == 0
    ref_rc, slots2, words2, _, _ = rc, *_pack(lib, labs, imgs, cap, 40 * H * W, threads=1)[1:3], None, None
    assert np.array_equal(slots, slots2) and np.array_equal(words, words2)
    for f in range(len(labs)):
        want = [imgs[f][sl].ravel() for sl in ndi.find_objects(np.where(labs[f] > 0, labs[f], 0)) if sl is not None]
        want = np.concatenate(want) if want else np.zeros(0, np.uint16)
        assert int(px[f]) == len(want)
        assert np.array_equal(patches[f, :len(want)], want)
        assert (patches[f, len(want):] == 0xABCD).all()                  # nothing written past the used pixels
    # a label cap below the largest label: the labels above it are left out (the device reports them)
    rc, _, _, patches3, px3 = _pack(lib, labs[:1], imgs[:1], 5, 40 * H * W)
    want = np.concatenate([imgs[0][sl].ravel() for sl in ndi.find_objects(np.where((labs[0] > 0) & (labs[0] <= 5), labs[0], 0))
                           if sl is not None])
    assert rc == 0 and int(px3[0]) == len(want) and np.array_equal(patches3[0, :len(want)], want)
    # rectangles that do not fit the slot: flagged per field, the runs are still there
    rc, _, words4, _, px4 = _pack(lib, labs, imgs, cap, 16)
    assert rc == 0 and (words4 > 0).all() and px4[2] == 0xFFFFFFFF and px4[1] == 0      # field 1 has no positive label


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 64), (37, 101), (128, 1030), (512, 2048)])
def test_rle_expand_gpu(shape):
    import torch
    from cell_image_analysis_b200.screening import Engine
    eng = Engine()
    H, W = shape
    labs = torch.from_numpy(_fields(H, W))
    F = labs.shape[0]
    sw = 2 * H * W + H + 4
    h_slots = torch.zeros((F, sw), dtype=torch.int32).pin_memory()
    d_slots = torch.zeros((F, sw), dtype=torch.int32, device="cuda")
    words = np.zeros(F, np.uint32)
    assert eng.rle_encode(labs, h_slots, words, 2)
    out = torch.full((F, H, W), -7, dtype=torch.int32, device="cuda")
    eng.rle_upload_expand(h_slots, words, d_slots, out)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), labs)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 64), (37, 101), (128, 1030), (512, 2048)])
def test_label_scan_from_runs_equals_dense_scan_gpu(shape):
    """cia_label_scan_rle (closed-form sums over runs) builds the same region table, byte for
    byte, as cia_label_scan over the dense field -- including labels outside [0, max_label]
    (ignored + CIA_E_LABEL raised) and fields with one run per 2 pixels."""
    import torch
    from cell_image_analysis_b200.screening import Engine
    eng = Engine()
    H, W = shape
    labs_np = _fields(H, W)
    labs_np[4] = labs_np[0][::-1].copy()                      # replace the 2^31-1 field: keep labels in range here
    labs_np[3] %= 997
    labs = torch.from_numpy(np.ascontiguousarray(labs_np))
    F = labs.shape[0]
    max_label = int(labs.max())
    sw = 2 * H * W + H + 4
    h_slots = torch.zeros((F, sw), dtype=torch.int32).pin_memory()
    d_slots = torch.zeros((F, sw), dtype=torch.int32, device="cuda")
    words = np.zeros(F, np.uint32)
    assert eng.rle_encode(labs, h_slots, words, 2)
    eng.rle_upload(h_slots, words, d_slots)
    dense = eng.label_scan(labs.cuda(), max_label)
    runs = eng.label_scan_rle(d_slots, H, W, max_label)
    torch.cuda.synchronize()
    eng.check_status()
    assert int((dense != 0).sum()) > 0
    assert torch.equal(dense, runs)
    # out-of-range labels: same table as the dense scan, and the status word is raised
    dense2 = eng.label_scan(labs.cuda(), max_label // 2)
    with pytest.raises(_lib.CiaError):
        eng.check_status()
    runs2 = eng.label_scan_rle(d_slots, H, W, max_label // 2)
    torch.cuda.synchronize()
    with pytest.raises(_lib.CiaError):
        eng.check_status()
    assert torch.equal(dense2, runs2)
    eng.close()


@pytest.mark.gpu
def test_run_host_transport_variants_agree(model_dir):
    """runs scanned directly, runs expanded first, and a mixed raw/run-length pass: same results."""
    import torch
    from cell_image_analysis_b200.artifacts import load_model_dir
    from cell_image_analysis_b200.batch import BatchScreen
    from cell_image_analysis_b200.screening import Engine
    from cell_image_analysis_b200.synth import make_fields
    eng = Engine()
    eng.load_artifacts(load_model_dir(model_dir))
    fields = make_fields(range(6), "tiny")
    g = torch.from_numpy(np.stack([f[0] for f in fields]).view(np.int16)).pin_memory()
    lab = torch.from_numpy(np.stack([f[1] for f in fields])).pin_memory()
    H, W = lab.shape[1:]
    res = []
    for kw in (dict(label_transport="raw"), dict(label_transport="rle"), dict(label_transport="rle", image_transport="patches"),
               dict(label_transport="rle", scan_runs=False),
               dict(label_transport="rle", rle_fraction=0.5),
               dict(label_transport="rle", rle_fraction=0.5, image_transport="patches")):
        bs = BatchScreen(eng, H, W, int(lab.max()), chunk_fields=1, **kw)
        bs.run_host(g, lab, 6)
        bs.sync()
        eng.check_status()
        res.append(bs.collect_host())
        h2d = bs.host_bytes_per_pass(6)[0]
        if kw == dict(label_transport="rle", image_transport="patches"):    # bbox rectangles + runs only
            assert h2d < 6 * H * W * 2
        if kw.get("rle_fraction") == 0.5 and "image_transport" not in kw:      # three chunks raw, three as runs
            assert 3 * H * W * 6 + 3 * H * W * 2 < h2d < 3 * H * W * 6 + 3 * H * W * 3
            assert bs.last_rle_share == 0.5
    assert res[0]["n_cells"] > 0
    for r in res[1:]:
        for k in ("cells", "mse", "mae", "dec_cons", "dec_mod", "pred_cons", "pred_mod", "field_counts"):
            assert np.array_equal(res[0][k], r[k]), k


@pytest.mark.gpu
def test_run_host_rle_equals_raw(golden_config1, model_dir):
    """The whole host pass gives identical cells and scores with either label transport."""
    import torch
    from cell_image_analysis_b200.artifacts import load_model_dir
    from cell_image_analysis_b200.batch import BatchScreen
    from cell_image_analysis_b200.screening import Engine
    from cell_image_analysis_b200.synth import make_fields
    eng = Engine()
    eng.load_artifacts(load_model_dir(model_dir))
    fields = make_fields(range(4), "config1")
    g = torch.from_numpy(np.stack([f[0] for f in fields]).view(np.int16)).pin_memory()
    lab = torch.from_numpy(np.stack([f[1] for f in fields])).pin_memory()
    H, W = lab.shape[1:]
    res = {}
    for mode in ("raw", "rle"):
        bs = BatchScreen(eng, H, W, int(lab.max()), chunk_fields=2, label_transport=mode)
        bs.run_host(g, lab, 4)
        bs.sync()
        res[mode] = bs.collect_host()
        h2d = bs.host_bytes_per_pass(4)[0]
        assert (h2d < 4 * H * W * 3) == (mode == "rle")
    a, b = res["raw"], res["rle"]
    assert a["n_cells"] == b["n_cells"] > 0
    for k in ("cells", "mse", "mae", "dec_cons", "dec_mod", "pred_cons", "pred_mod", "field_counts"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_run_host_incompressible_labels_fall_back_to_raw(model_dir):
    """A chunk whose labels do not fit the run-length slots is copied raw; results are unchanged."""
    import torch
    from cell_image_analysis_b200.artifacts import load_model_dir
    from cell_image_analysis_b200.batch import BatchScreen
    from cell_image_analysis_b200.screening import Engine
    eng = Engine()
    eng.load_artifacts(load_model_dir(model_dir))
    H = W = 128
    rng = np.random.default_rng(3)
    cells = make_field(11, H, W, 9, 9.0, 14.0)
    noise = rng.integers(1, 400, (H, W), dtype=np.int32)             # one run per pixel: does not compress
    g = torch.from_numpy(np.stack([cells[0], cells[0]]).view(np.int16)).pin_memory()
    lab = torch.from_numpy(np.stack([cells[1], noise])).pin_memory()
    res = {}
    for mode in ("raw", "rle"):
        bs = BatchScreen(eng, H, W, 400, chunk_fields=1, label_transport=mode)
        bs.run_host(g, lab, 2)
        bs.sync()
        res[mode] = bs.collect_host()
        if mode == "rle":      # field 0 went as runs, field 1 raw
            assert 2 * H * W * 2 + 4 * H * W < bs.host_bytes_per_pass(2)[0] < 2 * H * W * 6
    a, b = res["raw"], res["rle"]
    assert a["n_cells"] == b["n_cells"]
    for k in ("cells", "mse", "dec_cons", "pred_mod", "field_counts"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_run_host_patch_transport_falls_back_when_rectangles_do_not_fit(model_dir):
    """Two long diagonal regions: their bbox rectangles cover the whole field twice, more than a patch slot
    (half a field) holds, so that chunk's images are copied densely; results equal the dense pass."""
    import torch
    from cell_image_analysis_b200.artifacts import load_model_dir
    from cell_image_analysis_b200.batch import BatchScreen
    from cell_image_analysis_b200.screening import Engine
    from cell_image_analysis_b200.synth import make_fields
    eng = Engine()
    eng.load_artifacts(load_model_dir(model_dir))
    fields = make_fields(range(4), "tiny")
    g = np.stack([f[0] for f in fields])
    lab = np.stack([f[1] for f in fields]).copy()
    H, W = lab.shape[1:]
    top = int(lab.max())
    for k in range(min(H, W)):                       # two one-pixel diagonals with full-field bboxes in field 1
        lab[1, k, k] = top + 1
        lab[1, k, W - 1 - k] = top + 2
    gp = torch.from_numpy(g.view(np.int16)).pin_memory()
    lp = torch.from_numpy(lab).pin_memory()
    res = {}
    for mode in ("dense", "patches"):
        bs = BatchScreen(eng, H, W, top + 2, chunk_fields=1, label_transport="rle", image_transport=mode)
        bs.run_host(gp, lp, 4)
        bs.sync()
        res[mode] = (bs.collect_host(), bs.host_bytes_per_pass(4)[0])
    (a, ha), (b, hb) = res["dense"], res["patches"]
    assert a["n_cells"] == b["n_cells"] > 0
    for k in ("cells", "mse", "mae", "dec_cons", "dec_mod", "pred_cons", "pred_mod", "field_counts"):
        assert np.array_equal(a[k], b[k]), k
    assert 2 * H * W < hb < ha                       # one field went densely, three as rectangles
