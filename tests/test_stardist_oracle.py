"""CPU tests of the segmentation oracle (oracle/stardist.py, oracle/stardist_post.c) and of the host logic of
cell_image_analysis_b200/stardist.py (layer ordering, model-folder reading).  No GPU.

The oracle is a restatement ("parity unpinned", see its header); what is checked here are the properties any
faithful implementation has: np.percentile IS the normalisation's order statistic, the point-in-polygon rule on
known cases, exact intersection areas on shapes with closed forms, one instance per well-separated object, labels
in descending probability, idempotence of the suppression."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import stardist as sd   # noqa: E402

CFG = dict(n_channel_in=1, grid=[2, 2], n_rays=32, unet_n_depth=3, unet_n_filter_base=32,
           unet_n_conv_per_depth=2, net_conv_after_unet=128)


def _regular_polygon(cy, cx, r):
    """a candidate whose 32 distances are all r"""
    pts = np.array([[cy, cx]], np.int32)
    return sd.polygons(np.full((1, 32), r, np.float32), pts), pts


def test_normalize_is_numpy_percentile():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 4000, (97, 131)).astype(np.uint16)
    y = sd.normalize(x)
    lo, hi = np.float32(np.percentile(x, 3)), np.float32(np.percentile(x, 99.8))
    assert y.dtype == np.float32
    assert np.array_equal(y, (x.astype(np.float32) - lo) / (hi - lo + np.float32(1e-20)))
    assert abs(np.mean(y < 0) - 0.03) < 0.01 and abs(np.mean(y > 1) - 0.002) < 0.002     # no clipping


def test_layer_plan_of_2d_versatile_fluo():
    plan = sd.layer_plan(CFG)
    assert [p[1:3] for p in plan] == [(1, 32), (32, 32), (32, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128),
                                      (128, 256), (256, 128), (256, 128), (128, 64), (128, 64), (64, 32), (64, 32),
                                      (32, 32), (32, 128), (128, 1), (128, 32)]
    flops = sum(2 * k * k * ci * co for _, ci, co, k in plan)
    assert flops > 0


def test_network_shapes_and_dist_floor():
    w = sd.random_model(CFG, seed=1)
    x = np.random.default_rng(1).uniform(0, 1, (32, 48)).astype(np.float32)
    prob, dist = sd.unet_forward(CFG, w, x)
    assert prob.shape == (16, 24) and dist.shape == (16, 24, 32)
    assert prob.min() > 0 and prob.max() < 1 and dist.min() >= np.float32(1e-3)
    p16, d16 = sd.unet_forward(CFG, w, x, half_activations=True)
    assert np.abs(prob - p16).max() < 1e-2            # the fp16-activation twin stays close to float32


def test_pnpoly_rule():
    L = sd.lib()
    xp = np.array([0.0, 4.0, 4.0, 0.0]); yp = np.array([0.0, 0.0, 4.0, 4.0])
    f = lambda x, y: L.sd_pnpoly(4, sd._p(xp), sd._p(yp), C.c_double(x), C.c_double(y))
    assert f(2, 2) == 1 and f(5, 2) == 0 and f(-1e-9, 2) == 0
    assert f(0, 0) == 3 and f(4, 4) == 3              # vertices
    assert f(0, 2) == 2 and f(2, 4) == 2              # edges count as inside for skimage.draw.polygon
    # the example in skimage.draw.polygon's docstring: polygon(r=[1, 2, 8], c=[1, 7, 4]) on a 10 x 10 image
    r = np.array([1.0, 2.0, 8.0]); c = np.array([1.0, 7.0, 4.0])
    img = np.array([[int(L.sd_pnpoly(3, sd._p(c), sd._p(r), C.c_double(x), C.c_double(y)) != 0) for x in range(10)]
                    for y in range(10)])
    doc = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                    [0, 1, 0, 0, 0, 0, 0, 0, 0, 0],
                    [0, 0, 1, 1, 1, 1, 1, 1, 0, 0],
                    [0, 0, 1, 1, 1, 1, 1, 0, 0, 0],
                    [0, 0, 0, 1, 1, 1, 1, 0, 0, 0],
                    [0, 0, 0, 1, 1, 1, 0, 0, 0, 0],
                    [0, 0, 0, 0, 1, 1, 0, 0, 0, 0],
                    [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
                    [0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
                    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]])
    assert np.array_equal(img, doc)


def test_overlap_closed_forms():
    L = sd.lib()
    # two identical regular 32-gons: overlap 1; disjoint: 0; a 32-gon inside a larger one: 1 (area of the smaller)
    d = np.stack([np.full(32, 10, np.float32), np.full(32, 10, np.float32), np.full(32, 4, np.float32),
                  np.full(32, 10, np.float32)])
    pts = np.array([[50, 50], [50, 50], [52, 51], [50, 90]], np.int32)
    vy, vx, area = sd.polygons(d, pts)
    ov = lambda a, b: L.sd_overlap(sd._p(vy), sd._p(vx), sd._p(pts), sd._p(area), C.c_int(a), C.c_int(b))
    exact = 0.5 * 32 * 100 * np.sin(2 * np.pi / 32)
    assert abs(area[0] - exact) < 1e-3
    assert abs(ov(0, 1) - 1) < 1e-9 and ov(0, 3) == 0.0 and abs(ov(0, 2) - 1) < 1e-9
    assert abs(ov(0, 2) - ov(2, 0)) < 1e-12
    # two unit-offset squares-as-stars are hard to write with 32 rays; instead: shifted copies, monotone in the shift
    prev = 1.0
    for s in (2, 6, 10, 14, 18):
        pts2 = np.array([[50, 50], [50, 50 + s]], np.int32)
        vy2, vx2, a2 = sd.polygons(d[:2], pts2)
        o = L.sd_overlap(sd._p(vy2), sd._p(vx2), sd._p(pts2), sd._p(a2), C.c_int(0), C.c_int(1))
        # lens of two discs of radius ~10 at distance s (the 32-gon is within 0.5 % of the disc)
        lens = 2 * 100 * np.arccos(s / 20) - s / 2 * np.sqrt(400 - s * s)
        assert o < prev and abs(o * a2[0] - lens) < 0.02 * lens + 0.5
        prev = o


def _ellipses(H, W, n_side, seed):
    rng = np.random.default_rng(seed)
    pitch = min(H, W) / n_side
    return [((gy + 0.5) * pitch + rng.uniform(-3, 3), (gx + 0.5) * pitch + rng.uniform(-3, 3),
             (a := rng.uniform(0.2, 0.4) * pitch), a * rng.uniform(0.5, 1.0), rng.uniform(0, np.pi))
            for gy in range(n_side) for gx in range(n_side)]


def test_one_instance_per_object_and_label_order():
    H, W = 384, 448
    cells = _ellipses(H, W, 5, 2)
    prob, dist = sd.star_maps_from_ellipses(H, W, 2, cells)
    labels, det = sd.instances_from_prediction(prob, dist, 2, (H, W), 0.4, 0.3)
    assert labels.dtype == np.int32 and labels.shape == (H, W)
    assert len(det["prob"]) == len(cells) == labels.max()
    assert np.all(np.diff(det["prob"]) <= 0)                     # labels follow descending probability
    yy, xx = np.mgrid[0:H, 0:W]
    for cy, cx, a, b, th in cells:
        c, s = np.cos(th), np.sin(th)
        u, v = ((xx - cx) * c + (yy - cy) * s) / a, (-(xx - cx) * s + (yy - cy) * c) / b
        m = u * u + v * v < 1
        lab = np.bincount(labels[m]).argmax()
        assert lab > 0
        iou = (m & (labels == lab)).sum() / (m | (labels == lab)).sum()
        assert iou > 0.93, iou
    # a better polygon wins the pixels it shares with a worse one
    k = np.argmax(det["prob"])
    assert k == 0 and labels[tuple(det["points"][0])] == 1


def test_suppression_is_idempotent_and_threshold_monotone():
    rng = np.random.default_rng(3)
    prob = rng.uniform(0, 1, (64, 64)).astype(np.float32)
    dist = rng.uniform(2, 8, (64, 64, 32)).astype(np.float32)
    _, d1 = sd.instances_from_prediction(prob, dist, 2, (128, 128), 0.7, 0.3)
    # running the suppression again on the survivors removes nothing
    vy, vx = np.ascontiguousarray(d1["coord"][:, 0]), np.ascontiguousarray(d1["coord"][:, 1])
    pts = np.ascontiguousarray(d1["points"])
    area = np.empty(len(pts)); keep = np.zeros(len(pts), np.uint8)
    dsel = dist.reshape(-1, 32)[(pts[:, 0] // 2) * 64 + pts[:, 1] // 2]
    vy2, vx2, area = sd.polygons(dsel, pts)
    assert np.array_equal(vy2, vy) and np.array_equal(vx2, vx)
    sd.lib().sd_nms(C.c_int(len(pts)), sd._p(vy), sd._p(vx), sd._p(pts), sd._p(area), C.c_double(0.3), sd._p(keep))
    assert keep.all()
    _, d2 = sd.instances_from_prediction(prob, dist, 2, (128, 128), 0.7, 0.6)
    assert len(d2["prob"]) >= len(d1["prob"])                   # a laxer threshold keeps at least as many
    _, d0 = sd.instances_from_prediction(prob, dist, 2, (128, 128), 1.5, 0.3)
    assert len(d0["prob"]) == 0


def test_model_folder_round_trip(tmp_path):
    """config.json / thresholds.json / weights_best.h5 -> the layer list in application order (host logic of
    cell_image_analysis_b200.stardist, no GPU)."""
    from cell_image_analysis_b200 import stardist as prod
    w = sd.random_model(CFG, seed=4)
    folder = tmp_path / "2D_demo"
    sd.write_model_folder(str(folder), CFG, w, prob_thresh=0.47, nms_thresh=0.3)
    got = prod.load_weights_h5(str(folder / "weights_best.h5"))
    assert set(got) == set(w)
    for k in w:
        assert np.array_equal(got[k][0], w[k][0]) and np.array_equal(got[k][1], w[k][1])
    order = sorted(got, key=prod._layer_key)
    assert order == [p[0] for p in sd.layer_plan(CFG)]
    assert prod._layer_key("batch_normalization_3") is None
    rs, rc = sd.ray_tables(32)
    assert np.array_equal(np.sin(prod.ray_angles(32)), rs) and np.array_equal(np.cos(prod.ray_angles(32)), rc)
    with pytest.raises(FileNotFoundError):
        prod.StarDist2D.from_pretrained("2D_versatile_fluo", engine=object())


def test_round_trip_labels_to_star_maps_to_instances():
    """labels -> (prob, dist) as StarDist's training targets define them -> instances: every object comes back once,
    with its shape (the size-independent property of the segmentation's post-processing)."""
    from cell_image_analysis_b200 import synth
    H, W, n, lo, hi, lu = synth.FIELD_CONFIGS["tiny"]
    _, truth = synth.make_field(3, H, W, n, lo, hi, lu)
    prob, dist = synth.star_maps_from_labels(truth, 2)
    labels, det = sd.instances_from_prediction(prob, dist, 2, truth.shape, 0.479071, 0.3)
    ids = np.unique(truth)[1:]
    assert labels.max() == len(ids) == len(det["prob"])
    for i in ids:
        m = truth == i
        k = np.bincount(labels[m]).argmax()
        assert k > 0 and (m & (labels == k)).sum() / (m | (labels == k)).sum() > 0.9



def test_weights_reader_takes_both_keras_layouts(tmp_path):
    """legacy H5 weights (/<layer>/<layer>/kernel:0) and the Keras-3 ``*.weights.h5`` layout (/layers/<layer>/vars/0|1)"""
    from cell_image_analysis_b200 import stardist as prod
    from oracle.h5write import write_h5
    w = sd.random_model(CFG, seed=6)
    tree = {"layers": {name: {"vars": {"0": k, "1": b}} for name, (k, b) in w.items()}}
    path = tmp_path / "model.weights.h5"
    path.write_bytes(write_h5(tree))
    got = prod.load_weights_h5(str(path))
    assert set(got) == set(w)
    assert all(np.array_equal(got[k][0], w[k][0]) and np.array_equal(got[k][1], w[k][1]) for k in w)


def test_c_abi_segmentation_entry_points_reject_a_null_handle():
    """no GPU needed: every cia_seg_* call returns CIA_E_ARG (-2) on a null handle instead of touching the device"""
    from cell_image_analysis_b200 import _lib
    lib = _lib.load()
    assert lib.cia_seg_load(None, None, 0, None, None, None, None, None) == -2
    assert lib.cia_seg_normalize(None, None, 16, 16, 3.0, 99.8, None, None, None) == -2
    assert lib.cia_seg_predict(None, None, 16, 16, None, None, None) == -2
    assert lib.cia_seg_instances(None, None, None, 8, 8, 2, 16, 16, 0.5, 0.3, None, None, None) == -2
    assert lib.cia_seg_details(None, 0, None, None, None, None) == -2
    assert lib.cia_seg_layer_info(None, 0, None) == -2
    assert lib.cia_seg_debug_layer(None, 0, None, None, None, 16, 16, None, None, None, None) == -2
