"""CPU oracle for the segmentation step -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench CPU arms).

Restates ``normalize(seg_channel)`` + ``StarDist2D.predict_instances`` as the reference calls them
(/root/reference/improved_detection.py:44, 62-63; CAE_improved_modeltrain.py:54-55).

PARITY UNPINNED.  csbdeep (normalize, unet_block) and stardist (StarDist2D._build, predict_instances, the
C++ NMS, polygons_to_label) are third-party packages that are neither vendored in /root/reference nor
installable here; the reference pins no versions and holds no golden vectors, and the pretrained
``2D_versatile_fluo`` weights it downloads are not available offline.  What IS the real thing:
``np.percentile`` (NumPy itself) inside ``normalize`` and the float32 convolutions (torch-CPU).  The
network topology, the candidate rule, the greedy NMS and the rendering are restated from the packages'
published sources (csbdeep 0.7-0.8 ``internals/blocks.py::unet_block``, stardist 0.8-0.9
``models/model2d.py``, ``nms.py``, ``geometry/geom2d.py``; scikit-image ``draw/_draw.pyx::_polygon``);
the compiled half lives in oracle/stardist_post.c, whose header states the one known deviation
(exact polygon intersection instead of Clipper's integer-snapped one).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, "stardist_post.c")
_SO = os.path.join(HERE, "libstardist_post.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off: the statements of stardist_post.c evaluated as written."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.sd_overlap.restype = C.c_double
        _lib.sd_pnpoly.restype = C.c_int
        _lib.sd_pnpoly.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


# ---- csbdeep.utils.normalize (det:62) -------------------------------------------------------------
def normalize(x, pmin=3, pmax=99.8, eps=1e-20):
    """percentile-based scaling to float32 without clipping: np.percentile over the whole image, then
    normalize_mi_ma with dtype float32: (x - mi) / (ma - mi + eps)."""
    mi = np.percentile(x, pmin, keepdims=True)
    ma = np.percentile(x, pmax, keepdims=True)
    x = x.astype(np.float32, copy=False)
    mi = mi.astype(np.float32, copy=False)
    ma = ma.astype(np.float32, copy=False)
    return (x - mi) / (ma - mi + np.float32(eps))


# ---- the network (StarDist2D._build over csbdeep unet_block) ------------------------------------------
def layer_plan(cfg: dict):
    """[(name, cin, cout, ksize)] in application order for a StarDist2D config."""
    grid, depth, nconv = int(cfg["grid"][0]), int(cfg["unet_n_depth"]), int(cfg["unet_n_conv_per_depth"])
    base, after, rays = int(cfg["unet_n_filter_base"]), int(cfg["net_conv_after_unet"]), int(cfg["n_rays"])
    plan, cin, k = [], int(cfg.get("n_channel_in", 1)), 0
    pooled = 1
    while pooled < grid:
        for _ in range(nconv):
            plan.append((f"conv2d_{k}" if k else "conv2d", cin, base, 3)); cin = base; k += 1
        pooled *= 2
    skips = []
    for n in range(depth):
        for i in range(nconv):
            plan.append((f"down_level_{n}_no_{i}", cin, base << n, 3)); cin = base << n
        skips.append(cin)
    for i in range(nconv - 1):
        plan.append((f"middle_{i}", cin, base << depth, 3)); cin = base << depth
    plan.append((f"middle_{nconv}", cin, base << max(0, depth - 1), 3)); cin = base << max(0, depth - 1)
    for n in reversed(range(depth)):
        cin += skips[n]
        for i in range(nconv - 1):
            plan.append((f"up_level_{n}_no_{i}", cin, base << n, 3)); cin = base << n
        plan.append((f"up_level_{n}_no_{nconv}", cin, base << max(0, n - 1), 3)); cin = base << max(0, n - 1)
    plan.append(("features", cin, after, 3))
    plan.append(("prob", after, 1, 1))
    plan.append(("dist", after, rays, 1))
    return plan


def random_model(cfg: dict, seed: int = 11, dist_bias: float = 6.0):
    """Synthetic weights {name: (kernel HWIO, bias)}: He-normal convolutions; the heads are scaled so that prob
    spreads over (0, 1) and dist stays in a few pixels' range."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, cin, cout, k in layer_plan(cfg):
        std = np.sqrt(2.0 / (k * k * cin))
        kern = (rng.standard_normal((k, k, cin, cout)) * std).astype(np.float32)
        bias = (rng.standard_normal(cout) * 0.05).astype(np.float32)
        if name == "prob":
            kern *= 1.5
            bias[:] = -1.0
        if name == "dist":
            kern *= 0.8
            bias[:] = dist_bias
        w[name] = (kern, bias)
    return w


def unet_forward(cfg: dict, weights: dict, x: np.ndarray, half_activations: bool = False):
    """x float32 [H, W] (already normalized) -> prob [H/g, W/g], dist [H/g, W/g, n_rays] in float32, as
    StarDist2D.predict returns them (dist = max(1e-3, dist)).  ``half_activations`` rounds the weights and every
    layer output to float16 like the tensor-core path stores them (a diagnostic twin, not the oracle)."""
    import torch
    import torch.nn.functional as F
    grid, depth, nconv = int(cfg["grid"][0]), int(cfg["unet_n_depth"]), int(cfg["unet_n_conv_per_depth"])
    names = [p[0] for p in layer_plan(cfg)]
    it = iter(names)

    def q(t):
        return t.half().float() if half_activations else t

    def conv(t, name, act=True):
        k, b = weights[name]
        kt = torch.from_numpy(np.ascontiguousarray(k.transpose(3, 2, 0, 1)))
        if half_activations and name != names[0]:
            kt = kt.half().float()
        y = F.conv2d(t, kt, torch.from_numpy(b), padding=k.shape[0] // 2)
        return F.relu(y) if act else y

    with torch.no_grad():
        t = torch.from_numpy(np.ascontiguousarray(x, np.float32))[None, None]
        pooled = 1
        while pooled < grid:
            for _ in range(nconv):
                t = q(conv(t, next(it)))
            t = F.max_pool2d(t, 2)
            pooled *= 2
        skips = []
        for n in range(depth):
            for _ in range(nconv):
                t = q(conv(t, next(it)))
            skips.append(t)
            t = F.max_pool2d(t, 2)
        for _ in range(nconv):
            t = q(conv(t, next(it)))
        for n in reversed(range(depth)):
            t = torch.cat([F.interpolate(t, scale_factor=2, mode="nearest"), skips[n]], dim=1)
            for _ in range(nconv):
                t = q(conv(t, next(it)))
        t = q(conv(t, next(it)))                       # features
        prob = torch.sigmoid(conv(t, "prob", act=False))[0, 0].numpy()
        dist = conv(t, "dist", act=False)[0].permute(1, 2, 0).contiguous().numpy()
    return prob, np.maximum(np.float32(1e-3), dist)


# ---- _instances_from_prediction -------------------------------------------------------------------------
def ray_tables(n_rays: int = 32):
    phis = np.linspace(0, 2 * np.pi, n_rays, endpoint=False)      # stardist ray_angles
    return np.sin(phis), np.cos(phis)


def candidates(prob, prob_thresh, grid, b=2):
    """points (y, x) on the image lattice, prob and flat indices of prob > prob_thresh outside the b-border,
    in descending probability (ties: the later pixel first)."""
    mask = prob > np.float32(prob_thresh)
    inner = np.zeros_like(mask)
    inner[b:-b, b:-b] = True
    idx = np.flatnonzero(mask & inner)
    p = prob.ravel()[idx]
    order = np.argsort(p, kind="stable")[::-1]
    idx = idx[order]
    pts = np.stack(np.unravel_index(idx, prob.shape), 1).astype(np.int32) * np.int32(grid)
    return pts, prob.ravel()[idx], idx


def polygons(dist_sel, pts):
    n = len(pts)
    rs, rc = ray_tables(dist_sel.shape[1])
    vy = np.empty((n, 32), np.float32); vx = np.empty((n, 32), np.float32); area = np.empty(n, np.float64)
    d = np.ascontiguousarray(dist_sel, np.float32); pts = np.ascontiguousarray(pts, np.int32)
    lib().sd_polygons(C.c_int(n), _p(d), _p(pts), _p(rs), _p(rc), _p(vy), _p(vx), _p(area))
    return vy, vx, area


def instances_from_prediction(prob, dist, grid, shape, prob_thresh, nms_thresh):
    """-> labels int32 [H, W], details {'points', 'prob', 'coord'} (kept polygons in label order)."""
    assert dist.shape[-1] == 32
    pts, p, idx = candidates(prob, prob_thresh, grid)
    d = dist.reshape(-1, 32)[idx]
    vy, vx, area = polygons(d, pts)
    n = len(pts)
    keep = np.zeros(n, np.uint8)
    pts = np.ascontiguousarray(pts)
    lib().sd_nms(C.c_int(n), _p(vy), _p(vx), _p(pts), _p(area), C.c_double(nms_thresh), _p(keep))
    k = keep.astype(bool)
    vy, vx = np.ascontiguousarray(vy[k]), np.ascontiguousarray(vx[k])
    labels = np.zeros(shape, np.int32)
    lib().sd_render(C.c_int(int(k.sum())), _p(vy), _p(vx), C.c_int(shape[0]), C.c_int(shape[1]), _p(labels))
    return labels, {"points": pts[k], "prob": p[k], "coord": np.stack([vy, vx], 1)}


def predict_instances(cfg, weights, x_normalized, prob_thresh, nms_thresh):
    prob, dist = unet_forward(cfg, weights, x_normalized)
    return instances_from_prediction(prob, dist, int(cfg["grid"][0]), x_normalized.shape, prob_thresh, nms_thresh)


# synthetic (prob, dist) maps from ground-truth ellipses live with the other shared input generators
from cell_image_analysis_b200.synth import star_maps_from_ellipses   # noqa: E402,F401


# ---- a StarDist model folder (config.json, thresholds.json, weights_best.h5) for the loader tests -----------
def write_model_folder(path, cfg, weights, prob_thresh=0.479071, nms_thresh=0.3):
    import json
    from .h5write import write_h5
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    with open(os.path.join(path, "thresholds.json"), "w") as f:
        json.dump({"prob": prob_thresh, "nms": nms_thresh}, f)
    # Keras legacy H5 weights layout: /<layer>/<layer>/kernel:0 and bias:0
    tree = {name: {name: {"kernel:0": k, "bias:0": b}} for name, (k, b) in weights.items()}
    with open(os.path.join(path, "weights_best.h5"), "wb") as f:
        f.write(write_h5(tree))
