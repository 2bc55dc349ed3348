"""Seeded synthetic microscopy fields (SURVEY.md §8d "synthetic field generator").

Pure NumPy, shared by the CUDA path, the tests and the CPU baseline so that both
sides score the *same* bytes.  A field is the pair the reference holds at
improved_detection.py:63-66: ``green`` uint16[H, W] (single-channel case of
det:57-59) and ``labels`` int32[H, W] as StarDist would return them.

* labels: filled ellipses on a jittered grid; a later ellipse never overwrites an
  earlier one (so labels never overlap, but a crowded field clips shapes);
  ids 1..n with ~1 % deleted (absent ids), ~5 % pushed against the 10-px margin.
* image: Poisson(100) background + per-cell amplitude * exp(-r^2), Poisson noise
  again, clipped to 65535.
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_field", "make_fields", "FIELD_CONFIGS", "star_maps_from_ellipses", "star_maps_from_labels", "ellipse_lattice"]

# name -> (H, W, n_cells, a_lo, a_hi, log_uniform)
FIELD_CONFIGS = {
    "config1": (2048, 2048, 520, 9.0, 30.0, False),   # BASELINE.json configs[0]/[1]
    "config3": (4096, 4096, 5000, 8.0, 50.0, True),   # dense, wide bbox distribution
    "tiny": (256, 256, 16, 9.0, 20.0, False),          # unit tests
}


def make_field(seed: int, H: int = 2048, W: int = 2048, n_cells: int = 520,
               a_lo: float = 9.0, a_hi: float = 30.0, log_uniform: bool = False):
    """Return (green uint16[H,W], labels int32[H,W]) for ``seed``."""
    rng = np.random.default_rng(seed)
    g = int(np.ceil(np.sqrt(n_cells)))
    pitch_r, pitch_c = H / g, W / g
    labels = np.zeros((H, W), np.int32)
    signal = np.zeros((H, W), np.float32)

    if log_uniform:
        a = np.exp(rng.uniform(np.log(a_lo), np.log(a_hi), n_cells))
    else:
        a = rng.uniform(a_lo, a_hi, n_cells)
    b = a * rng.uniform(0.5, 1.0, n_cells)
    theta = rng.uniform(0.0, np.pi, n_cells)
    amp = rng.uniform(500.0, 4000.0, n_cells)
    jit = rng.uniform(-3.0, 3.0, (n_cells, 2))
    deleted = rng.random(n_cells) < 0.01
    to_margin = rng.random(n_cells) < 0.05
    slots = rng.permutation(g * g)[:n_cells]

    for i in range(n_cells):
        gr, gc = divmod(int(slots[i]), g)
        cr = (gr + 0.5) * pitch_r + jit[i, 0]
        cc = (gc + 0.5) * pitch_c + jit[i, 1]
        ct, st = np.cos(theta[i]), np.sin(theta[i])
        # half extents of the rotated ellipse's bbox
        er = np.sqrt((a[i] * st) ** 2 + (b[i] * ct) ** 2)
        ec = np.sqrt((a[i] * ct) ** 2 + (b[i] * st) ** 2)
        if to_margin[i]:
            # slide towards the nearest image edge so the bbox enters the 10-px margin
            if gr < g // 2:
                cr = er + rng.uniform(1.0, 9.0)
            else:
                cr = H - 1 - er - rng.uniform(1.0, 9.0)
        if cr - er < 0 or cc - ec < 0 or cr + er > H - 1 or cc + ec > W - 1:
            continue  # bbox would leave the image: rejected (id stays absent)
        ext_r, ext_c = 2.5 * er, 2.5 * ec
        r0, r1 = max(int(cr - ext_r), 0), min(int(cr + ext_r) + 2, H)
        c0, c1 = max(int(cc - ext_c), 0), min(int(cc + ext_c) + 2, W)
        rr = np.arange(r0, r1, dtype=np.float64)[:, None] - cr
        cx = np.arange(c0, c1, dtype=np.float64)[None, :] - cc
        u = (cx * ct + rr * st) / a[i]
        v = (-cx * st + rr * ct) / b[i]
        r2 = u * u + v * v
        signal[r0:r1, c0:c1] += (amp[i] * np.exp(-r2)).astype(np.float32)
        if not deleted[i]:
            win = labels[r0:r1, c0:c1]
            win[(r2 <= 1.0) & (win == 0)] = i + 1

    lam = rng.poisson(100.0, (H, W)).astype(np.float32) + signal
    img = rng.poisson(lam)
    green = np.minimum(img, 65535).astype(np.uint16)
    return green, labels


def make_fields(seeds, config: str = "config1"):
    H, W, n, lo, hi, lu = FIELD_CONFIGS[config]
    return [make_field(int(s), H, W, n, lo, hi, lu) for s in seeds]


# ---- segmentation inputs: what a trained StarDist network emits for a field of ellipses ----
def star_maps_from_ellipses(H, W, grid, cells, n_rays=32):
    """cells: [(cy, cx, a, b, theta)].  prob = 1 - normalized elliptical radius (object probability falling off
    towards the boundary, as StarDist's edt_prob), dist = exact distance to the ellipse boundary along each ray."""
    Hg, Wg = H // grid, W // grid
    prob = np.zeros((Hg, Wg), np.float32)
    dist = np.full((Hg, Wg, n_rays), 1e-3, np.float32)
    phis = np.linspace(0, 2 * np.pi, n_rays, endpoint=False)
    rs, rc = np.sin(phis), np.cos(phis)
    for cy, cx, a, b, th in cells:
        r = int(np.ceil(max(a, b))) + 1
        y0, y1 = max(0, int((cy - r) // grid)), min(Hg, int((cy + r) // grid) + 2)
        x0, x1 = max(0, int((cx - r) // grid)), min(Wg, int((cx + r) // grid) + 2)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dy, dx = yy * grid - cy, xx * grid - cx
        c, s = np.cos(th), np.sin(th)
        u, v = (dx * c + dy * s) / a, (-dx * s + dy * c) / b
        rho = np.sqrt(u * u + v * v)
        inside = rho < 1
        # ray (sin, cos) in (y, x): solve |(p + t d)|_ellipse = 1
        du = (rc[None, None] * c + rs[None, None] * s) / a
        dv = (-rc[None, None] * s + rs[None, None] * c) / b
        A = du * du + dv * dv
        B = 2 * (u[..., None] * du + v[..., None] * dv)
        Cc = (u * u + v * v - 1)[..., None]
        t = (-B + np.sqrt(np.maximum(B * B - 4 * A * Cc, 0))) / (2 * A)
        sub_p = prob[y0:y1, x0:x1]; sub_d = dist[y0:y1, x0:x1]
        sub_p[inside] = (1 - rho[inside]).astype(np.float32)
        sub_d[inside] = np.maximum(t[inside], 1e-3).astype(np.float32)
    return prob, dist


def ellipse_lattice(H, W, n_side, seed):
    """[(cy, cx, a, b, theta)]: one ellipse per cell of an n_side x n_side jittered lattice, not touching."""
    rng = np.random.default_rng(seed)
    pitch = min(H, W) / n_side
    cells = []
    for gy in range(n_side):
        for gx in range(n_side):
            a = rng.uniform(0.15, 0.42) * pitch
            cells.append(((gy + 0.5) * pitch + rng.uniform(-4, 4), (gx + 0.5) * pitch + rng.uniform(-4, 4), a,
                          a * rng.uniform(0.5, 1.0), rng.uniform(0, np.pi)))
    return cells


def star_maps_from_labels(labels, grid=2, n_rays=32, max_steps=256):
    """(prob, dist) a trained StarDist network is trained to emit for a label image (its training targets): prob =
    the Euclidean distance transform normalised per object (stardist ``edt_prob``), dist[k] = the distance from the
    pixel to the object's boundary along ray k (stardist ``star_dist``: unit steps until the label changes, then the
    library's sub-pixel correction), sampled on the grid.  labels -> maps -> instances is the round trip of the
    segmentation tests."""
    from scipy import ndimage
    H, W = labels.shape
    lab = labels[::grid, ::grid]
    Hg, Wg = lab.shape
    edt = ndimage.distance_transform_edt(labels > 0)
    # touching objects: distance to the object's own boundary (objects of the synthetic fields do not touch;
    # the generic form costs one EDT per object, only used when labels touch)
    mx = ndimage.maximum(edt, labels, index=np.arange(0, int(labels.max()) + 1))
    prob = np.where(labels > 0, edt / np.maximum(mx[labels], 1e-9), 0.0)[::grid, ::grid].astype(np.float32)
    dist = np.full((Hg, Wg, n_rays), 1e-3, np.float32)
    ys, xs = np.nonzero(lab)
    if len(ys) == 0:
        return prob, dist
    py, px = ys * grid, xs * grid
    me = labels[py, px]
    phis = np.linspace(0, 2 * np.pi, n_rays, endpoint=False)
    for k in range(n_rays):
        dy, dx = np.float32(np.sin(phis[k])), np.float32(np.cos(phis[k]))
        y = py.astype(np.float32); x = px.astype(np.float32)
        steps = np.zeros(len(py), np.float32)
        alive = np.ones(len(py), bool)
        for _ in range(max_steps):
            y = np.where(alive, y + dy, y); x = np.where(alive, x + dx, x)
            ii = np.rint(y).astype(np.int64); jj = np.rint(x).astype(np.int64)
            inside = (ii >= 0) & (ii < H) & (jj >= 0) & (jj < W)
            same = np.zeros(len(py), bool)
            same[inside] = labels[ii[inside], jj[inside]] == me[inside]
            steps = np.where(alive & same, steps + 1, steps)
            alive &= same
            if not alive.any():
                break
        # stardist's correction: the boundary lies about half a pixel before the first foreign pixel
        t_corr = 1 - 0.5 / max(abs(float(dy)), abs(float(dx)))
        d = (steps + 1) + np.float32(t_corr) - 1
        dist[ys, xs, k] = np.maximum(d * np.float32(np.hypot(dy, dx)), 1e-3).astype(np.float32)
    return prob, dist

