"""Oracle for ``regionprops`` (subset) and the quality gates.

Follows improved_detection.py:66-95 (training twin CAE_improved_modeltrain.py:57-88).
skimage is absent here; this restates skimage.measure.regionprops' use of the real
``scipy.ndimage.find_objects``.  Test infrastructure only.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

# literals of improved_detection.py:76-95
BORDER_MARGIN = 10
AREA_MIN, AREA_MAX = 200, 8000
ECC_MAX = 0.95
MEAN_MIN, STD_MIN = 0.5, 0.1


def region_table(labels: np.ndarray):
    """det:67 ``regionprops(labels)``: ascending present labels with half-open bbox and area.

    Returns int64 array [K, 6]: label, minr, minc, maxr, maxc, area.
    """
    objs = ndi.find_objects(labels)
    rows = []
    for i, sl in enumerate(objs):
        if sl is None:
            continue
        lab = i + 1
        mask = labels[sl] == lab
        rows.append((lab, sl[0].start, sl[1].start, sl[0].stop, sl[1].stop, int(mask.sum())))
    return np.array(rows, dtype=np.int64).reshape(-1, 6)


def eccentricity_skimage(mask: np.ndarray) -> float:
    """``prop.eccentricity`` (det:84) the way skimage computes it: float central
    moments about the centroid in bbox-local coordinates, inertia tensor,
    ``eigvalsh``, ``sqrt(1 - l2/l1)``."""
    img = mask.astype(np.float64)
    n = img.sum()
    r = np.arange(mask.shape[0], dtype=np.float64)
    c = np.arange(mask.shape[1], dtype=np.float64)
    cr = (img.sum(1) @ r) / n
    cc = (img.sum(0) @ c) / n
    dr, dc = r - cr, c - cc
    mu20 = (dr ** 2) @ img.sum(1)
    mu02 = (dc ** 2) @ img.sum(0)
    mu11 = dr @ img @ dc
    T = np.array([[mu02, -mu11], [-mu11, mu20]]) / n
    ev = np.clip(np.sort(np.linalg.eigvalsh(T))[::-1], 0, None)
    if ev[0] == 0:
        return 0.0
    return float(np.sqrt(1 - ev[1] / ev[0]))


def eccentricity_closed_form(n, m10, m01, m20, m02, m11) -> float:
    """Same quantity from exact integer raw moments (any origin) with the 2x2
    symmetric eigenvalues in closed form -- the formula the CUDA filter uses.
    n*mu20 = n*M20 - M10^2 etc. are exact integers."""
    n = int(n)
    a20 = int(n) * int(m20) - int(m10) * int(m10)   # n * mu20 (rows)
    a02 = int(n) * int(m02) - int(m01) * int(m01)   # n * mu02 (cols)
    a11 = int(n) * int(m11) - int(m10) * int(m01)
    nn = float(n) * float(n)
    a = float(a02) / nn       # T[0,0] = mu02 / n
    c = float(a20) / nn       # T[1,1] = mu20 / n
    b = -float(a11) / nn
    half_tr = 0.5 * (a + c)
    rad = np.sqrt((0.5 * (a - c)) ** 2 + b * b)
    l1, l2 = half_tr + rad, max(half_tr - rad, 0.0)
    if l1 <= 0:
        return 0.0
    return float(np.sqrt(max(1.0 - l2 / l1, 0.0)))


def convex_hull_image(mask: np.ndarray) -> np.ndarray:
    """``skimage.morphology.convex_hull_image(mask)`` with its defaults (``offset_coordinates=True``,
    ``include_borders=True``) as regionprops' ``image_convex`` uses it: every mask pixel contributes
    the four midpoints of its edges, the convex hull of those points is computed with the REAL Qhull
    (scipy.spatial.ConvexHull, the library skimage itself calls) and every pixel centre inside or on
    the hull is set.  skimage is absent here [R]; pinned by skimage's own unit-test vector
    (tests/test_oracle.py::test_convex_hull_image_skimage_vector)."""
    from scipy.spatial import ConvexHull
    mask = np.asarray(mask, bool)
    out = np.zeros(mask.shape, bool)
    rr, cc = np.nonzero(mask)
    if rr.size == 0:
        return out
    pts = np.stack([rr, cc], 1).astype(np.float64)
    offs = np.array([[-0.5, 0.0], [0.5, 0.0], [0.0, -0.5], [0.0, 0.5]])
    pts = np.unique((pts[:, None, :] + offs[None]).reshape(-1, 2), axis=0)
    hull = ConvexHull(pts)
    gr, gc = np.mgrid[0:mask.shape[0], 0:mask.shape[1]]
    g = np.stack([gr.ravel(), gc.ravel()], 1).astype(np.float64)
    # inside or on every facet: normal . x + offset <= 0 (coordinates are multiples of 0.5: a point ON a
    # facet evaluates to ~1e-16, far inside the tolerance)
    inside = np.all(g @ hull.equations[:, :2].T + hull.equations[:, 2] <= 1e-9, axis=1)
    return inside.reshape(mask.shape)


def solidity(mask: np.ndarray) -> float:
    """``prop.solidity`` (det:106, train:101): area / area_convex."""
    return float(mask.sum()) / float(convex_hull_image(mask).sum())


def raw_moments(labels: np.ndarray, lab: int, bbox):
    minr, minc, maxr, maxc = bbox
    rr, cc = np.nonzero(labels[minr:maxr, minc:maxc] == lab)
    rr = rr.astype(np.int64) + minr
    cc = cc.astype(np.int64) + minc
    return (rr.size, int(rr.sum()), int(cc.sum()), int((rr * rr).sum()),
            int((cc * cc).sum()), int((rr * cc).sum()))


def quality_regions(green: np.ndarray, labels: np.ndarray, ecc_fn=None):
    """det:72-95: iterate regions in ascending label order and apply the gates.

    Returns (kept, table): ``kept`` is a list of dicts (label, bbox, area,
    eccentricity, mean_intensity, std_intensity) for regions that pass; ``table``
    is the full region_table for bit-exact scan checks.
    """
    H, W = labels.shape
    tab = region_table(labels)
    kept = []
    for lab, minr, minc, maxr, maxc, area in tab.tolist():
        if minr < BORDER_MARGIN or minc < BORDER_MARGIN or maxr > H - BORDER_MARGIN \
                or maxc > W - BORDER_MARGIN:                      # det:76
            continue
        if area < AREA_MIN or area > AREA_MAX:                     # det:80
            continue
        mask = labels[minr:maxr, minc:maxc] == lab
        ecc = eccentricity_skimage(mask) if ecc_fn is None else ecc_fn(mask)
        if ecc > ECC_MAX:                                          # det:84
            continue
        cell = green[minr:maxr, minc:maxc]                         # det:88 (unmasked)
        mean, std = float(np.mean(cell)), float(np.std(cell))      # det:91-92
        if mean < MEAN_MIN or std < STD_MIN:                       # det:94
            continue
        kept.append(dict(label=lab, bbox=(minr, minc, maxr, maxc), area=area,
                         eccentricity=ecc, solidity=solidity(mask), mean_intensity=mean, std_intensity=std))
    return kept, tab
