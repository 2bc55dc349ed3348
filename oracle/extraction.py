"""Oracle for ``extract_quality_cells`` from the label mask on
(improved_detection.py:66-111; training twin CAE_improved_modeltrain.py:57-107).
Test infrastructure only."""
from __future__ import annotations

import numpy as np

from . import clahe, regions, resize


def extract_quality_cells_from_labels(green: np.ndarray, labels: np.ndarray):
    """Returns (cells: list of float64 [64,64], stats: list of dict, kept regions, table)."""
    kept, tab = regions.quality_regions(green, labels)
    cells, stats = [], []
    for k in kept:
        minr, minc, maxr, maxc = k["bbox"]
        cell = green[minr:maxr, minc:maxc]                        # det:88
        eq = clahe.equalize_adapthist(cell, clip_limit=0.02)      # det:98
        cells.append(resize.resize(eq, (64, 64)))                 # det:99
        stats.append({"area": k["area"], "eccentricity": k["eccentricity"], "solidity": k["solidity"],
                      "mean_intensity": k["mean_intensity"],
                      "std_intensity": k["std_intensity"]})       # det:103-109
    return cells, stats, kept, tab
