"""Host-side helpers of the multi-GPU slab decomposition: where to cut, who gets what.

Pure index arithmetic on positions (no SPH step arithmetic); used by bench.py, the multi-GPU
tests and the CPU protocol test.
"""
import numpy as np


def equal_count_planes(x, world):
    """x planes cutting the particles into `world` slabs of (nearly) equal count.

    Returns world+1 values with -inf / +inf at the ends; plane r is the midpoint between the last
    particle of slab r-1 and the first of slab r in sorted-x order."""
    x = np.asarray(x, dtype=np.float64)
    order = np.sort(x)
    planes = [-np.inf]
    for r in range(1, world):
        k = (x.size * r) // world
        planes.append(float(0.5 * (order[k - 1] + order[k])))
    planes.append(np.inf)
    return np.asarray(planes, dtype=np.float32)


def slab_of(x, planes):
    """Rank owning each x (plane r <= x < plane r+1)."""
    return np.clip(np.searchsorted(np.asarray(planes[1:-1], dtype=np.float64), np.asarray(x, dtype=np.float64), side="right"),
                   0, len(planes) - 2).astype(np.int32)


def cell_planes(planes, min_x, cell, grid_size_x):
    """Cell index of each plane for the current grid, as k_grid_setup snaps it:
    floor((plane - min_x) / cell + 0.5) clamped to [0, grid_size_x]; ends map to 0 and INT_MAX."""
    out = []
    for k, p in enumerate(planes):
        if k == 0:
            out.append(0)
        elif k == len(planes) - 1:
            out.append(0x7FFFFFFF)
        else:
            v = np.floor((np.float32(p) - np.float32(min_x)) / np.float32(cell) + np.float32(0.5))
            out.append(int(max(0, min(grid_size_x, int(v)))))
    return out
