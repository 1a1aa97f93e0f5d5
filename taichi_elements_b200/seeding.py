"""Host-side helpers of the setup path (seeding shapes that have no device kernel).
Same counter-based generator as libmpm_b200."""
import math

import numpy as np


def _splitmix64(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over='ignore'):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def rand01(seed, ids, draw):
    ids = np.asarray(ids, np.uint64)
    z = _splitmix64(np.uint64(seed) ^ _splitmix64((ids << np.uint64(16)) | np.uint64(draw)))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def polygon_points(seed, id0, n, sides, angle):
    """random_point_in_unit_polygon (reference engine/mpm_solver.py:918-931): uniform points of the
    unit regular polygon with `sides` sides, first vertex direction `angle`, by rejection."""
    ids = np.arange(id0, id0 + n, dtype=np.uint64)
    out = np.zeros((n, 2), np.float32)
    todo = np.ones(n, bool)
    central = 2 * math.pi / sides
    phi = central / 2
    for t in range(1000):
        if not todo.any():
            break
        sel = ids[todo]
        px = rand01(seed, sel, 2 * t) * np.float32(2) - np.float32(1)
        py = rand01(seed, sel, 2 * t + 1) * np.float32(2) - np.float32(1)
        theta = np.mod(np.arctan2(py, px) - angle, central)
        dist = np.sqrt(px * px + py * py)
        ok = dist < np.cos(phi) / np.cos(phi - theta)
        idx = np.nonzero(todo)[0]
        out[idx[ok], 0], out[idx[ok], 1] = px[ok], py[ok]
        todo[idx[ok]] = False
    return out
