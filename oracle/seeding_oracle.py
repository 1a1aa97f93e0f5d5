"""Oracle for the setup path: the counter-based seeding generator of
libmpm_b200 and the mesh voxelizer of the reference
(/root/reference/engine/voxelizer.py:46-109, engine/mpm_solver.py:1017-1047),
restated in NumPy / pure Python.  TEST INFRASTRUCTURE ONLY."""
import math

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over='ignore'):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def rand01(seed, ids, draw):
    """24-bit uniform in [0, 1): splitmix64(seed ^ splitmix64(id << 16 | draw)) >> 40."""
    ids = np.asarray(ids, np.uint64)
    z = splitmix64(np.uint64(seed) ^ splitmix64((ids << np.uint64(16)) | np.uint64(draw)))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def seed_cube(seed, id0, n, lower, size):
    """seed (engine/mpm_solver.py:840-850): x_k = lower_k + u * size_k in f32."""
    ids = np.arange(id0, id0 + n, dtype=np.uint64)
    d = len(lower)
    x = np.empty((n, d), np.float32)
    for k in range(d):
        x[:, k] = np.float32(lower[k]) + rand01(seed, ids, k) * np.float32(size[k])
    return x


def seed_ellipsoid(seed, id0, n, center, radius):
    """seed_ellipsoid / random_point_in_unit_sphere (:959-976): rejection sampling."""
    ids = np.arange(id0, id0 + n, dtype=np.uint64)
    d = len(center)
    r = np.zeros((n, d), np.float32)
    todo = np.ones(n, bool)
    for t in range(1000):
        if not todo.any():
            break
        cand = np.stack([rand01(seed, ids[todo], t * d + k) * np.float32(2) - np.float32(1) for k in range(d)], 1)
        nsq = np.zeros(cand.shape[0], np.float32)
        for k in range(d):
            nsq = nsq + cand[:, k] * cand[:, k]
        ok = nsq <= 1
        idx = np.nonzero(todo)[0]
        r[idx] = cand            # the device keeps the last draw until one is accepted
        todo[idx[ok]] = False
    return (np.array(center, np.float32)[None] + r * np.array(radius, np.float32)[None]).astype(np.float32)


def _inside_ccw(p, a, b, c):
    cr = lambda u, v: u[0] * v[1] - u[1] * v[0]
    return cr(a - p, b - p) >= 0 and cr(b - p, c - p) >= 0 and cr(c - p, a - p) >= 0


def voxelize(triangles, res, dx, padding=3):
    """Voxelizer.voxelize_triangles (engine/voxelizer.py:46-109), f64; returns a dense int32
    array of the super-sampled grid (res already includes the super-sampling factor)."""
    vox = np.zeros(res, np.int32)
    inv_dx = 1.0 / dx
    jitter = np.array([-0.057616723909439505, -0.25608986292614977, 0.06716309129743714]) * 1e-8
    tris = np.asarray(triangles, np.float64)
    for t in tris:
        a, b, c = t[0:3] + jitter, t[3:6] + jitter, t[6:9] + jitter
        bmin, bmax = np.minimum(np.minimum(a, b), c), np.maximum(np.maximum(a, b), c)
        p_min = max(padding, int(math.floor(bmin[0] * inv_dx)))
        p_max = min(res[0] - padding, int(math.floor(bmax[0] * inv_dx)) + 1)
        q_min = max(padding, int(math.floor(bmin[1] * inv_dx)))
        q_max = min(res[1] - padding, int(math.floor(bmax[1] * inv_dx)) + 1)
        nrm = np.cross(b - a, c - a)
        nrm = nrm / np.linalg.norm(nrm)
        if abs(nrm[2]) < 1e-10:
            continue
        for p in range(p_min, p_max):
            for q in range(q_min, q_max):
                pos = np.array([(p + 0.5) * dx, (q + 0.5) * dx])
                if _inside_ccw(pos, a[:2], b[:2], c[:2]) or _inside_ccw(pos, a[:2], c[:2], b[:2]):
                    base = np.array([pos[0], pos[1], 0.0])
                    height = int(-nrm.dot(base - a) / nrm[2] * inv_dx + 0.5)
                    height = min(height, res[1] - padding)          # res[1]: the reference's quirk
                    inc = 1 if nrm[2] > 0 else -1
                    if height > padding:
                        vox[p, q, padding:height] += inc
    return vox


def icosphere(center, radius, subdiv=2):
    """Closed triangle mesh (n, 9) float32, outward normals (stand-in for the PLY assets)."""
    t = (1 + 5**0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        nf, cache = [], {}

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = v[i] + v[j]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    V = np.array(v) * radius + np.array(center, np.float64)
    return np.array([np.concatenate([V[a], V[b], V[c]]) for a, b, c in f], np.float32)
