"""Seeded scenes shared by the parity tests: the same particles are injected
into the oracle and into the CUDA solver through add_particles."""
import numpy as np


def mixed_scene(dim, n_per=400, seed=0, res=32, spread=0.12):
    """Five materials in separate blobs with non-trivial v; returns list of
    (positions, material, velocity)."""
    rng = np.random.default_rng(seed)
    out = []
    for m in range(5):
        lo = 0.18 + 0.13 * m
        p = (rng.random((n_per, dim)).astype(np.float32) * np.float32(spread) + np.float32(lo)).astype(np.float32)
        vel = [0.5 - 0.2 * m, -1.0 + 0.3 * m, 0.25][:dim]
        out.append((p, m, vel))
    return out


def build_pair(dim, scene, res=32, colliders=(), gravity=None, unbounded=False, size=1, **kw):
    from oracle.mpm_oracle import OracleMPM
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    o = OracleMPM((res, ) * dim, size=size, unbounded=unbounded, **kw)
    s = MPMSolver((res, ) * dim, size=size, unbounded=unbounded, **kw)
    s.substep_batch = 1
    for kind, args in colliders:
        getattr(o, kind)(*args)
        getattr(s, kind)(*args)
    if gravity is not None:
        o.set_gravity(gravity)
        s.set_gravity(gravity)
    for p, m, vel in scene:
        o.add_particles(p, m, velocity=vel)
        s.add_particles(p, m, velocity=vel)
    return o, s


def rel_err(a, b, scale=0.0):
    """max over particles of ||a_p - b_p||_inf / max(||b_p||_inf, scale): relative to the
    particle's own magnitude, floored by the characteristic scale of the field."""
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    if a.shape[0] == 0:
        return 0.0
    num = np.abs(a - b).max(axis=1)
    den = np.maximum(np.abs(b).max(axis=1), max(scale, 1e-30))
    return float((num / den).max())


def state_errors(s, o):
    """Per-field error of the CUDA solver `s` against oracle `o` (insertion order)."""
    vs = max(float(np.abs(o.v).max()), 1e-6)
    return {
        'x': rel_err(s.x.to_numpy(), o.x, 1.0),
        'v': rel_err(s.v.to_numpy(), o.v, vs),
        'F': rel_err(s.F.to_numpy(), o.F, 1.0),
        'C': rel_err(s.C.to_numpy(), o.C, 4 * o.inv_dx * vs),
        'Jp': float(np.abs(s.Jp.to_numpy() - o.Jp).max()),
    }
