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


def rel_err(a, b, floor=0.0):
    """max over particles of ||a_p - b_p||_inf / max(||b_p||_inf, floor): relative to the particle's OWN
    magnitude (north_star: per-particle relative error); `floor` only guards a division by ~0."""
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    if a.shape[0] == 0:
        return 0.0
    num = np.abs(a - b).max(axis=1)
    den = np.maximum(np.abs(b).max(axis=1), max(floor, 1e-30))
    return float((num / den).max())


def state_errors(s, o):
    """Per-particle relative error of the CUDA solver `s` against oracle `o` (insertion order).

    x, v, F: ||a_p - b_p||_inf / ||b_p||_inf of that particle.  Floors (only where the own magnitude can
    vanish): x -- one grid cell (a particle at the origin of an unbounded domain has ||x_p|| ~ 0; what
    matters is its place in the cell); v -- 1e-3 of the fastest particle (particles at rest carry f32
    round-off noise of the moving grid nodes they touch); F -- none needed (||F_p||_inf ~ 1).
    C: C only acts through C * dpos, |dpos| <= 1.5 dx, on the particle's momentum per unit mass, i.e. next to
    v_p; its error is therefore measured as ||dC_p||_inf * dx against max(||v_p||_inf, ||C_p||_inf * dx,
    v floor) -- for a rigidly translating particle C is pure round-off (~1e-7 |v| / dx) and a ratio of two
    noise terms would be meaningless.  Jp: absolute (it is O(1), and exactly 0 for fresh sand)."""
    vmax = max(float(np.abs(o.v).max()), 1e-6)
    vfloor = 1e-3 * vmax
    sv, sC = s.v.to_numpy(), s.C.to_numpy()
    n = len(o.x)
    dC = np.abs(sC.astype(np.float64) - o.C).reshape(n, -1).max(axis=1) * o.dx if n else np.zeros(0)
    denC = np.maximum(np.maximum(np.abs(o.v).max(axis=1), np.abs(o.C).reshape(n, -1).max(axis=1) * o.dx), vfloor) \
        if n else np.ones(0)
    return {
        'x': rel_err(s.x.to_numpy(), o.x, o.dx),
        'v': rel_err(sv, o.v, vfloor),
        'F': rel_err(s.F.to_numpy(), o.F),
        'C': float((dC / denC).max()) if n else 0.0,
        'Jp': float(np.abs(s.Jp.to_numpy() - o.Jp).max()) if n else 0.0,
    }


def tracking_errors(s, o):
    """Drift of the CUDA solver against the oracle after MANY substeps, measured against the scale of each
    field (max |v| etc.), not per particle: two correct f32 implementations decorrelate at round-off level
    per substep (atomic order, stress amplification of the last bits of F), so a slow particle next to a fast
    one legitimately shows a large error relative to its own small velocity."""
    vs = max(float(np.abs(o.v).max()), 1e-6)
    return {
        'x': rel_err(s.x.to_numpy(), o.x, 1.0),
        'v': rel_err(s.v.to_numpy(), o.v, vs),
        'F': rel_err(s.F.to_numpy(), o.F, 1.0),
        'C': rel_err(s.C.to_numpy(), o.C, 4 * o.inv_dx * vs),
        'Jp': float(np.abs(s.Jp.to_numpy() - o.Jp).max()),
    }
