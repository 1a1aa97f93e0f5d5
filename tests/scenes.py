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


def stress_noise_dv(o, dt=None, ulps=4):
    """Velocity change that a `ulps`-ulp perturbation of F (|F| ~ 1) causes in ONE substep through the stiffest
    stress term, 2 mu_0 (F - R) F^T:  dv = dt * (4 inv_dx^2 p_vol / p_mass) * 2 mu_0 eps * (dx / 2)
    = 4 dt inv_dx (mu_0 / rho) eps.  It is an ABSOLUTE floor of any f32 implementation of this algorithm (summation
    order of the F update alone moves F by an ulp), independent of how slowly the particle moves: ~1e-5 m/s at the
    reference's E = 1e6, res 256.  Two correct f32 substeps cannot agree better than this on v."""
    dt = o.default_dt if dt is None else dt
    eps = ulps * 2.0**-23
    return 4.0 * dt * o.inv_dx * (o.mu_0 / o.p_rho) * eps


def state_errors(s, o, dt=None):
    """Per-particle error of the CUDA solver `s` against oracle `o` after ONE substep (insertion order), as
    north_star words it: ||a_p - b_p||_inf relative to ||b_p||_inf of THAT particle.

    Floors / allowances, each only where the own magnitude can vanish:
      x  -- denominator floored by one grid cell (a particle at the origin of an unbounded domain has ||x_p|| ~ 0);
      v  -- the f32 stress-noise increment `stress_noise_dv` is subtracted from the absolute error first (a
            particle at rest next to stiff material legitimately differs by ~1e-5 m/s between two correct f32
            implementations), and the denominator is floored by 1e-3 of the fastest particle;
      F  -- none (||F_p||_inf ~ 1);
      C  -- C acts only through C * dpos (|dpos| <= 1.5 dx) next to v_p in the particle's momentum per unit mass,
            so its error is ||dC_p||_inf * dx against max(||v_p||_inf, ||C_p||_inf * dx, v floor), with the same
            noise allowance (C = 4 inv_dx sum w v (o - fx) carries the grid's velocity noise times 4 inv_dx);
      Jp -- absolute (O(1); exactly 0 for fresh sand)."""
    vmax = max(float(np.abs(o.v).max()), 1e-6)
    vfloor = 1e-3 * vmax
    eta = stress_noise_dv(o, dt)
    n = len(o.x)
    if n == 0:
        return {'x': 0.0, 'v': 0.0, 'F': 0.0, 'C': 0.0, 'Jp': 0.0}
    sv, sC = s.v.to_numpy().astype(np.float64), s.C.to_numpy().astype(np.float64)
    vp = np.abs(o.v).max(axis=1)
    dv = np.maximum(np.abs(sv - o.v).max(axis=1) - eta, 0.0)
    dC = np.maximum(np.abs(sC - o.C).reshape(n, -1).max(axis=1) * o.dx - 4 * eta, 0.0)
    denC = np.maximum(np.maximum(vp, np.abs(o.C).reshape(n, -1).max(axis=1) * o.dx), vfloor)
    return {
        'x': rel_err(s.x.to_numpy(), o.x, o.dx),
        'v': float((dv / np.maximum(vp, vfloor)).max()),
        'F': rel_err(s.F.to_numpy(), o.F),
        'C': float((dC / denC).max()),
        'Jp': float(np.abs(s.Jp.to_numpy() - o.Jp).max()),
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
