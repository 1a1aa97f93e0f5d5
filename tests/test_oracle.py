"""The oracle itself: analytic known-answer tests, invariants and the
NumPy <-> C cross-check (the reference holds no golden vectors for this path,
SURVEY.md section 4; these are what pins the restatement)."""
import numpy as np
import pytest

from oracle.mpm_oracle import (MATERIAL_ELASTIC, MATERIAL_SAND, MATERIAL_SNOW, MATERIAL_STATIONARY,
                               MATERIAL_WATER, OracleMPM, svd2d, svd3d)
from scenes import mixed_scene, rel_err


def test_constants_match_reference_formulas():
    o = OracleMPM((128, 128))
    assert o.dx == 1 / 128 and o.default_dt == 2e-2 / 128
    assert abs(o.mu_0 - 416666.6666) < 1e-3 and abs(o.lambda_0 - 277777.7777) < 1e-3
    assert abs(o.alpha - 0.5035998) < 1e-6
    assert o.offset == (-2048, -2048) and o.leaf_block_size == 16
    o3 = OracleMPM((4096, 4096, 4096), unbounded=True)
    assert o3.grid_size == 16384 and o3.leaf_block_size == 4    # :143-148


@pytest.mark.parametrize('frame_dt,res,count', [(8e-3, 128, 53), (3e-3, 256, 40), (4e-3, 24, 6), (8e-3, 24, 10),
                                                 (1e-2, 256, 130)])
def test_substep_schedule_quirk(frame_dt, res, count):
    # `while frame_time_left > 0` often runs substeps+1 iterations
    _, n = OracleMPM.substep_schedule(frame_dt, 2e-2 / res)
    assert n == count


def test_svd3_convention_and_reconstruction():
    rng = np.random.default_rng(0)
    F = (rng.normal(size=(500, 3, 3)) + 2 * np.eye(3)).astype(np.float32)
    F[:50, :, 0] *= -1                                          # inverted elements
    U, s, V = svd3d(F)
    assert np.allclose(np.linalg.det(U), 1, atol=1e-5) and np.allclose(np.linalg.det(V), 1, atol=1e-5)
    assert np.all(np.abs(s[:, 0]) >= np.abs(s[:, 1]) - 1e-6) and np.all(np.abs(s[:, 1]) >= np.abs(s[:, 2]) - 1e-6)
    assert np.all(s[:, :2] >= 0)
    assert np.array_equal(np.sign(s[:, 2]), np.sign(np.linalg.det(F.astype(np.float64))))
    rec = np.einsum('nik,nk,njk->nij', U, s, V)
    assert np.abs(rec - F).max() < 5e-6


def test_svd2_known_answers():
    # diagonal, rotation, and a shear with analytic singular values
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], np.float32)
    F = np.stack([np.diag([2.0, 0.5]).astype(np.float32), R, np.array([[1, 1], [0, 1]], np.float32),
                  np.array([[2, 0], [0, -1]], np.float32)])
    U, s, V = svd2d(F)
    assert np.allclose(s[0], [2, 0.5]) and np.allclose(s[1], [1, 1], atol=1e-6)
    gold = (1 + np.sqrt(5)) / 2
    assert np.allclose(s[2], [gold, 1 / gold], atol=1e-6)
    assert np.allclose(s[3], [2, -1])                           # signed: reflection carried by sigma
    rec = np.einsum('nik,nk,njk->nij', U, s, V)
    assert np.abs(rec - F).max() < 1e-6
    assert np.allclose(np.linalg.det(U), 1, atol=1e-6) and np.allclose(np.linalg.det(V), 1, atol=1e-6)


def test_sand_projection_known_answers():
    o = OracleMPM((32, 32, 32))
    # expansion (tr >= 0): sigma -> 1, Jp <- tr
    sig, jp = o.sand_projection(np.array([[1.1, 1.05, 1.0]], np.float32), np.zeros(1, np.float32))
    assert np.allclose(sig, 1) and np.isclose(jp[0], np.log(1.1) + np.log(1.05), atol=1e-6)
    # isotropic compression: eps_hat = 0 -> stays on the cone axis, sigma unchanged
    s0 = np.array([[0.9, 0.9, 0.9]], np.float32)
    sig, jp = o.sand_projection(s0, np.zeros(1, np.float32))
    assert np.allclose(sig, s0, atol=1e-6) and jp[0] == 0
    # shear with compression: delta_gamma formula
    s1 = np.array([[1.0, 0.8, 0.7]], np.float32)
    sig, jp = o.sand_projection(s1, np.zeros(1, np.float32))
    eps = np.log(s1[0].astype(np.float64))
    tr = eps.sum()
    eh = eps - tr / 3
    nrm = np.linalg.norm(eh)
    dg = nrm + (3 * o.lambda_0 + 2 * o.mu_0) / (2 * o.mu_0) * tr * o.alpha
    want = np.exp(eps - max(0, dg) / nrm * eh)
    assert np.allclose(sig[0], want, rtol=1e-5)


@pytest.mark.parametrize('dim', [2, 3])
def test_p2g_conserves_mass_and_momentum(dim):
    o = OracleMPM((32, ) * dim)
    for p, m, vel in mixed_scene(dim, seed=1):
        o.add_particles(p, m, velocity=vel)
    # give the particles non-trivial C and F so the affine/stress terms are active
    rng = np.random.default_rng(2)
    o.C = rng.normal(size=o.C.shape).astype(np.float32)
    o.F = (o.F + 0.05 * rng.normal(size=o.F.shape)).astype(np.float32)
    mass, mom = o.particle_mass().sum(), o.total_momentum()
    o.p2g(o.default_dt)
    assert np.isclose(o.grid_m.sum(dtype=np.float64), mass, rtol=1e-5)
    # sum_i w_ip (x_i - x_p) = 0 kills the affine and stress terms
    scale = np.abs(o._affine).max() * o.dx * o.n_particles
    assert np.abs(o.grid_v.sum(0, dtype=np.float64) - mom).max() <= 1e-5 * max(scale, np.abs(mom).max())


@pytest.mark.parametrize('dim', [2, 3])
def test_free_fall_and_rigid_translation(dim):
    o = OracleMPM((32, ) * dim)
    rng = np.random.default_rng(3)
    p = (rng.random((600, dim)) * 0.1 + 0.45).astype(np.float32)
    vel = [0.3, 0.0, -0.2][:dim]
    o.add_particles(p, MATERIAL_ELASTIC, velocity=vel)
    g = (0.0, -9.8, 0.0)[:dim]
    dt = o.default_dt
    x0 = o.x.copy()
    for _ in range(10):
        o.substep(dt)
    # an unconstrained blob in free flight: every particle has v = v0 + g t
    want = np.array(vel) + np.array(g) * 10 * dt
    assert np.abs(o.v - want).max() < 2e-4
    assert np.abs(o.F - np.eye(dim)).max() < 1e-4 and np.abs(o.C).max() < 2e-3
    # zero gravity: rigid translation is a fixed point of the scheme
    o2 = OracleMPM((32, ) * dim)
    o2.set_gravity(tuple([0.0] * dim))
    o2.add_particles(p, MATERIAL_SNOW, velocity=vel)
    for _ in range(5):
        o2.substep(dt)
    assert np.abs(o2.v - np.array(vel, np.float32)).max() < 1e-5
    assert np.abs(o2.x - (x0 + 5 * np.float32(dt) * np.array(vel, np.float32))).max() < 1e-5


def test_stationary_and_water_quirks():
    o = OracleMPM((32, 32))
    p = (np.random.default_rng(4).random((300, 2)) * 0.1 + 0.4).astype(np.float32)
    o.add_particles(p, MATERIAL_STATIONARY, velocity=[1.0, 0.0])
    o.add_particles(p + np.float32(0.15), MATERIAL_WATER)
    x0, v0 = o.x[:300].copy(), o.v[:300].copy()
    for _ in range(3):
        o.substep(o.default_dt)
    assert np.array_equal(o.x[:300], x0) and np.array_equal(o.v[:300], v0)   # :722
    Fw = o.F[300:]
    assert np.all(Fw[:, 0, 1] == 0) and np.all(Fw[:, 1, 0] == 0) and np.all(Fw[:, 1, 1] == 1)
    assert np.array_equal(Fw[:, 0, 0], o.Jp[300:])                             # :537-542
    s = OracleMPM((32, 32))
    s.add_particles(p, MATERIAL_SAND)
    assert np.all(s.Jp == 0)                                                    # :832-833


def test_colliders_known_answers():
    o = OracleMPM((32, 32, 32))
    I = np.array([[10, 2, 10], [10, 29, 10], [10, 10, 10], [2, 10, 30]], np.int64)
    v = np.array([[1, -1, 1], [1, 1, 1], [1, -1, 1], [-1, 0, 2]], np.float32)
    out = o._bbox(v, I, False)
    assert np.array_equal(out, np.array([[1, 0, 1], [1, 0, 1], [1, -1, 1], [0, 0, 0]], np.float32))
    with pytest.raises(ValueError):
        o.add_surface_collider((0, 0, 0), (0, 1, 0), 0, friction=0.5)         # :657-658
    o.add_surface_collider((0.0, 0.5, 0.0), (0.0, 2.0, 0.0), 1, friction=0.5)
    c = o.colliders[-1]
    assert np.allclose(c.normal, [0, 1, 0])
    I = np.array([[5, 3, 5], [5, 20, 5]], np.int64)                            # below / above the plane
    v = np.array([[3.0, -4.0, 0.0], [3.0, -4.0, 0.0]], np.float32)
    out = o._plane(v, I, c)
    assert np.allclose(out[0], [1.0, 0.0, 0.0])     # slip removes v_n; friction: |v|=3 -> 3 + (-4)(0.5) = 1
    assert np.array_equal(out[1], v[1])


@pytest.mark.parametrize('dim', [2, 3])
def test_c_restatement_matches_numpy_oracle(dim):
    from oracle.c_oracle import COracle
    a, b = OracleMPM((32, ) * dim), COracle((32, ) * dim)
    for o in (a, b):
        o.add_sphere_collider((0.3, ) * dim, 0.1, 1)
        o.add_surface_collider((0.5, 0.25, 0.5)[:dim], (0.2, 1.0, 0.1)[:dim], 2, 0.3)
        for p, m, vel in mixed_scene(dim, seed=3):
            o.add_particles(p, m, velocity=vel)
    dt = a.default_dt
    a.substep(dt)
    b.substep(dt)
    vs = float(np.abs(a.v).max())
    assert rel_err(b.x, a.x, 1.0) < 1e-6 and rel_err(b.v, a.v, vs) < 1e-5 and rel_err(b.F, a.F, 1.0) < 1e-5
    for _ in range(10):
        a.substep(dt)
        b.substep(dt)
    assert rel_err(b.x, a.x, 1.0) < 1e-4 and rel_err(b.v, a.v, vs) < 5e-3 and rel_err(b.F, a.F, 1.0) < 1e-3
    assert np.abs(b.Jp - a.Jp).max() < 1e-3


def test_quantised_split_substep_rounds_at_the_stores():
    """quant=True without use_g2p2g in 3D (ref :101-114, 216-247, 567, 723-724): after every substep x, v and F lie on
    their fixed-point / shared-exponent grids, C stays f32, and x advanced with the ROUNDED velocity."""
    from oracle import quant_oracle as q
    o, ref = OracleMPM((32, ) * 3, quant=True), OracleMPM((32, ) * 3)
    for m in (o, ref):
        for p, mat, vel in mixed_scene(3, n_per=60, seed=4):
            m.add_particles(p, mat, velocity=vel)
    assert np.array_equal(o.x, q.round_x(ref.x))
    dt = o.default_dt
    x0 = o.x.copy()
    o.substep(dt)
    ref.substep(dt)
    assert np.array_equal(q.round_x(o.x), o.x) and np.array_equal(q.round_v(o.v), o.v) and np.array_equal(q.round_F(o.F), o.F)
    mov = o.material != MATERIAL_STATIONARY
    assert np.array_equal(o.x[mov], q.round_x((x0 + np.float32(dt) * o.v).astype(np.float32))[mov])
    assert np.abs(o.C).max() > 0 and not np.array_equal(q.round_F(o.C), o.C)            # C is not quantised
    assert rel_err(o.x, ref.x) < 1e-5 and np.abs(o.F - ref.F).max() < 2e-4                # close to the f32 run
    # 2D keeps f32 storage in this build
    o2 = OracleMPM((32, ) * 2, quant=True)
    o2.add_particles(np.float32([[0.51234567, 0.5]]), MATERIAL_ELASTIC)
    assert o2.x[0, 0] == np.float32(0.51234567)


def test_c_restatement_matches_numpy_oracle_with_quantised_storage():
    """The two restatements round at the same stores (quant=True, split substep, 3D): after a substep they agree to one
    step of each field's grid (a value within round-off of a rounding boundary may land one step apart), and both lie
    exactly on the grids."""
    from oracle import quant_oracle as q
    from oracle.c_oracle import COracle
    a, b = OracleMPM((32, ) * 3, quant=True), COracle((32, ) * 3, quant=True)
    for o in (a, b):
        o.add_surface_collider((0.5, 0.25, 0.5), (0.2, 1.0, 0.1), 2, 0.3)
        for p, m, vel in mixed_scene(3, seed=3):
            o.add_particles(p, m, velocity=vel)
    assert np.array_equal(a.x, b.x) and np.array_equal(a.F, b.F)
    dt = a.default_dt
    for it in range(4):
        a.substep(dt)
        b.substep(dt)
        for o in (a, b):
            assert np.array_equal(q.round_x(o.x), o.x) and np.array_equal(q.round_v(o.v), o.v)
            assert np.array_equal(q.round_F(o.F), o.F)
        if it == 0:
            vmax = np.abs(a.v).max(axis=1, keepdims=True)
            assert np.abs(b.x - a.x).max() <= 2.0 / 2**20 * 1.01                         # one step of the 21-bit grid
            assert (np.abs(b.v - a.v) <= np.maximum(vmax, 1e-3) * 2.0**-17 + 2e-5).all()   # one step of the 19-bit fractions
            assert np.abs(b.F - a.F).max() <= 4.1 / 2**15 * 1.01                         # one step of the 16-bit grid
    vs = float(np.abs(a.v).max())
    assert rel_err(b.x, a.x, 1.0) < 1e-4 and rel_err(b.v, a.v, vs) < 5e-3 and rel_err(b.F, a.F, 1.0) < 1e-3
