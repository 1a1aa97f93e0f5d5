"""The exact device math (csrc/mpm_math.cuh, compiled for the host by the
test-only harness) against the oracle: SVD and the per-particle P2G update."""
import ctypes

import numpy as np
import pytest

from oracle.mpm_oracle import OracleMPM, svd2d, svd3d
from scenes import mixed_scene


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_device_svd3_matches_lapack(host_math):
    rng = np.random.default_rng(1)
    n = 20000
    Q = np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]
    S = np.exp(rng.normal(scale=0.5, size=(n, 3)))
    F = ((Q * S[:, None, :]) @ np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]).astype(np.float32)
    F[:100] = np.eye(3, dtype=np.float32)
    U, V, sig = np.empty_like(F), np.empty_like(F), np.empty((n, 3), np.float32)
    host_math.host_svd3(P(F), n, P(U), P(sig), P(V))
    rec = np.einsum('nik,nk,njk->nij', U, sig, V)
    assert np.abs(rec - F).max() < 1e-5
    assert np.abs(np.linalg.det(U) - 1).max() < 1e-5 and np.abs(np.linalg.det(V) - 1).max() < 1e-5
    Uo, so, Vo = svd3d(F)
    assert np.abs(sig - so).max() < 1e-5
    assert np.abs(np.einsum('nik,njk->nij', U, V) - np.einsum('nik,njk->nij', Uo, Vo)).max() < 1e-5
    assert np.array_equal(U[:100], np.tile(np.eye(3, dtype=np.float32), (100, 1, 1)))   # F = I is exact


def test_device_svd2_is_the_closed_form(host_math):
    rng = np.random.default_rng(2)
    n = 5000
    F = (rng.normal(size=(n, 2, 2)) + np.eye(2)).astype(np.float32)
    U, V, sig = np.empty_like(F), np.empty_like(F), np.empty((n, 2), np.float32)
    host_math.host_svd2(P(F), n, P(U), P(sig), P(V))
    Uo, so, Vo = svd2d(F)
    assert np.abs(U - Uo).max() < 1e-6 and np.abs(sig - so).max() < 1e-5 and np.abs(V - Vo).max() < 1e-6


@pytest.mark.parametrize('dim', [2, 3])
@pytest.mark.parametrize('plastic,fused', [(True, 0), (False, 0), (True, 1), (True, 3)])
def test_particle_update_matches_oracle(host_math, dim, plastic, fused):
    o = OracleMPM((32, ) * dim, support_plasticity=plastic, use_g2p2g=bool(fused), quant=bool(fused & 2))
    for p, m, vel in mixed_scene(dim, seed=5):
        o.add_particles(p, m, velocity=vel)
    rng = np.random.default_rng(6)
    n = o.n_particles
    o.C = (rng.normal(size=o.C.shape) * 5).astype(np.float32)
    o.F = (o.F + 0.08 * rng.normal(size=o.F.shape)).astype(np.float32)
    o.Jp = (o.Jp + 0.02 * rng.normal(size=n)).astype(np.float32)
    F, C, Jp, mat = o.F.copy(), o.C.copy(), o.Jp.copy(), o.material.copy()
    dt = o.default_dt
    o.p2g(dt, g2p2g=bool(fused))
    consts = np.array([o.dx, o.inv_dx, o.p_vol, o.p_mass, o.mu_0, o.lambda_0, o.alpha,
                       (dim * o.lambda_0 + 2 * o.mu_0) / (2 * o.mu_0), o.water_density, o.inv_dx**2,
                       4 * o.inv_dx, fused], np.float32)
    aff, mass = np.empty_like(F), np.empty(n, np.float32)
    host_math.host_particle_update(dim, P(consts), int(plastic), ctypes.c_float(dt), n, P(mat), P(F), P(C), P(Jp),
                                   P(aff), P(mass), None)
    assert np.abs(F - o.F).max() < 2e-5
    assert np.abs(Jp - o.Jp).max() < 2e-5
    assert np.array_equal(mass, o._mass)
    # affine = stress*scale + mass*C: the stress part carries the f32 cancellation of (F - R)
    scale = max(1.0, float(np.abs(o._affine).max()))
    assert np.abs(aff - o._affine).max() < 2e-4 * scale


@pytest.mark.parametrize('plastic,fused', [(True, 0), (False, 0), (True, 1)])
def test_svd_free_paths_match_oracle(host_math, plastic, fused):
    """Small deformations: SNOW inside its clamp interval and expanding SAND (tr >= 0) avoid the SVD in the 3D
    kernels (csrc/mpm_math.cuh particle_update_fast); particles right at the clamp bounds and at tr ~ 0 may take
    either branch.  Results must equal the oracle's full-SVD arithmetic to round-off in every case."""
    dim = 3
    o = OracleMPM((32, ) * dim, support_plasticity=plastic, use_g2p2g=bool(fused))
    for p, m, vel in mixed_scene(dim, n_per=3000, seed=8):
        o.add_particles(p, m, velocity=vel)
    rng = np.random.default_rng(9)
    n = o.n_particles
    o.C = (rng.normal(size=o.C.shape) * 2).astype(np.float32)
    # strains of 0 .. 3e-2: straddles the snow clamp bounds (-2.5e-2, +4.5e-3)
    amp = (rng.random(n) * 3e-2).astype(np.float32)[:, None, None]
    o.F = (o.F + amp * rng.normal(size=o.F.shape)).astype(np.float32)
    # random rotations on top, so that R is not the identity
    Q = np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]
    Q *= np.sign(np.linalg.det(Q))[:, None, None]
    o.F = (Q @ o.F).astype(np.float32)
    o.Jp = (o.Jp + 0.02 * rng.normal(size=n)).astype(np.float32)          # sand: tr of both signs
    F, C, Jp, mat = o.F.copy(), o.C.copy(), o.Jp.copy(), o.material.copy()
    dt = o.default_dt
    o.p2g(dt, g2p2g=bool(fused))
    consts = np.array([o.dx, o.inv_dx, o.p_vol, o.p_mass, o.mu_0, o.lambda_0, o.alpha,
                       (dim * o.lambda_0 + 2 * o.mu_0) / (2 * o.mu_0), o.water_density, o.inv_dx**2,
                       4 * o.inv_dx, fused], np.float32)
    aff, mass, fast = np.empty_like(F), np.empty(n, np.float32), np.zeros(n, np.int32)
    host_math.host_particle_update(dim, P(consts), int(plastic), ctypes.c_float(dt), n, P(mat), P(F), P(C), P(Jp),
                                   P(aff), P(mass), P(fast))
    fast = fast.astype(bool)
    for m_, lo, hi in ((0, 1.0, 1.0), (1, 0.99, 1.0), (4, 0.99, 1.0), (2, 0.05, 0.6)):
        frac = fast[mat == m_].mean()
        assert lo <= frac <= hi, (m_, frac)              # water/elastic/stationary always, snow only inside the clamp
    if plastic:
        assert 0.2 < fast[mat == 3].mean() < 0.8         # sand: only the expanding particles
    else:
        assert not fast[mat == 3].any()
    assert np.abs(F - o.F).max() < 2e-5
    assert np.abs(Jp - o.Jp).max() < 2e-5
    assert np.array_equal(mass, o._mass)
    scale = max(1.0, float(np.abs(o._affine).max()))
    err = np.abs(aff - o._affine).reshape(n, -1).max(1)
    assert err.max() < 2e-4 * scale, (err.argmax(), mat[err.argmax()], fast[err.argmax()])


def test_quantised_storage_codecs_match_the_restatement(host_math):
    """csrc/mpm_quant.cuh (x: 3 x 21-bit fixed, v: shared-exponent 19-bit fractions, F: 9 x 16-bit fixed; ref
    engine/mpm_solver.py:106-114, 216-247) against oracle/quant_oracle.py: same rounded values, values inside the
    representable range within half a step, saturation outside, idempotent."""
    from oracle import quant_oracle as q
    rng = np.random.default_rng(11)
    n = 20000
    x = (rng.random((n, 3)) * 4.4 - 2.2).astype(np.float32)
    x[:4] = [[0, 0, 0], [1.9999999, -2.0, 2.0], [1e-7, -1e-7, 0.5], [3.0, -3.0, 0.123]]
    v = (rng.normal(size=(n, 3)) * np.exp(rng.normal(size=(n, 1)) * 6)).astype(np.float32)
    v[:3] = [[0, 0, 0], [1.0, 1e-9, -1.0], [-5.5, 0.25, 1e-3]]
    F = (np.eye(3).reshape(1, 9) + rng.normal(size=(n, 9)) * 1.5).astype(np.float32)
    F[0] = np.eye(3).ravel()
    F[1] = [4.05, -4.05, 4.2, -4.2, 100, -100, 0, 1, -1]
    for kind, a, ref, nw in ((0, x, q.round_x(x), 2), (1, v, q.round_v(v), 2), (2, F, q.round_F(F.reshape(n, 3, 3)).reshape(n, 9), 5)):
        out = np.empty_like(a)
        words = np.zeros((n, nw), np.uint32)
        host_math.host_quant_round(kind, n, P(np.ascontiguousarray(a)), P(out), P(words))
        assert np.array_equal(out, ref), kind
        again = np.empty_like(a)
        host_math.host_quant_round(kind, n, P(out), P(again), P(words))
        assert np.array_equal(again, out), kind                       # a stored value is a fixed point
    inside = np.abs(x).max(1) < 1.99
    assert np.abs(q.round_x(x) - x)[inside].max() <= 2.0 / 2**20 / 2 * 1.001
    assert np.abs(q.round_x(x)).max() <= 2.0
    rel = np.abs(q.round_v(v) - v).max(1) / np.abs(v).max(1).clip(1e-30)
    assert rel[np.abs(v).max(1) > 1e-15].max() <= 2.0**-17 * 0.5 * 1.001      # 18 significant bits of the largest component
    assert abs(float(q.round_F(np.array([1.0], np.float32))[0]) - 1.0) < 4.1 / 2**15    # the identity is not exact (as in the reference)
