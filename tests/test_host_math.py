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
