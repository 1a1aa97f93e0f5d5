"""CPU oracle for the MLS-MPM substep of taichi_elements (NumPy, float32).

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package may import this
module; it is the checker used by tests/, __graft_entry__.smoke() and the
`cpu_baseline` leg of bench.py.

PARITY UNPINNED: the reference (`/root/reference/engine/mpm_solver.py`) is a
set of Taichi kernels and Taichi is neither installed nor installable in this
image, and the reference's own tests hold no value assertions or golden
vectors.  This file therefore restates the reference
line by line; it is pinned only by analytic known-answer tests and invariants
(tests/test_oracle_*.py) and by cross-checking against the independent C
restatement in oracle/mpm_oracle.c.

Every function cites the reference lines it follows (paths relative to
/root/reference).  All kernel arithmetic is float32, as in the reference
(`ti.init` default_fp=f32); Python floats captured by a Taichi kernel become
f32 constants, which NumPy's weak-scalar promotion reproduces.
"""
import math

import numpy as np

f32 = np.float32

MATERIAL_WATER = 0
MATERIAL_ELASTIC = 1
MATERIAL_SNOW = 2
MATERIAL_SAND = 3
MATERIAL_STATIONARY = 4

SURFACE_STICKY = 0
SURFACE_SLIP = 1
SURFACE_SEPARATE = 2


# ----------------------------------------------------------------------------
# ti.svd restated (Taichi 1.1.0 is not under /root/reference)
# ----------------------------------------------------------------------------
def svd2d(F):
    """2x2 f32 SVD with Taichi's closed form (polar decomposition, then a
    Jacobi rotation of the symmetric factor).  U, V are rotations, sigma is
    signed with sig[0] >= sig[1].  Call sites: engine/mpm_solver.py:428,525."""
    F = np.asarray(F, dtype=f32)
    n = F.shape[0]
    a = F[:, 0, 0] + F[:, 1, 1]
    b = F[:, 1, 0] - F[:, 0, 1]
    with np.errstate(divide='ignore', invalid='ignore'):
        s = f32(1.0) / np.sqrt(a * a + b * b)
    rc, rs = a * s, b * s
    # R = [[rc, -rs], [rs, rc]] ; S = R^T F
    S00 = rc * F[:, 0, 0] + rs * F[:, 1, 0]
    S01 = rc * F[:, 0, 1] + rs * F[:, 1, 1]
    S11 = -rs * F[:, 0, 1] + rc * F[:, 1, 1]
    c = np.ones(n, f32)
    sn = np.zeros(n, f32)
    s1 = S00.copy()
    s2 = S11.copy()
    nz = np.abs(S01) >= f32(1e-5)
    if nz.any():
        tau = f32(0.5) * (S00 - S11)
        w = np.sqrt(tau * tau + S01 * S01)
        with np.errstate(divide='ignore', invalid='ignore'):
            t = np.where(tau > 0, S01 / (tau + w), S01 / (tau - w)).astype(f32)
        cc = f32(1.0) / np.sqrt(t * t + f32(1.0))
        ss = -t * cc
        c = np.where(nz, cc, c).astype(f32)
        sn = np.where(nz, ss, sn).astype(f32)
        c2, s2_ = c * c, sn * sn
        cs2 = f32(2.0) * c * sn * S01
        s1 = np.where(nz, c2 * S00 - cs2 + s2_ * S11, s1).astype(f32)
        s2 = np.where(nz, s2_ * S00 + cs2 + c2 * S11, s2).astype(f32)
    V = np.empty((n, 2, 2), f32)
    swap = s1 < s2
    sig1 = np.where(swap, s2, s1)
    sig2 = np.where(swap, s1, s2)
    V[:, 0, 0] = np.where(swap, -sn, c)
    V[:, 0, 1] = np.where(swap, c, sn)
    V[:, 1, 0] = np.where(swap, -c, -sn)
    V[:, 1, 1] = np.where(swap, -sn, c)
    R = np.empty((n, 2, 2), f32)
    R[:, 0, 0] = rc
    R[:, 0, 1] = -rs
    R[:, 1, 0] = rs
    R[:, 1, 1] = rc
    U = np.einsum('nij,njk->nik', R, V).astype(f32)
    sig = np.stack([sig1, sig2], axis=1).astype(f32)
    return U, sig, V


def svd3d(F):
    """3x3 SVD in the Sifakis/Taichi convention: U, V proper rotations,
    |sigma| sorted descending, sign of det(F) carried by the last sigma.
    The decomposition itself is LAPACK in f64 (any accurate algorithm is within
    the 1e-4 parity budget)."""
    F = np.asarray(F, dtype=f32)
    U, s, Vh = np.linalg.svd(F.astype(np.float64))
    V = np.swapaxes(Vh, 1, 2).copy()
    du = np.linalg.det(U) < 0
    U[du, :, 2] *= -1
    s[du, 2] *= -1
    dv = np.linalg.det(V) < 0
    V[dv, :, 2] *= -1
    s[dv, 2] *= -1
    return U.astype(f32), s.astype(f32), V.astype(f32)


def svd(F):
    return svd2d(F) if F.shape[-1] == 2 else svd3d(F)


def _diag(sig):
    n, d = sig.shape
    out = np.zeros((n, d, d), f32)
    for i in range(d):
        out[:, i, i] = sig[:, i]
    return out


def _mm(a, b):
    return np.einsum('nij,njk->nik', a, b).astype(f32)


def _T(a):
    return np.swapaxes(a, 1, 2)


class Collider:
    """One entry of MPMSolver.grid_postprocess (engine/mpm_solver.py:300-302,
    618-692): kind in {'bbox', 'sphere', 'plane'}."""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


class OracleMPM:
    """Restatement of MPMSolver's state, constants and substep
    (engine/mpm_solver.py:44-310, 344-361, 487-616, 618-735, 748-805)."""

    def __init__(self, res, size=1, padding=3, unbounded=False, dt_scale=1,
                 E_scale=1, water_density=1.0, support_plasticity=True,
                 use_g2p2g=False, v_clamp_g2p2g=True, g2p2g_allowed_cfl=0.9, quant=False):
        self.use_g2p2g = use_g2p2g                                # :57, 69
        self.v_clamp_g2p2g = v_clamp_g2p2g
        self.g2p2g_allowed_cfl = g2p2g_allowed_cfl
        self.quant = quant
        self.F_bound = 4.0                                        # :99
        self.last_time_final_particles = 0                        # :124
        self._prev_grid = None                                    # (sorted linear cell keys, velocities)
        self.dim = len(res)
        assert self.dim in (2, 3)
        self.res = tuple(res)
        self.grid_size = 4096                                    # :74
        self.dx = size / res[0]                                  # :82
        self.inv_dx = 1.0 / self.dx                              # :83
        self.default_dt = 2e-2 * self.dx / size * dt_scale       # :84
        self.p_vol = self.dx ** self.dim                         # :85
        self.p_rho = 1000
        self.p_mass = self.p_vol * self.p_rho                    # :87
        self.water_density = water_density
        self.support_plasticity = support_plasticity
        self.padding = padding
        self.unbounded = unbounded
        if unbounded:                                            # :143-148
            while self.grid_size <= 2 * max(self.res):
                self.grid_size *= 2
        self.offset = tuple(-self.grid_size // 2 for _ in range(self.dim))  # :149
        self.leaf_block_size = 16 if self.dim == 2 else 4        # :155-159
        self.E, self.nu = 1e6 * size * E_scale, 0.2              # :201
        self.mu_0 = self.E / (2 * (1 + self.nu))                 # :203
        self.lambda_0 = self.E * self.nu / ((1 + self.nu) * (1 - 2 * self.nu))
        friction_angle = math.radians(45)                        # :208-210
        sin_phi = math.sin(friction_angle)
        self.alpha = math.sqrt(2 / 3) * 2 * sin_phi / (3 - sin_phi)
        self.gravity = np.zeros(self.dim, f32)
        self.gravity[1] = f32(-9.8)                              # :283,296
        self.t = 0.0
        self.total_substeps = 0
        self.all_time_max_velocity = 0.0
        self.colliders = [Collider('bbox', unbounded=unbounded)]  # :300-302
        d = self.dim
        self.x = np.zeros((0, d), f32)
        self.v = np.zeros((0, d), f32)
        self.F = np.zeros((0, d, d), f32)
        self.C = np.zeros((0, d, d), f32)
        self.Jp = np.zeros((0,), f32)
        self.material = np.zeros((0,), np.int32)
        self.color = np.zeros((0,), np.int32)
        # last grid, for tests
        self.grid_cells = None
        self.grid_m = None
        self.grid_v = None

    def _packed(self):
        """Bit-packed x / v / F storage: quant=True in 3D (:106-114, 216-247); C stays f32 in the split mode (:101-102)."""
        return bool(self.quant and self.dim == 3)

    @property
    def n_particles(self):
        return self.x.shape[0]

    # ---- API pieces needed to drive the substep -------------------------
    def set_gravity(self, g):                                    # :316-319
        assert isinstance(g, (tuple, list)) and len(g) == self.dim
        self.gravity = np.array(g, dtype=f32)

    def add_particles(self, particles, material, color=0xFFFFFF, velocity=None):
        """seed_from_external_array + seed_particle (:823-838, 1081-1104)."""
        p = np.asarray(particles, dtype=f32).reshape(-1, self.dim)
        n, d = p.shape
        vel = np.zeros(d, f32) if velocity is None else np.array(velocity, f32)
        self.x = np.concatenate([self.x, p])
        self.v = np.concatenate([self.v, np.tile(vel, (n, 1))])
        eye = np.tile(np.eye(d, dtype=f32), (n, 1, 1))
        self.F = np.concatenate([self.F, eye])
        self.C = np.concatenate([self.C, np.zeros((n, d, d), f32)])
        jp = f32(0.0) if material == MATERIAL_SAND else f32(1.0)  # :831-835
        self.Jp = np.concatenate([self.Jp, np.full(n, jp, f32)])
        self.material = np.concatenate(
            [self.material, np.full(n, material, np.int32)])
        self.color = np.concatenate([self.color, np.full(n, color, np.int32)])
        if self._packed():                                        # seed_particle stores into the quantised fields (:826-830)
            from .quant_oracle import round_F, round_v, round_x
            self.x, self.v, self.F = round_x(self.x), round_v(self.v), round_F(self.F, self.F_bound)

    def add_sphere_collider(self, center, radius, surface=SURFACE_STICKY):
        self.colliders.append(Collider('sphere', center=list(center),
                                       radius=radius, surface=surface))

    def add_surface_collider(self, point, normal, surface=SURFACE_STICKY,
                             friction=0.0):
        normal_scale = 1.0 / math.sqrt(sum(x ** 2 for x in normal))  # :654
        normal = list(normal_scale * x for x in normal)
        if surface == SURFACE_STICKY and friction != 0:              # :657
            raise ValueError('friction must be 0 on sticky surfaces.')
        self.colliders.append(Collider('plane', point=list(point),
                                       normal=normal, surface=surface,
                                       friction=friction))

    def clear_grid_postprocess(self):
        self.colliders.clear()

    # ---- binning (build_pid, :344-361) -----------------------------------
    def base_index(self, x=None):
        x = self.x if x is None else x
        # int(ti.floor(x * inv_dx - 0.5)): f32 mul, f32 sub, floor
        return np.floor(x * f32(self.inv_dx) - f32(0.5)).astype(np.int64)

    def binning(self):
        """Returns (block (n,d) int64 in [0, grid_size/leaf), base (n,d))."""
        base = self.base_index()
        off = np.array(self.offset, np.int64)
        block = (base - off) // self.leaf_block_size             # rescale_index :360
        return block, base

    @staticmethod
    def unique_rows(a):
        if a.shape[0] == 0:
            return a, np.zeros((0,), np.int64)
        u, counts = np.unique(a, axis=0, return_counts=True)
        return u, counts

    def active_blocks(self):
        """Set of leaf blocks activated by P2G writes (:582-584); SURVEY A-1."""
        base = self.base_index()
        d = self.dim
        offs = np.stack(np.meshgrid(*([np.arange(3)] * d), indexing='ij'),
                        -1).reshape(-1, d)
        cells = (base[:, None, :] + offs[None]).reshape(-1, d)
        off = np.array(self.offset, np.int64)
        blocks = (cells - off) // self.leaf_block_size
        return np.unique(blocks, axis=0)

    # ---- sand_projection (:321-342) --------------------------------------
    def sand_projection(self, sig, Jp):
        d = self.dim
        eps = np.log(np.maximum(np.abs(sig), f32(1e-4))).astype(f32)
        tr = (eps.sum(axis=1, dtype=f32) + Jp).astype(f32)
        eps_hat = (eps - (tr / f32(d))[:, None]).astype(f32)
        eps_hat_norm = (np.sqrt((eps_hat * eps_hat).sum(axis=1, dtype=f32))
                        + f32(1e-20)).astype(f32)
        coef = f32((d * self.lambda_0 + 2 * self.mu_0) / (2 * self.mu_0))
        delta_gamma = (eps_hat_norm + coef * tr * f32(self.alpha)).astype(f32)
        new_Jp = np.where(tr >= 0, tr, f32(0.0)).astype(f32)
        proj = np.exp(eps - (np.maximum(f32(0), delta_gamma) /
                             eps_hat_norm)[:, None] * eps_hat).astype(f32)
        sig_out = np.where((tr >= 0)[:, None], f32(1.0), proj).astype(f32)
        return sig_out, new_Jp

    # ---- helpers ----------------------------------------------------------
    def _stencil(self):
        d = self.dim
        return np.stack(np.meshgrid(*([np.arange(3)] * d), indexing='ij'),
                        -1).reshape(-1, d)

    def _weights(self, fx):
        return [f32(0.5) * (f32(1.5) - fx) ** 2,
                f32(0.75) - (fx - f32(1.0)) ** 2,
                f32(0.5) * (fx - f32(0.5)) ** 2]

    # ---- P2G (:487-584) -----------------------------------------------------
    def p2g(self, dt, g2p2g=False):
        """P2G of the split path (:487-584); g2p2g=True switches to the P2G half of the fused
        kernel (:405-483), whose differences are marked [g2p2g]."""
        dt = f32(dt)
        d = self.dim
        n = self.n_particles
        x, v, C, mat = self.x, self.v, self.C, self.material
        inv_dx = f32(self.inv_dx)
        base = self.base_index()
        fx = (x * inv_dx - base.astype(f32)).astype(f32)          # :503
        w = self._weights(fx)                                     # :505
        I = np.eye(d, dtype=f32)
        F = self.F.copy()                                         # :507
        water = mat == MATERIAL_WATER
        if water.any() and not g2p2g:                             # :508-511 ([g2p2g] keeps the stored F, :414)
            Fw = np.tile(I, (int(water.sum()), 1, 1))
            if self.support_plasticity:
                Fw[:, 0, 0] = self.Jp[water]
            F[water] = Fw
        F = _mm((I[None] + dt * C).astype(f32), F)                # :513
        if g2p2g and self.quant:                                  # [g2p2g] :415-416
            F = np.maximum(f32(-self.F_bound), np.minimum(f32(self.F_bound), F)).astype(f32)
            if self._packed():                                    # self.F[p] = new_F rounds to the 16-bit grid (:416)
                from .quant_oracle import round_F
                F = round_F(F, self.F_bound)
        h = np.ones(n, f32)                                       # :515-521
        if self.support_plasticity:
            with np.errstate(over='ignore'):
                hh = np.exp(f32(10) * (f32(1.0) - self.Jp)).astype(f32)
            h = hh if g2p2g else np.where(~water, hh, h).astype(f32)   # [g2p2g] hardens water too (:419-421)
        h = np.where(mat == MATERIAL_ELASTIC, f32(0.3), h).astype(f32)
        mu = (f32(self.mu_0) * h).astype(f32)
        la = (f32(self.lambda_0) * h).astype(f32)
        mu = np.where(water, f32(0.0), mu).astype(f32)            # :523-524
        U, sig, V = svd(F)                                        # :525
        Jp = self.Jp.copy()
        J = np.ones(n, f32)
        not_sand = mat != MATERIAL_SAND
        snow = mat == MATERIAL_SNOW
        sig_ns = sig.copy()
        for k in range(d):                                        # :527-536
            new_sig = sig[:, k].copy()
            clamped = np.minimum(np.maximum(sig[:, k], f32(1 - 2.5e-2)),
                                 f32(1 + 4.5e-3)).astype(f32)
            new_sig = np.where(snow, clamped, new_sig).astype(f32)
            if self.support_plasticity:
                with np.errstate(divide='ignore', invalid='ignore'):
                    Jp = np.where(not_sand, Jp * (sig[:, k] / new_sig),
                                  Jp).astype(f32)
            sig_ns[:, k] = np.where(not_sand, new_sig, sig[:, k])
            J = np.where(not_sand, J * new_sig, J).astype(f32)
        sig = sig_ns
        if water.any():                                           # :537-542
            Fw = np.tile(I, (int(water.sum()), 1, 1))
            Fw[:, 0, 0] = J[water]
            F[water] = Fw
            if self.support_plasticity and not g2p2g:             # [g2p2g] does not reset Jp (:440-444)
                Jp[water] = J[water]
        if snow.any():                                            # :543-545
            F[snow] = _mm(_mm(U[snow], _diag(sig[snow])), _T(V[snow]))
        stress = np.zeros((n, d, d), f32)
        ns = not_sand
        if ns.any():                                              # :549-551
            R = _mm(U[ns], _T(V[ns]))
            t1 = _mm(((f32(2) * mu[ns])[:, None, None] * (F[ns] - R)).astype(f32),
                     _T(F[ns]))
            t2 = I[None] * (la[ns] * J[ns] * (J[ns] - f32(1)))[:, None, None]
            stress[ns] = (t1 + t2).astype(f32)
        sand = ~ns
        if sand.any() and self.support_plasticity:                # :553-566
            sg, jp_new = self.sand_projection(sig[sand], Jp[sand])
            Jp[sand] = jp_new
            Fs = _mm(_mm(U[sand], _diag(sg)), _T(V[sand]))
            F[sand] = Fs
            logs = np.log(sg).astype(f32)
            log_sum = logs.sum(axis=1, dtype=f32)
            inv = (f32(1) / sg).astype(f32)
            center = (f32(2.0) * f32(self.mu_0) * logs * inv).astype(f32)
            center = (center + f32(self.lambda_0) * log_sum[:, None] * inv
                      ).astype(f32)
            stress[sand] = _mm(_mm(_mm(U[sand], _diag(center)), _T(V[sand])),
                               _T(Fs))
        self.F = F.astype(f32)                                    # :567
        if self._packed() and not g2p2g:                          # the store rounds F to its 16-bit grid (:113-114)
            from .quant_oracle import round_F
            self.F = round_F(self.F, self.F_bound)
        self.Jp = Jp.astype(f32)
        scale = f32(f32(f32(-dt * f32(self.p_vol)) * f32(4)) *
                    f32(self.inv_dx ** 2))                        # :569
        stress = (scale * stress).astype(f32)
        mass = np.full(n, f32(self.p_mass), f32)                  # :571-573
        if not g2p2g:                                             # [g2p2g] has no water_density (:472, 570)
            mass = np.where(water, mass * f32(self.water_density), mass).astype(f32)
        affine = (stress + mass[:, None, None] * C).astype(f32)   # :574
        self._affine, self._mass = affine, mass

        # scatter (:577-584) onto the set of touched cells
        offs = self._stencil()
        k27 = offs.shape[0]
        cells = (base[:, None, :] + offs[None]).reshape(-1, d)
        shift = cells - np.array(self.offset, np.int64)
        gs = self.grid_size
        lin = np.zeros(cells.shape[0], np.int64)
        for a in range(d):
            lin = lin * gs + shift[:, a]
        ulin, inv_idx = np.unique(lin, return_inverse=True)
        ncell = ulin.shape[0]
        ucells = np.empty((ncell, d), np.int64)
        rem = ulin.copy()
        for a in reversed(range(d)):
            ucells[:, a] = rem % gs
            rem //= gs
        ucells += np.array(self.offset, np.int64)
        grid_mv = np.zeros((ncell, d), f32)
        grid_m = np.zeros(ncell, f32)
        mv = (mass[:, None] * v).astype(f32)
        inv_idx = inv_idx.reshape(n, k27)
        for k, o in enumerate(offs):
            dpos = ((o.astype(f32)[None] - fx) * f32(self.dx)).astype(f32)
            weight = np.ones(n, f32)
            for a in range(d):
                weight = (weight * w[o[a]][:, a]).astype(f32)
            contrib = (weight[:, None] *
                       (mv + np.einsum('nij,nj->ni', affine, dpos).astype(f32))
                       ).astype(f32)
            np.add.at(grid_mv, inv_idx[:, k], contrib)
            np.add.at(grid_m, inv_idx[:, k], (weight * mass).astype(f32))
        self.grid_cells = ucells
        self.grid_m = grid_m
        self.grid_v = grid_mv
        self._inv_idx = inv_idx
        self._fx = fx
        self._w = w

    # ---- grid ops (:586-616, 618-687) ---------------------------------------
    def grid_op(self, dt, t=0.0):
        dt = f32(dt)
        d = self.dim
        m, v, I = self.grid_m, self.grid_v, self.grid_cells
        pos = m > 0                                               # :591-593
        with np.errstate(divide='ignore', invalid='ignore'):
            vn = ((f32(1) / m)[:, None] * v).astype(f32)
        vn = (vn + dt * self.gravity[None]).astype(f32)
        v = np.where(pos[:, None], vn, v).astype(f32)
        if self.g2p2g_allowed_cfl > 0 and self.use_g2p2g and self.v_clamp_g2p2g:   # :589, 596-598
            v_allowed = f32(f32(self.dx * self.g2p2g_allowed_cfl) / dt)
            v = np.minimum(np.maximum(v, -v_allowed), v_allowed).astype(f32)
        for c in self.colliders:
            if c.kind == 'bbox':
                v = self._bbox(v, I, c.unbounded)
            elif c.kind == 'sphere':
                v = self._sphere(v, I, c)
            else:
                v = self._plane(v, I, c)
        self.grid_v = v

    def _bbox(self, v, I, unbounded):                             # :600-616
        v = v.copy()
        for a in range(self.dim):
            if unbounded:
                lo = -self.grid_size // 2 + self.padding
                hi = self.grid_size // 2 - self.padding
            else:
                lo = self.padding
                hi = self.res[a] - self.padding
            m1 = (I[:, a] < lo) & (v[:, a] < 0)
            v[m1, a] = 0
            m2 = (I[:, a] >= hi) & (v[:, a] > 0)
            v[m2, a] = 0
        return v

    def _sphere(self, v, I, c):                                   # :621-640
        offset = (I.astype(f32) * f32(self.dx) - np.array(c.center, f32)[None]
                  ).astype(f32)
        nsq = (offset * offset).sum(axis=1, dtype=f32)
        inside = nsq < f32(c.radius * c.radius)
        if c.surface == SURFACE_STICKY:
            out = np.zeros_like(v)
        else:
            # Vector.normalized(eps) = v * (1 / (norm + eps))  [Taichi, EXT]
            normal = (offset * (f32(1) / (np.sqrt(nsq) + f32(1e-5)))[:, None]
                      ).astype(f32)
            nc = (normal * v).sum(axis=1, dtype=f32)
            if c.surface == SURFACE_SLIP:
                out = (v - normal * nc[:, None]).astype(f32)
            else:
                out = (v - normal * np.minimum(nc, f32(0))[:, None]).astype(f32)
        return np.where(inside[:, None], out, v).astype(f32)

    def _plane(self, v, I, c):                                    # :660-685
        n = np.array(c.normal, f32)
        offset = (I.astype(f32) * f32(self.dx) - np.array(c.point, f32)[None]
                  ).astype(f32)
        inside = (offset * n[None]).sum(axis=1, dtype=f32) < 0
        if c.surface == SURFACE_STICKY:
            out = np.zeros_like(v)
        else:
            nc = (v * n[None]).sum(axis=1, dtype=f32)
            if c.surface == SURFACE_SLIP:
                out = (v - n[None] * nc[:, None]).astype(f32)
            else:
                out = (v - n[None] * np.minimum(nc, f32(0))[:, None]).astype(f32)
            norm = np.sqrt((out * out).sum(axis=1, dtype=f32)).astype(f32)
            fr = (nc < 0) & (norm > f32(1e-30))                   # :679-683
            with np.errstate(divide='ignore', invalid='ignore'):
                unit = (out * (f32(1) / norm)[:, None]).astype(f32)
            scaled = (unit * np.maximum(f32(0), norm + nc * f32(c.friction)
                                        )[:, None]).astype(f32)
            out = np.where(fr[:, None], scaled, out).astype(f32)
        return np.where(inside[:, None], out, v).astype(f32)

    # ---- G2P (:694-724) -------------------------------------------------------
    def g2p(self, dt):
        dt = f32(dt)
        d = self.dim
        n = self.n_particles
        fx, w, inv_idx = self._fx, self._w, self._inv_idx
        offs = self._stencil()
        new_v = np.zeros((n, d), f32)
        new_C = np.zeros((n, d, d), f32)
        four_inv = f32(4 * self.inv_dx)   # `4 * self.inv_dx` is a Python const
        for k, o in enumerate(offs):
            dpos = (o.astype(f32)[None] - fx).astype(f32)
            g_v = self.grid_v[inv_idx[:, k]]
            weight = np.ones(n, f32)
            for a in range(d):
                weight = (weight * w[o[a]][:, a]).astype(f32)
            new_v = (new_v + weight[:, None] * g_v).astype(f32)
            new_C = (new_C + (four_inv * weight)[:, None, None] *
                     (g_v[:, :, None] * dpos[:, None, :])).astype(f32)
        mov = self.material != MATERIAL_STATIONARY                # :722-724
        self.v = np.where(mov[:, None], new_v, self.v).astype(f32)
        self.C = np.where(mov[:, None, None], new_C, self.C).astype(f32)
        if self._packed():                                        # quantised fields: `self.v[p] = new_v` rounds, and
            from .quant_oracle import round_v, round_x            # `self.x[p] += dt * self.v[p]` reads it back (:723-724)
            self.v = round_v(self.v)
        self.x = np.where(mov[:, None], self.x + dt * self.v, self.x).astype(f32)
        if self._packed():
            self.x = round_x(self.x)

    # ---- fused mode (:363-485, host :773-787) ------------------------------------
    def _lin(self, cells):
        shift = cells - np.array(self.offset, np.int64)
        lin = np.zeros(cells.shape[0], np.int64)
        for a in range(self.dim):
            lin = lin * self.grid_size + shift[:, a]
        return lin

    def substep_g2p2g(self, dt):
        """One use_g2p2g substep: gather from the previous output grid with the OLD positions,
        advect, then scatter with the NEW positions into the other grid; normalise, clamp,
        post-process."""
        dtf = f32(dt)
        d, n = self.dim, self.n_particles
        inv_dx = f32(self.inv_dx)
        base = self.base_index()
        fx = (self.x * inv_dx - base.astype(f32)).astype(f32)
        w = self._weights(fx)
        new_v = np.zeros((n, d), f32)
        C = np.zeros((n, d, d), f32)
        old = np.arange(n) < self.last_time_final_particles         # :396-399
        four_inv = f32(4 * self.inv_dx)
        if self._prev_grid is not None and old.any():
            keys, gv = self._prev_grid
            for o in self._stencil():
                lin = self._lin(base[old] + o[None])
                pos = np.searchsorted(keys, lin)
                assert np.array_equal(keys[pos], lin)
                g_v = gv[pos]
                dpos = (o.astype(f32)[None] - fx[old]).astype(f32)
                weight = np.ones(int(old.sum()), f32)
                for a in range(d):
                    weight = (weight * w[o[a]][old, a]).astype(f32)
                new_v[old] = (new_v[old] + weight[:, None] * g_v).astype(f32)
                C[old] = (C[old] + (four_inv * weight)[:, None, None] * (g_v[:, :, None] * dpos[:, None, :])).astype(f32)
        new_v[~old] = self.v[~old]
        mov = self.material != MATERIAL_STATIONARY                # :401-403
        self.v = np.where(mov[:, None], new_v, self.v).astype(f32)
        if self._packed():                                        # quantised fields (:106-114): a store rounds
            from .quant_oracle import round_v, round_x, round_F
            self.v = round_v(self.v)
        self.x = np.where(mov[:, None], self.x + dtf * self.v, self.x).astype(f32)
        if self._packed():
            self.x = round_x(self.x)
        self.C = C                                                # a register value in the reference (:385)
        self.p2g(dt, g2p2g=True)
        if self._packed():
            self.F = round_F(self.F, self.F_bound)                # the final store of F (:445, 466)
        self.last_time_final_particles = n                        # :485
        self.grid_op(dt, self.t)
        self.t += float(dt)
        order = np.argsort(self._lin(self.grid_cells))
        self._prev_grid = (self._lin(self.grid_cells)[order], self.grid_v[order])

    def compute_max_grid_velocity(self):                          # :737-746
        if self.grid_v is None or len(self.grid_v) == 0:
            return 0.0
        return float(np.abs(self.grid_v).max())

    def compute_max_velocity(self):                               # :726-735
        if self.n_particles == 0:
            return 0.0
        return float(np.abs(self.v).max())

    # ---- host loop (:748-805) ---------------------------------------------------
    def substep(self, dt):
        if self.use_g2p2g:
            return self.substep_g2p2g(dt)
        self.p2g(dt)
        self.grid_op(dt, self.t)
        self.t += float(dt)
        self.g2p(dt)

    @staticmethod
    def substep_schedule(frame_dt, default_dt):
        """Replays the host arithmetic of step() (:752-771); returns (dt, count)."""
        substeps = int(frame_dt / default_dt) + 1
        dt = frame_dt / substeps
        left = frame_dt
        count = 0
        while left > 0:
            count += 1
            left -= dt
        return dt, count

    def step_adaptive(self, frame_dt, allowed_cfl=0.9):
        """step() with use_adaptive_dt=True (:752-771): dt only ever shrinks inside a frame."""
        substeps = int(frame_dt / self.default_dt) + 1
        dt = frame_dt / substeps
        left = frame_dt
        dts = []
        while left > 0:
            self.total_substeps += 1
            max_grid_v = self.compute_max_grid_velocity()
            cfl_dt = allowed_cfl * self.dx / (max_grid_v + 1e-6)
            dt = min(dt, cfl_dt, left)
            left -= dt
            self.substep(dt)
            dts.append(dt)
            self.all_time_max_velocity = max(self.all_time_max_velocity, self.compute_max_velocity())
        return dts

    def step(self, frame_dt):
        dt, count = self.substep_schedule(frame_dt, self.default_dt)
        for _ in range(count):
            self.total_substeps += 1
            self.substep(dt)
            cur = self.compute_max_velocity()
            self.all_time_max_velocity = max(self.all_time_max_velocity, cur)

    # ---- diagnostics used by invariant tests -------------------------------------
    def particle_mass(self):
        mass = np.full(self.n_particles, self.p_mass, np.float64)
        mass[self.material == MATERIAL_WATER] *= self.water_density
        return mass

    def total_momentum(self):
        return (self.particle_mass()[:, None] * self.v.astype(np.float64)).sum(0)
