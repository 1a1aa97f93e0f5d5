"""ctypes wrapper of oracle/mpm_oracle.c (TEST INFRASTRUCTURE ONLY; see the
header of mpm_oracle.c).  `COracle` mirrors oracle.mpm_oracle.OracleMPM for the
substep and is what bench.py times as the CPU baseline."""
import ctypes
import os
import subprocess

import numpy as np

from .mpm_oracle import OracleMPM

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, 'libmpm_oracle.so')


class _Params(ctypes.Structure):
    _fields_ = [('dim', ctypes.c_int), ('res', ctypes.c_int * 3), ('grid_size', ctypes.c_int),
                ('padding', ctypes.c_int), ('quant', ctypes.c_int), ('support_plasticity', ctypes.c_int)] + \
               [(k, ctypes.c_float) for k in ('dx', 'inv_dx', 'p_vol', 'p_mass', 'mu_0', 'lambda_0', 'alpha',
                                              'sand_coef', 'water_density', 'inv_dx2', 'four_inv_dx')] + \
               [('gravity', ctypes.c_float * 3)]


class _Collider(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int), ('surface', ctypes.c_int), ('a', ctypes.c_float * 3),
                ('b', ctypes.c_float * 3), ('r2', ctypes.c_float), ('friction', ctypes.c_float)]


def build(force=False):
    src = os.path.join(_HERE, 'mpm_oracle.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-s', '-C', _HERE, '-B', 'libmpm_oracle.so'], check=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build()                      # (re)compiles when the library is missing or older than its source
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_substep.restype = ctypes.c_int
        _lib.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.oracle_set_threads.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class COracle(OracleMPM):
    """OracleMPM whose substep runs in C/OpenMP on a dense grid box."""

    def _params(self):
        p = _Params()
        p.dim = self.dim
        for d in range(3):
            p.res[d] = self.res[d] if d < self.dim else 1
        p.grid_size, p.padding, p.support_plasticity = self.grid_size, self.padding, int(self.support_plasticity)
        p.quant = int(self._packed())          # split substep only (this class has no g2p2g mode)
        p.dx, p.inv_dx, p.p_vol, p.p_mass = self.dx, self.inv_dx, self.p_vol, self.p_mass
        p.mu_0, p.lambda_0, p.alpha = self.mu_0, self.lambda_0, self.alpha
        p.sand_coef = (self.dim * self.lambda_0 + 2 * self.mu_0) / (2 * self.mu_0)
        p.water_density = self.water_density
        p.inv_dx2, p.four_inv_dx = self.inv_dx**2, 4 * self.inv_dx
        for d in range(self.dim):
            p.gravity[d] = float(self.gravity[d])
        return p

    def _collider_table(self):
        t = (_Collider * max(1, len(self.colliders)))()
        for i, c in enumerate(self.colliders):
            if c.kind == 'bbox':
                t[i].kind, t[i].a[0] = 0, 1.0 if c.unbounded else 0.0
            elif c.kind == 'sphere':
                t[i].kind, t[i].surface, t[i].r2 = 1, c.surface, c.radius * c.radius
                for d in range(self.dim):
                    t[i].a[d] = c.center[d]
            else:
                t[i].kind, t[i].surface, t[i].friction = 2, c.surface, c.friction
                for d in range(self.dim):
                    t[i].a[d], t[i].b[d] = c.point[d], c.normal[d]
        return t

    def substep(self, dt, margin=8):
        assert not self.use_g2p2g, 'the C restatement has the split substep only'
        lib = load()
        n = self.n_particles
        if n == 0:
            return
        base = self.base_index()
        lo = (base.min(axis=0) - margin).astype(np.int32)
        hi = (base.max(axis=0) + 3 + margin).astype(np.int32)
        if not self.unbounded:   # cover the whole bounded domain when it is affordable
            full_lo = np.minimum(lo, 0)
            full_hi = np.maximum(hi, np.array(self.res, np.int32))
            if np.prod((full_hi - full_lo).astype(np.int64)) <= 2**28:
                lo, hi = full_lo.astype(np.int32), full_hi.astype(np.int32)
        ext = (hi - lo).astype(np.int64)
        ncell = int(np.prod(ext))
        if getattr(self, '_gv', None) is None or self._gv.shape[0] < ncell:
            self._gv = np.empty((ncell, self.dim), np.float32)
            self._gm = np.empty((ncell, ), np.float32)
        for name in ('x', 'v', 'F', 'C', 'Jp', 'material'):
            setattr(self, name, np.ascontiguousarray(getattr(self, name)))
        lo3 = (ctypes.c_int * 3)(*[int(c) for c in lo], *([0] * (3 - self.dim)))
        hi3 = (ctypes.c_int * 3)(*[int(c) for c in hi], *([1] * (3 - self.dim)))
        prm = self._params()
        rc = lib.oracle_substep(ctypes.byref(prm), ctypes.c_float(dt), ctypes.c_int64(n), _p(self.x), _p(self.v),
                                _p(self.F), _p(self.C), _p(self.Jp), _p(self.material), lo3, hi3, _p(self._gv),
                                _p(self._gm), self._collider_table(), len(self.colliders))
        if rc != 0:
            raise RuntimeError('oracle_substep: a stencil left the grid box')
        self.t += float(dt)
        self._box = (lo, hi)

    @staticmethod
    def num_threads():
        return load().oracle_num_threads()
