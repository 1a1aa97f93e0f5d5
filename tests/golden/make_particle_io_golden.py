"""Generates tests/golden/particle_io_*.npz by running the REFERENCE's own
ParticleIO (/root/reference/engine/particle_io.py, pure NumPy) on seeded
arrays.  The reference module imports `taichi` and (via engine.mesh_io)
`plyfile` at import time without using them on this path; both are stubbed.
Run here (the reference tree does not exist on the GPU box):

    python tests/golden/make_particle_io_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ('taichi', 'plyfile'):
    m = types.ModuleType(name)
    m.PlyData = object
    sys.modules[name] = m
import importlib.util  # noqa: E402

# engine/__init__.py imports the Taichi solver; load only the two pure-NumPy files
pkg = types.ModuleType('engine')
pkg.__path__ = ['/root/reference/engine']
sys.modules['engine'] = pkg
for mod in ('mesh_io', 'particle_io'):
    spec = importlib.util.spec_from_file_location(f'engine.{mod}', f'/root/reference/engine/{mod}.py')
    m = importlib.util.module_from_spec(spec)
    sys.modules[f'engine.{mod}'] = m
    spec.loader.exec_module(m)
ParticleIO = sys.modules['engine.particle_io'].ParticleIO   # the reference's


class _Scalar:
    def __init__(self, v):
        self.v = v

    def __getitem__(self, k):
        return self.v


class _Field:
    def __init__(self, arr):
        self.arr = arr

    def get_scalar_field(self, d):
        return _Field(self.arr[:, d])


class FakeSolver:
    """Just what ParticleIO.write_particles touches (reference :12-76)."""

    def __init__(self, x, v, color):
        self.dim = x.shape[1]
        self.n_particles = _Scalar(len(x))
        self.x, self.v, self.color = _Field(x), _Field(v), _Field(color)

    def copy_ranged(self, np_x, input_x, begin, end):
        np_x[:end - begin] = input_x.arr[begin:end]     # Taichi casts to the ndarray's dtype


def make(dim, n, seed):
    rng = np.random.default_rng(seed)
    x = (rng.random((n, dim)) * 1.7 - 0.4).astype(np.float32)
    v = rng.normal(size=(n, dim)).astype(np.float32) * 3
    if dim == 3:
        v[:, 2] = 0.25                                   # degenerate range on one axis
    color = rng.integers(0, 1 << 24, size=n).astype(np.int32)
    return x, v, color


if __name__ == '__main__':
    for dim, n, seed, slice_size in ((3, 2500, 11, 1000), (2, 777, 12, 1000000)):
        x, v, color = make(dim, n, seed)
        out = os.path.join(HERE, f'particle_io_ref_{dim}d.npz')
        ParticleIO.write_particles(FakeSolver(x, v, color), out, slice_size)
        np.savez(os.path.join(HERE, f'particle_io_input_{dim}d.npz'), x=x, v=v, color=color,
                 slice_size=slice_size)
        rx, rv, rc = ParticleIO.read_particles(out, dim)
        np.savez(os.path.join(HERE, f'particle_io_read_{dim}d.npz'), x=rx, v=rv, color=rc)
        print('wrote', out)
