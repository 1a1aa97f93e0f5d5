"""particle_io .npz format against golden files written by the reference's own
ParticleIO (tests/golden/make_particle_io_golden.py): byte-exact."""
import os

import numpy as np
import pytest

from taichi_elements_b200.engine.particle_io import ParticleIO

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


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


class HostSolver:
    """Host stand-in with the attributes write_particles touches."""

    def __init__(self, x, v, color):
        self.dim, self.n_particles = x.shape[1], _Scalar(len(x))
        self.x, self.v, self.color = _Field(x), _Field(v), _Field(color)

    def copy_ranged(self, np_x, f, begin, end):
        np_x[:end - begin] = f.arr[begin:end]


@pytest.mark.parametrize('dim', [2, 3])
def test_writer_is_byte_exact(dim, tmp_path):
    inp = np.load(os.path.join(G, f'particle_io_input_{dim}d.npz'))
    ref = np.load(os.path.join(G, f'particle_io_ref_{dim}d.npz'))
    out = str(tmp_path / 'p.npz')
    ParticleIO.write_particles(HostSolver(inp['x'], inp['v'], inp['color']), out, int(inp['slice_size']))
    got = np.load(out)
    assert sorted(got.files) == sorted(ref.files) == ['color', 'ranges', 'x_and_v']
    for k in ref.files:
        assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape
        assert got[k].tobytes() == ref[k].tobytes(), k


@pytest.mark.parametrize('dim', [2, 3])
def test_reader_matches_reference(dim):
    want = np.load(os.path.join(G, f'particle_io_read_{dim}d.npz'))
    x, v, color = ParticleIO.read_particles(os.path.join(G, f'particle_io_ref_{dim}d.npz'), dim)
    assert x.tobytes() == want['x'].tobytes() and v.tobytes() == want['v'].tobytes()
    assert np.array_equal(color, want['color'])
    inp = np.load(os.path.join(G, f'particle_io_input_{dim}d.npz'))
    span = inp['x'].max(0) - inp['x'].min(0)
    assert np.abs(x - inp['x']).max() <= span.max() * 2.0**-23      # 24-bit positions


def test_ply_point_cloud_layout(tmp_path):
    from taichi_elements_b200.engine.mesh_io import write_point_cloud
    pts = np.arange(12, dtype=np.float32).reshape(3, 4)
    fn = str(tmp_path / 'c.ply')
    write_point_cloud(fn, pts)
    raw = open(fn, 'rb').read()
    head, body = raw.split(b'end_header\n')
    assert b'element vertex 3' in head and b'property uchar placeholder' in head
    assert body == pts.tobytes()
