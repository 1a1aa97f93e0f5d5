"""The MPMSolver API on the GPU: seeding, the reference's own test scenes,
colliders, read-back / export, capacity growth, conservation over many substeps."""
import ctypes
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solver(*a, **k):
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    return MPMSolver(*a, **k)


def _seed_of(s, call):
    return (int(s.rng_seed) * 0x9E3779B97F4A7C15 + call) & 0xFFFFFFFFFFFFFFFF


@pytest.mark.parametrize('dim', [2, 3])
def test_add_cube_and_ellipsoid_counts_and_positions(dim):
    from oracle import seeding_oracle as so
    s = _solver((64, ) * dim)
    s.rng_seed = 7
    lower, size = [0.2, 0.3, 0.25][:dim], [0.3, 0.2, 0.1][:dim]
    s.add_cube(lower, size, s.material_water, color=0x112233, velocity=[1, 2, 3][:dim])
    n1 = int(2**dim * np.prod(size) / s.dx**dim + 1)                      # reference :873
    assert s.n_particles[None] == n1
    center, radius = [0.6, 0.6, 0.6][:dim], [0.1, 0.05, 0.08][:dim]
    s.add_ellipsoid(center, radius, s.material_sand)
    vol = math.pi if dim == 2 else 4 / 3 * math.pi
    for r in radius:
        vol *= r * s.inv_dx
    n2 = int(math.ceil(vol * 2**dim))                                       # reference :997-1005
    assert s.n_particles[None] == n1 + n2
    info = s.particle_info()
    x = info['position']
    assert np.array_equal(x[:n1], so.seed_cube(_seed_of(s, 1), 0, n1, lower, size))       # bit-exact
    assert np.array_equal(x[n1:], so.seed_ellipsoid(_seed_of(s, 2), n1, n2, center, radius))
    q = (x[n1:] - np.array(center, np.float32)) / np.array(radius, np.float32)
    assert (q * q).sum(1).max() <= 1 + 1e-5
    # uniformity: each axis of the cube is flat to a few sigma
    for k in range(dim):
        h, _ = np.histogram(x[:n1, k], bins=8, range=(lower[k], lower[k] + size[k]))
        assert abs(h - n1 / 8).max() < 6 * math.sqrt(n1 / 8)
    assert np.all(info['material'][:n1] == 0) and np.all(info['material'][n1:] == 3)
    assert np.all(info['color'][:n1] == 0x112233) and np.all(info['color'][n1:] == 0xFFFFFF)
    assert np.allclose(info['velocity'][:n1], [1, 2, 3][:dim]) and np.all(info['velocity'][n1:] == 0)
    assert np.all(s.Jp.to_numpy()[:n1] == 1) and np.all(s.Jp.to_numpy()[n1:] == 0)   # :831-835
    assert np.array_equal(s.F.to_numpy(), np.tile(np.eye(dim, dtype=np.float32), (n1 + n2, 1, 1)))


def test_reference_scene_test_2d():
    """tests/test_2d.py:8-16 of the reference: water inflow at res 24."""
    s = _solver(res=(24, 24))
    for frame in range(5):
        s.step(8e-3)
        s.add_cube(lower_corner=[0.3, 0.7], cube_size=[0.2, 0.01], material=s.material_water,
                   velocity=[math.sin(frame * 0.1), 0])
        p = s.particle_info()
    assert s.total_substeps == 50                 # 5 frames x 10 substeps
    assert p['position'].shape == (5 * 5, 2) and np.isfinite(p['position']).all()


def test_config0_demo_2d_against_oracle():
    """BASELINE configs[0] = demo/demo_2d.py:28-45 of the reference (res 128, three ELASTIC cubes, a SAND
    jet every frame, a WATER sheet from frame 11), 30 frames x 53 substeps through MPMSolver.step against
    the NumPy oracle fed the same seeded particles: cubes in free fall, jet and sheet landing on them."""
    from oracle.mpm_oracle import OracleMPM
    s = _solver(res=(128, 128))
    o = OracleMPM((128, 128))
    seen = [0]

    def mirror(material, velocity=None):
        pos = s.particle_info()['position']
        o.add_particles(pos[seen[0]:], material, velocity=velocity)
        seen[0] = len(pos)

    for i in range(3):
        s.add_cube(lower_corner=[0.2 + i * 0.1, 0.3 + i * 0.1], cube_size=[0.1, 0.1], material=s.material_elastic)
        mirror(1)
    assert seen[0] == 3 * 656                      # SURVEY 8(d) cfg 1
    for frame in range(30):
        s.step(8e-3)
        o.step(8e-3)
        s.add_cube(lower_corner=[0.1, 0.8], cube_size=[0.01, 0.05], velocity=[1, 0], material=s.material_sand)
        mirror(3, [1, 0])
        if 10 < frame < 100:
            vel = [math.sin(frame * 0.1), 0]
            s.add_cube(lower_corner=[0.6, 0.7], cube_size=[0.2, 0.01], material=s.material_water, velocity=vel)
            mirror(0, vel)
    assert s.total_substeps == o.total_substeps == 30 * 53      # App. C-1: int(8e-3 / dt) + 1
    s.step(8e-3)
    o.step(8e-3)
    p = s.particle_info()
    assert len(p['position']) == len(o.x) == 3 * 656 + 30 * 33 + 19 * 132
    assert np.array_equal(p['material'], o.material)
    assert np.abs(p['position'] - o.x).max() < 1e-4              # measured 2e-6 after 1643 substeps
    assert np.abs(p['velocity'] - o.v).max() < 5e-3 * max(1.0, np.abs(o.v).max())   # measured 1e-4 of max |v|


def test_reference_scene_test_3d_against_oracle():
    """tests/test_3d.py:9-22: snow ball + elastic bar, size=10, g=(0,-50,0); same particles
    injected into the oracle."""
    from oracle.mpm_oracle import OracleMPM
    s = _solver(res=(24, 24, 24), size=10)
    s.add_ellipsoid(center=[2, 4, 3], radius=1, material=s.material_snow, velocity=[0, -10, 0])
    s.add_cube(lower_corner=[2, 6, 3], cube_size=[1, 1, 3], material=s.material_elastic)
    s.set_gravity((0, -50, 0))
    info = s.particle_info()
    o = OracleMPM((24, 24, 24), size=10)
    o.set_gravity((0, -50, 0))
    snow = info['material'] == 2
    o.add_particles(info['position'][snow], 2, velocity=[0, -10, 0])
    o.add_particles(info['position'][~snow], 1)
    for frame in range(2):
        s.step(4e-3)
        o.step(4e-3)
    assert s.total_substeps == o.total_substeps == 12
    p = s.particle_info()
    assert np.abs(p['position'] - o.x).max() < 2e-4 * 10
    assert np.abs(p['velocity'] - o.v).max() < 5e-3 * max(1.0, np.abs(o.v).max())
    assert abs(s.all_time_max_velocity - o.all_time_max_velocity) < 1e-2 * o.all_time_max_velocity


def test_reference_scene_colliders_against_oracle():
    """tests/test_3d_collider.py and test_3d_surface_collider.py: three sphere colliders (all
    surface modes) and a sticky plane with a non-axis normal."""
    from oracle.mpm_oracle import OracleMPM
    s = _solver(res=(24, 24, 24), size=1)
    o = OracleMPM((24, 24, 24), size=1)
    for m in (s, o):
        m.set_gravity((0, -20, 0))
        m.add_sphere_collider(center=(0.25, 0.5, 0.5), radius=0.1, surface=1)
        m.add_sphere_collider(center=(0.5, 0.5, 0.5), radius=0.1, surface=0)
        m.add_sphere_collider(center=(0.75, 0.5, 0.5), radius=0.1, surface=2)
        m.add_surface_collider((0.5, 0.3, 0.5), (1.0, 1.0, 0.0))
    with pytest.raises(ValueError):
        s.add_surface_collider((0, 0, 0), (0, 1, 0), s.surface_sticky, friction=0.3)
    n0 = 0
    for frame in range(3):
        for lo, col in (((0.2, 0.8, 0.45), 0x8888FF), ((0.45, 0.8, 0.45), 0xFF8888), ((0.7, 0.8, 0.45), 0xFFFFFF)):
            s.add_cube(lo, (0.1, 0.03, 0.1), s.material_water, color=col)
        x = s.particle_info()['position']
        o.add_particles(x[n0:], 0)
        n0 = len(x)
        s.step(4e-3)
        o.step(4e-3)
    p = s.particle_info()
    assert np.abs(p['position'] - o.x).max() < 2e-4
    assert np.abs(p['velocity'] - o.v).max() < 5e-3 * max(1.0, np.abs(o.v).max())


def test_add_mesh_voxelizer_matches_restatement():
    """tests/test_3d_mesh.py:23-38 with a procedural closed mesh (the PLY asset needs network)."""
    from oracle import seeding_oracle as so
    s = _solver(res=(32, 32, 32))
    tris = so.icosphere((0.5, 0.5, 0.5), 0.2, subdiv=2)
    s.add_mesh(triangles=tris, material=s.material_elastic, color=0xFFFF00, velocity=(0, -2, 0))
    vox, lo = s.voxelizer.voxels_numpy()
    ref = so.voxelize(tris, s.voxelizer.res, s.voxelizer.dx, padding=3)
    sub = ref[lo[0]:lo[0] + vox.shape[0], lo[1]:lo[1] + vox.shape[1], lo[2]:lo[2] + vox.shape[2]]
    assert np.array_equal(vox, sub) and ref.sum() == sub.sum()            # winding counts identical
    filled = int((ref > 0).sum())
    n = s.n_particles[None]
    assert n == filled                    # s = 8 / 2^3 = 1 particle per filled super-sampled voxel
    x = s.particle_info()['position']
    r = np.linalg.norm(x - 0.5, axis=1)
    assert r.max() < 0.2 + 2 * s.voxelizer.dx and abs(n * s.voxelizer.dx**3 / (4 / 3 * math.pi * 0.2**3) - 1) < 0.08
    cell = np.floor(x / np.float32(s.voxelizer.dx)).astype(np.int64)
    assert np.all(ref[cell[:, 0], cell[:, 1], cell[:, 2]] > 0)            # every particle sits in a filled voxel
    s.add_mesh(triangles=tris, material=s.material_snow, sample_density=16, translation=(0.1, 0.0, 0.0))
    assert s.n_particles[None] == n + 2 * filled
    s.step(4e-3)
    assert np.isfinite(s.particle_info()['position']).all()


def test_write_particles_round_trip(tmp_path):
    from taichi_elements_b200.engine.particle_io import ParticleIO
    s = _solver(res=(32, 32, 32))
    s.add_cube((0.3, 0.3, 0.3), (0.2, 0.2, 0.2), s.material_elastic, color=0xA0B0C0, velocity=(1, -2, 0.5))
    s.step(2e-3)
    fn = str(tmp_path / 'frame.npz')
    s.write_particles(fn, slice_size=1000)
    # packed on the device (mpm_pack_particles) == the reference's NumPy loop over copy_ranged, byte for byte
    class _HostPath:
        def __init__(self, solver):
            self._s = solver
        def __getattr__(self, name):
            if name == '_pack_particles':
                raise AttributeError(name)
            return getattr(self._s, name)
    fn_host = str(tmp_path / 'frame_host.npz')
    ParticleIO.write_particles(_HostPath(s), fn_host, slice_size=1000)
    dev, host = np.load(fn), np.load(fn_host)
    for key in ('ranges', 'x_and_v', 'color'):
        assert dev[key].dtype == host[key].dtype and dev[key].shape == host[key].shape
        assert np.array_equal(dev[key], host[key]), key
    x, v, color = ParticleIO.read_particles_3d(fn)
    info = s.particle_info()
    span = info['position'].max(0) - info['position'].min(0)
    assert np.abs(x - info['position']).max() <= span.max() * 2.0**-22
    vspan = info['velocity'].max(0) - info['velocity'].min(0)
    assert np.abs(v - info['velocity']).max() <= max(vspan.max(), 1e-5) / 255 * 0.51 + 1e-6
    assert np.all(color == np.array([0xA0, 0xB0, 0xC0], np.uint8))
    buf = np.empty(100, np.float32)
    s.copy_ranged(buf, s.x.get_scalar_field(1), 50, 150)
    assert np.array_equal(buf, info['position'][50:150, 1])
    s.write_particles_ply(str(tmp_path / 'frame.ply'))
    assert os.path.getsize(str(tmp_path / 'frame.ply')) > 16 * s.n_particles[None]


def test_large_add_particles_is_stored_block_sorted_but_reads_back_in_insertion_order():
    """add_particles with >= 2^15 host positions stores the rows sorted by leaf block (so the first substep
    does not gather at random); ids keep the insertion order, and the physics is the same as the oracle's."""
    from oracle.mpm_oracle import OracleMPM
    rng = np.random.default_rng(21)
    pos = (0.25 + 0.5 * rng.random((50000, 3))).astype(np.float32)
    s = _solver(res=(64, 64, 64))
    s.add_particles(pos[:30000], s.material_elastic, velocity=(0.5, -1.0, 0.25))   # small enough: input order
    s.add_particles(pos[30000:40000], s.material_water)
    s.add_particles(pos[10000:], s.material_snow, color=0x123456)                # 40000 rows: block-sorted
    want = np.concatenate([pos[:30000], pos[30000:40000], pos[10000:]])
    assert np.array_equal(s.x.to_numpy(), want)
    mat = s.material.to_numpy()
    assert np.all(mat[:30000] == 1) and np.all(mat[30000:40000] == 0) and np.all(mat[40000:] == 2)
    o = OracleMPM((64, 64, 64))
    o.add_particles(pos[:30000], 1, velocity=(0.5, -1.0, 0.25))
    o.add_particles(pos[30000:40000], 0)
    o.add_particles(pos[10000:], 2, color=0x123456)
    for _ in range(2):
        o.substep(o.default_dt)
    st = s._run_substeps(o.default_dt, 2)
    assert st.substeps_done == 2
    assert np.abs(s.x.to_numpy() - o.x).max() < 1e-5
    assert np.abs(s.v.to_numpy() - o.v).max() < 2e-3 * max(1.0, np.abs(o.v).max())
    assert np.array_equal(s.color.to_numpy()[40000:], np.full(40000, 0x123456))


def test_capacity_growth_and_batched_step_equivalence():
    rng = np.random.default_rng(5)
    pts = (rng.random((60000, 3)) * 0.5 + 0.25).astype(np.float32)        # > initial 16384 rows, > 1024 blocks at res 128
    a, b = _solver(res=(128, ) * 3), _solver(res=(128, ) * 3)
    b.substep_batch = 7
    for s in (a, b):
        s.add_particles(pts[:20000], s.material_water)
        s.add_particles(pts[20000:], s.material_elastic, velocity=(0, -1, 0))
        s.step(3e-3)
    assert a.total_substeps == b.total_substeps
    assert a.stats().n_grid_blocks > 1024                                 # the block workspace had to grow
    xa, xb = a.particle_info()['position'], b.particle_info()['position']
    assert np.abs(xa - xb).max() < 1e-5
    assert a.particle_info()['material'].sum() == 40000
    a.clear_particles()
    assert a.n_particles[None] == 0 and a.particle_info()['position'].shape == (0, 3)
    a.step(1e-3)                                                          # stepping an empty solver is a no-op


def test_conservation_over_1000_substeps():
    """north_star: bounded drift of mass, momentum and energy (zero gravity, free blob)."""
    s = _solver(res=(32, 32, 32))
    s.set_gravity((0, 0, 0))
    s.clear_grid_postprocess()
    rng = np.random.default_rng(6)
    pts = (rng.random((4000, 3)) * 0.2 + 0.4).astype(np.float32)
    s.add_particles(pts[:2000], s.material_elastic, velocity=(0.3, 0.1, -0.2))
    s.add_particles(pts[2000:], s.material_water, velocity=(-0.3, 0.0, 0.2))
    m = s.p_mass

    def momentum():
        return m * s.v.to_numpy().astype(np.float64).sum(0)

    def kinetic():
        return 0.5 * m * (s.v.to_numpy().astype(np.float64)**2).sum()

    p0, e0 = momentum(), kinetic()
    dt = s.default_dt
    for _ in range(10):
        s._run_substeps(dt, 100)
    p1, e1 = momentum(), kinetic()
    scale = m * 4000 * 0.4
    assert np.abs(p1 - p0).max() < 1e-3 * scale          # APIC transfers conserve linear momentum
    assert e1 < e0 * 1.001 and np.isfinite(e1)           # kinetic energy does not grow (it feeds strain energy)
    assert s.n_particles[None] == 4000 and np.isfinite(s.x.to_numpy()).all()
