"""Slab decomposition on the GPU.  The full exchange logic (shared-column sums,
particle migration, dropped/appended rows) is exercised on ONE device by driving
several slab solvers in lock-step and copying the message buffers by hand
(loop-back), and compared with the undecomposed solver; the NCCL path itself is
covered by tests/dist_worker.py under torchrun when >= 2 GPUs are present."""
import os
import subprocess
import sys

import numpy as np
import pytest

from scenes import mixed_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _loopback_run(world, scene, dim, res, dt, steps, colliders=(), batch=4):
    import torch
    from taichi_elements_b200.distributed import DistributedMPMSolver, SlabDecomposition
    allx = np.concatenate([p for p, _, _ in scene])
    probe = DistributedMPMSolver((res, ) * dim, cuts=[], world=1, rank=0)
    cuts = SlabDecomposition.balanced_cuts(allx[:, 0], world, probe.leaf_block_size, probe.grid_size, probe.inv_dx)
    del probe
    ranks = [DistributedMPMSolver((res, ) * dim, cuts=cuts, world=world, rank=r, mig_capacity=4096,
                                  halo_capacity=512) for r in range(world)]
    for s in ranks:
        for kind, args in colliders:
            getattr(s, kind)(*args)
        for p, m, vel in scene:
            s.add_particles(p, m, velocity=vel)
        s.reserve_blocks(4096)
    assert sum(s.n_particles[None] for s in ranks) == len(allx)

    def copy(kind):
        for r, s in enumerate(ranks):
            send = s._mig_send if kind == 'migration' else s._halo_send
            if s.slab.left is not None:
                dst = ranks[r - 1]._mig_recv if kind == 'migration' else ranks[r - 1]._halo_recv
                dst[1].copy_(send[0])
            if s.slab.right is not None:
                dst = ranks[r + 1]._mig_recv if kind == 'migration' else ranks[r + 1]._halo_recv
                dst[0].copy_(send[1])

    def global_box():
        boxes = [s._local_box() for s in ranks]
        lo = [min(b[0][d] for b in boxes) for d in range(3)]
        hi = [max(b[1][d] for b in boxes) for d in range(3)]
        return lo, hi

    def exchange(kind, solver):
        if kind == 'box':
            return global_box()
        raise AssertionError

    left, migrated = steps, 0
    while left > 0:
        nb = min(batch, left)
        glo, ghi = global_box()
        for s in ranks:
            s._batch_begin(glo, ghi)
        for _ in range(nb):
            copy('migration')
            migrated += sum(int(t[0].item()) for s in ranks for t in s._mig_recv if t is not None)
            for s in ranks:
                s._substep_pre(dt)
            copy('halo')
            for s in ranks:
                s._substep_post(dt)
        for s in ranks:
            assert s._batch_end() == 0, s._lib.mpm_last_error(s._ctx)
        left -= nb
    # deliver the last substep's leavers, then gather by global id
    glo, ghi = global_box()
    for s in ranks:
        s._batch_begin(glo, ghi)
    copy('migration')
    for s in ranks:
        ptr = lambda t: t.data_ptr() if t is not None else None
        s._check(s._lib.mpm_phase_unpack(s._ctx, ptr(s._mig_recv[0]), ptr(s._mig_recv[1]), s._stream()), 'unpack')
        assert s._batch_end() == 0
    parts = [s.local_rows() for s in ranks]
    merged = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    order = np.argsort(merged['id'], kind='stable')
    torch.cuda.synchronize()
    return {k: v[order] for k, v in merged.items()}, migrated, [len(p['id']) for p in parts]


@pytest.mark.parametrize('dim,world', [(3, 2), (3, 3), (2, 2)])
def test_slabs_match_single_domain(dim, world):
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    # blobs with strong +-x velocities so that particles really cross the cuts
    scene = []
    res = 32 if dim == 3 else 128          # 2D leaves are 16 cells wide: give the cuts room
    for i, (p, m, vel) in enumerate(mixed_scene(dim, n_per=500 if dim == 3 else 3000, seed=11)):
        vel = list(vel)
        vel[0] = 4.0 if i % 2 == 0 else -4.0
        scene.append((p, m, vel))
    cols = [('add_surface_collider', ((0.5, 0.2, 0.5)[:dim], (0.0, 1.0, 0.0)[:dim], 1, 0.2))]
    ref = MPMSolver((res, ) * dim)
    for kind, args in cols:
        getattr(ref, kind)(*args)
    for p, m, vel in scene:
        ref.add_particles(p, m, velocity=vel)
    dt, steps = ref.default_dt, 24 if dim == 3 else 60
    ref._run_substeps(dt, steps)
    got, migrated, counts = _loopback_run(world, scene, dim, res, dt, steps, colliders=cols)
    n = ref.n_particles[None]
    assert len(got['id']) == n and np.array_equal(got['id'], np.arange(n))     # nobody lost or duplicated
    assert migrated > 0 and all(c > 0 for c in counts)
    vs = float(np.abs(ref.v.to_numpy()).max())
    assert np.abs(got['x'] - ref.x.to_numpy()).max() <= 2e-5
    assert np.abs(got['v'] - ref.v.to_numpy()).max() <= 5e-3 * vs
    assert np.abs(got['F'] - ref.F.to_numpy()).max() <= 5e-3
    assert np.array_equal(got['material'], ref.material.to_numpy())


def test_phase_api_equals_substeps_single_rank():
    """world = 1: the phase sequence is the plain substep."""
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    scene = mixed_scene(3, n_per=400, seed=12)
    ref = MPMSolver((32, ) * 3)
    for p, m, vel in scene:
        ref.add_particles(p, m, velocity=vel)
    dt = ref.default_dt
    ref._run_substeps(dt, 6)
    got, migrated, _ = _loopback_run(1, scene, 3, 32, dt, 6)
    assert migrated == 0
    assert np.abs(got['x'] - ref.x.to_numpy()).max() <= 1e-6
    assert np.abs(got['v'] - ref.v.to_numpy()).max() <= 1e-4 * float(np.abs(ref.v.to_numpy()).max())


@pytest.mark.parametrize('comm', ['peer', 'nccl'])
def test_two_ranks_match_single_domain(comm):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    os.environ['MPM_COMM'] = comm
    port = 29600 + os.getpid() % 1000 + (7 if comm == 'peer' else 0)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_worker.py')]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'DIST_OK' in out.stdout
