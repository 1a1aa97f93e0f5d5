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


def _torchrun(world, env, timeout=600):
    port = 29600 + (os.getpid() * 7 + hash(tuple(sorted(env.items())))) % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_worker.py')]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'DIST_OK' in out.stdout, out.stdout[-3000:]
    return out.stdout


@pytest.mark.parametrize('comm', ['peer', 'nccl'])
def test_two_ranks_match_single_domain(comm):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    _torchrun(2, {'MPM_COMM': comm})


def test_two_ranks_rebalance_over_nccl():
    """The bulk move of a re-cut through NCCL send/recv of device buffers (the one-GPU variant below goes through gloo)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = _torchrun(2, {'MPM_COMM': 'peer', 'MPM_SCENE': 'rebalance'})
    assert 'rebalance moved' in out


@pytest.mark.parametrize('fused', ['1', '0'])
@pytest.mark.parametrize('world', [2, 3])
def test_peer_path_ranks_sharing_one_gpu(world, fused):
    """The transport the scaling runs use -- kernels writing into the neighbour's memory (CUDA IPC mapping), epoch
    words published with st.release.sys and awaited with ld.acquire.sys -- on ONE GPU: the ranks are separate
    processes that share device 0 (time-sliced), the control plane runs on gloo.  fused=1: halo inside P2G, waits
    inside the grid op / unpack kernels; fused=0: the pack / wait / add kernels.  Both must reproduce the
    undecomposed solver, with particles crossing the cuts."""
    out = _torchrun(world, {'MPM_COMM': 'peer', 'MPM_DIST_BACKEND': 'gloo', 'MPM_FUSED_HALO': fused})
    assert 'comm peer' in out


def test_step_with_an_initially_empty_rank():
    """DistributedMPMSolver.step() must run its collectives on a rank that holds no particle yet; here
    rank 1 starts empty and is filled by migration (public step() path, peer transport, one GPU)."""
    _torchrun(2, {'MPM_COMM': 'peer', 'MPM_DIST_BACKEND': 'gloo', 'MPM_SCENE': 'empty_rank'})


def test_block_workspace_is_sized_by_the_solver():
    """A block-capacity miss cannot be retried inside a distributed batch: the solver sizes the workspace from a dry
    run of the block discovery after seeding and keeps a margin between batches (here it starts far too small)."""
    out = _torchrun(2, {'MPM_COMM': 'peer', 'MPM_DIST_BACKEND': 'gloo', 'MPM_BLOCKS': '24'})
    assert 'blocks 24 ->' in out


@pytest.mark.parametrize('world', [2, 3])
def test_rebalance_moves_the_cuts_in_mid_run(world):
    """DistributedMPMSolver.rebalance: the cut planes are moved twice in mid-run (once by hand, once to the cuts
    balance_cuts proposes); the particles that change owner are moved in bulk and the run still reproduces the
    undecomposed solver (peer transport for the substeps, torch.distributed for the bulk moves; one GPU)."""
    out = _torchrun(world, {'MPM_COMM': 'peer', 'MPM_DIST_BACKEND': 'gloo', 'MPM_SCENE': 'rebalance'})
    assert 'rebalance moved' in out


def _loopback_ranks(world, res, cuts, **kw):
    from taichi_elements_b200.distributed import DistributedMPMSolver
    return [DistributedMPMSolver((res, ) * 3, cuts=cuts, world=world, rank=r, mig_capacity=1024, halo_capacity=256, **kw)
            for r in range(world)]


def test_slab_seeding_and_export_follow_the_host_partition():
    """mpm_seed_positions_slab / mpm_export_local select a rank's rows with the binning's f32 arithmetic: they must
    agree with SlabDecomposition.mine (NumPy mirror), including positions exactly on and next to the cut planes."""
    res, leaf, gs = 32, 4, 4096
    cuts = [2048 // 4 + 3, 2048 // 4 + 5]
    rng = np.random.default_rng(9)
    x = rng.random((60000, 3)).astype(np.float32)
    edges = (np.array([c * leaf - gs // 2 for c in cuts], np.float32) + np.float32(0.5)) / np.float32(res)
    near = np.array([np.nextafter(e, np.float32(d)) for e in edges for d in (-1, 2)] + list(edges), np.float32)
    x[:len(near), 0] = near
    ranks = _loopback_ranks(3, res, cuts)
    total = 0
    for s in ranks:
        s.add_particles(x[:40000], 1, color=0x123456, velocity=(1, 2, 3))
        s.add_particles(x[40000:], 2)
        info = s.particle_info()
        want = np.nonzero(s.slab.mine(x[:, 0]))[0]
        assert np.array_equal(np.sort(info['id']), want)
        order = np.argsort(info['id'])
        assert np.array_equal(info['position'][order], x[want])
        assert np.array_equal(info['material'][order], np.where(want < 40000, 1, 2))
        assert np.allclose(info['velocity'][order][want < 40000], [1, 2, 3])
        total += len(want)
        with pytest.raises(NotImplementedError):
            s.x.to_numpy()                      # insertion-order read-back is single-device
    assert total == len(x)


def test_distributed_add_cube_ellipsoid_mesh_equal_the_single_device_seeders():
    """add_cube / add_ellipsoid / add_mesh on the distributed solver (every rank the same calls) give exactly the
    particles of the single-device solver, split by slab (global id = insertion index)."""
    from oracle.seeding_oracle import icosphere
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    res = 64
    cuts = [2048 // 4 + 7]
    ref = MPMSolver((res, ) * 3)
    ranks = _loopback_ranks(2, res, cuts)
    mesh = icosphere((0.45, 0.6, 0.5), 0.09, 2)
    for s in [ref] + ranks:
        s.rng_seed = 5
        s.add_cube((0.3, 0.2, 0.3), (0.3, 0.15, 0.2), s.material_water, color=0x010203, velocity=(0, -1, 0))
        s.add_ellipsoid((0.45, 0.5, 0.4), (0.12, 0.05, 0.1), s.material_sand)
        s.add_mesh(mesh, s.material_elastic, velocity=(1, 0, 0), emmiter_id=3)
    want = ref.particle_info()
    n = ref.n_particles[None]
    parts = [s.particle_info() for s in ranks]
    assert all(len(p['id']) > 0 for p in parts) and sum(len(p['id']) for p in parts) == n
    ids = np.concatenate([p['id'] for p in parts])
    order = np.argsort(ids)
    assert np.array_equal(ids[order], np.arange(n))
    for k in ('position', 'velocity', 'material', 'color'):
        assert np.array_equal(np.concatenate([p[k] for p in parts])[order], want[k]), k
