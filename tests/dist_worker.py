"""torchrun worker: N ranks against the undecomposed solver on rank 0's GPU.

    MPM_DIST_BACKEND = nccl (default; one GPU per rank) | gloo (ranks may share GPU 0: the control plane -- IPC handle
                       exchange, per-batch all-reduce -- runs on gloo, the data path is CUDA peer memory as usual)
    MPM_COMM         = auto | peer | nccl          MPM_FUSED_HALO = 1 | 0 (peer path: fused exchange or legacy kernels)
    MPM_SCENE        = mixed (five blobs crossing the cuts) | empty_rank (rank > 0 starts with no particle and receives
                       them by migration, driven through the public step()) | rebalance (mixed, with the cut planes moved
                       twice in mid-run: DistributedMPMSolver.rebalance / balance_cuts)
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from scenes import mixed_scene  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    backend = os.environ.get('MPM_DIST_BACKEND', 'nccl')
    local = local % torch.cuda.device_count()
    torch.cuda.set_device(local)
    if backend == 'nccl':
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        dist.init_process_group('gloo')
    from taichi_elements_b200.distributed import DistributedMPMSolver, SlabDecomposition
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    dim, res, steps = 3, 32, 24
    comm = os.environ.get('MPM_COMM', 'auto')
    which = os.environ.get('MPM_SCENE', 'mixed')
    scene = []
    if which in ('mixed', 'rebalance'):
        for i, (p, m, vel) in enumerate(mixed_scene(dim, n_per=500, seed=11)):
            vel = list(vel)
            vel[0] = 4.0 if i % 2 == 0 else -4.0
            scene.append((p, m, vel))
        allx = np.concatenate([p for p, _, _ in scene])
        cuts = SlabDecomposition.balanced_cuts(allx[:, 0], world, 4, 4096, float(res)) if world > 1 else []
    else:
        rng = np.random.default_rng(12)
        for m, y in ((1, 0.3), (2, 0.5), (0, 0.7)):
            p = (rng.random((800, 3)) * np.array([0.12, 0.12, 0.2]) + np.array([0.30, y, 0.4])).astype(np.float32)
            scene.append((p, m, [6.0, 0.0, 0.5]))
        cut0 = (16 + 2048) // 4                       # first cut at cell 16 (x = 0.5): everything starts left of it
        cuts = [cut0 + 2 * k for k in range(world - 1)]
    s = DistributedMPMSolver((res, ) * dim, cuts=cuts, mig_capacity=4096, halo_capacity=1024, substep_batch=6,
                             device=local, comm=comm)
    for p, m, vel in scene:
        s.add_particles(p, m, velocity=vel)
    if 'MPM_BLOCKS' in os.environ:
        s._rebind(max_blocks=int(os.environ['MPM_BLOCKS']))      # too small on purpose: the solver must size it itself
    else:
        s.reserve_blocks(4096)
    dt = s.default_dt
    n0 = s.n_particles[None]
    if which == 'mixed':
        s._run_substeps(dt, steps)
    elif which == 'rebalance':
        s._run_substeps(dt, 8)
        moved = s.rebalance([c + 1 for c in cuts])      # shift the cuts by a column
        counts = s._allreduce_ints([s.n_particles[None]], dist.ReduceOp.SUM)
        s._run_substeps(dt, 8)
        # cost-balanced cuts from a "cost" that is just the particle count: the ranks end up with similar counts
        new_cuts = s.balance_cuts(float(len(s.particle_info()['id'])))
        moved2 = s.rebalance(new_cuts)
        s._run_substeps(dt, 8)
        mine = len(s.particle_info()['id'])
        cdev = 'cpu' if backend == 'gloo' else 'cuda'
        allc = [torch.zeros(1, dtype=torch.int64, device=cdev) for _ in range(world)]
        dist.all_gather(allc, torch.tensor([mine], dtype=torch.int64, device=cdev))
        if rank == 0:
            print('rebalance moved', moved, moved2, 'cuts', cuts, '->', new_cuts, 'counts', [int(c) for c in allc])
        assert moved > 0
    else:
        assert (n0 > 0) == (rank == 0), (rank, n0)
        for _ in range(8):
            s.step(6 * dt * 0.999)                     # the public path: 6 substeps per frame
        steps = s.total_substeps
        dt = 6 * dt * 0.999 / 6
    s.flush_migration()
    got = s.gather_rows()
    info = s.gather_particle_info()
    ok = True
    if rank == 0:
        ref = MPMSolver((res, ) * dim, device=local)
        for p, m, vel in scene:
            ref.add_particles(p, m, velocity=vel)
        if which in ('mixed', 'rebalance'):
            ref._run_substeps(dt, steps)
        else:
            for _ in range(8):
                ref.step(6 * s.default_dt * 0.999)
        n = ref.n_particles[None]
        vs = float(np.abs(ref.v.to_numpy()).max())
        ok = (len(got['id']) == n and np.array_equal(got['id'], np.arange(n))
              and np.abs(got['x'] - ref.x.to_numpy()).max() <= 2e-5
              and np.abs(got['v'] - ref.v.to_numpy()).max() <= 5e-3 * vs
              and np.array_equal(info['id'], np.arange(n)) and np.array_equal(info['position'], got['x'])
              and np.array_equal(info['material'], ref.material.to_numpy()))
        print('comm', s.comm, 'scene', which, 'substeps', steps, 'max dx', np.abs(got['x'] - ref.x.to_numpy()).max(),
              'max dv', np.abs(got['v'] - ref.v.to_numpy()).max(), 'launches', s.stats().launches)
    if 'MPM_BLOCKS' in os.environ:
        ok = ok and s._max_blocks > int(os.environ['MPM_BLOCKS'])
        if rank == 0:
            print('blocks', os.environ['MPM_BLOCKS'], '->', s._max_blocks)
    if which == 'empty_rank' and world > 1:
        ok = ok and (s.n_particles[None] > 0 or rank == 0)          # the other ranks received particles
    flag = torch.tensor([int(ok)], device='cuda' if backend == 'nccl' else 'cpu')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('DIST_OK' if int(flag.item()) == 1 else 'DIST_FAIL')
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
