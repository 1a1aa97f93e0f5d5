"""torchrun worker: N ranks (NCCL) against the undecomposed solver on rank 0's GPU."""
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
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from taichi_elements_b200.distributed import DistributedMPMSolver, SlabDecomposition
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    dim, res, steps = 3, 32, 24
    comm = os.environ.get('MPM_COMM', 'auto')
    scene = []
    for i, (p, m, vel) in enumerate(mixed_scene(dim, n_per=500, seed=11)):
        vel = list(vel)
        vel[0] = 4.0 if i % 2 == 0 else -4.0
        scene.append((p, m, vel))
    allx = np.concatenate([p for p, _, _ in scene])
    s = DistributedMPMSolver((res, ) * dim, cuts=[0] * 0 if world == 1 else
                             SlabDecomposition.balanced_cuts(allx[:, 0], world, 4, 4096, float(res)),
                             mig_capacity=4096, halo_capacity=512, substep_batch=6, device=local, comm=comm)
    for p, m, vel in scene:
        s.add_particles(p, m, velocity=vel)
    s.reserve_blocks(4096)
    dt = s.default_dt
    s._run_substeps(dt, steps)
    s.flush_migration()
    got = s.gather_rows()
    ok = True
    if rank == 0:
        ref = MPMSolver((res, ) * dim, device=local)
        for p, m, vel in scene:
            ref.add_particles(p, m, velocity=vel)
        ref._run_substeps(dt, steps)
        n = ref.n_particles[None]
        vs = float(np.abs(ref.v.to_numpy()).max())
        ok = (len(got['id']) == n and np.array_equal(got['id'], np.arange(n))
              and np.abs(got['x'] - ref.x.to_numpy()).max() <= 2e-5
              and np.abs(got['v'] - ref.v.to_numpy()).max() <= 5e-3 * vs)
        print('comm', s.comm, 'max dx', np.abs(got['x'] - ref.x.to_numpy()).max(), 'max dv', np.abs(got['v'] - ref.v.to_numpy()).max())
    flag = torch.tensor([int(ok)], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('DIST_OK' if int(flag.item()) == 1 else 'DIST_FAIL')
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
