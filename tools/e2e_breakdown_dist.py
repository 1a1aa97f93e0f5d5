"""Wall-clock breakdown of bench.py's end-to-end frame under torchrun (N > 1), rank 0 prints."""
import contextlib, io, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bench import workload
from taichi_elements_b200.distributed import DistributedMPMSolver

world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
w = workload('cube_drop_4m', rank=rank, world=world)
side, res = 0.25, 256
x0 = 0.5 - side * world / 2
cuts = [int((math.floor((x0 + side * k) * res) + 2048) // 4) for k in range(1, world)]
with contextlib.redirect_stdout(io.StringIO()):
    s = DistributedMPMSolver(res=w['res'], cuts=cuts, size=1, unbounded=False, device=local, mig_capacity=1 << 14,
                             halo_capacity=1 << 11, substep_batch=20)
    s.reserve_blocks(1 << 16)
s.set_gravity(w['gravity'])
shift = np.array([x0 + side * rank - (0.5 - side / 2), 0, 0], np.float32)
parts = []
for r in range(world):                                  # rows of this rank's slab, selected once (as bench.py does)
    wr = workload('cube_drop_4m', rank=r, world=world)
    sh = np.array([x0 + side * r - (0.5 - side / 2), 0, 0], np.float32)
    for x, m in wr['parts']:
        rows = (x + sh)[s.slab.mine((x + sh)[:, 0])]
        if len(rows):
            parts.append((torch.from_numpy(np.ascontiguousarray(rows)).pin_memory().numpy(), m))


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


dev = torch.device('cuda', local)
for it in range(6):     # bare small all-reduce + read-back, as _global_box does
    t0 = T()
    a = torch.tensor([1, 2, 3], dtype=torch.int64, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MIN)
    v = a.tolist()
    t1 = T()
    if rank == 0:
        print(f'bare all_reduce+tolist {1e3*(t1-t0):.2f} ms', flush=True)
for it in range(4):
    t0 = T(); s.clear_particles()
    t1 = T()
    for x, m in parts:
        s.add_particles(x, m)
    t2 = T()
    if it == 3 and rank == 0:
        import cProfile, pstats
        pr = cProfile.Profile(); pr.enable()
    ta = T(); lo, hi = s._local_box(); tb = T()
    t_lo = torch.tensor(lo, dtype=torch.int64, device=dev); t_hi = torch.tensor(hi, dtype=torch.int64, device=dev); tc = T()
    dist.all_reduce(t_lo, op=dist.ReduceOp.MIN); td = T()
    dist.all_reduce(t_hi, op=dist.ReduceOp.MAX); te = T()
    l1 = t_lo.tolist(); l2 = t_hi.tolist(); tf = T()
    dist.barrier(); tg = T()
    if rank == 0: print(f'  local_box {1e3*(tb-ta):.2f} tensors {1e3*(tc-tb):.2f} ar1 {1e3*(td-tc):.2f} ar2 {1e3*(te-td):.2f} tolist {1e3*(tf-te):.2f} barrier {1e3*(tg-tf):.2f} ms', flush=True)
    with contextlib.redirect_stdout(io.StringIO()):
        s.step(w['frame_dt'])
    if it == 3 and rank == 0:
        pr.disable()
        st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(18); print(st.getvalue()[:3500], flush=True)
    t3 = T(); s.flush_migration(); t4 = T(); info = s.particle_info(); t5 = T()
    if True:
        print(f'[rank {rank}] frame {it}: clear {1e3*(t1-t0):.1f}  add {1e3*(t2-t1):.1f}  step {1e3*(t3-t2):.1f}  flush {1e3*(t4-t3):.1f}  '
              f'particle_info {1e3*(t5-t4):.1f} ms  n {len(info["id"])}', flush=True)
dist.destroy_process_group()
