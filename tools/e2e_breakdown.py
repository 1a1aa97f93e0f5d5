"""Wall-clock breakdown of the end-to-end frame bench.py times (N=1)."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import workload
from taichi_elements_b200.engine.mpm_solver import MPMSolver
w = workload('cube_drop_4m')
with contextlib.redirect_stdout(io.StringIO()):
    s = MPMSolver(res=w['res'])
s.set_gravity(w['gravity'])
parts = [(torch.from_numpy(x).pin_memory().numpy(), m) for x, m in w['parts']]
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); s.clear_particles()
    for x, m in parts: s.add_particles(x, m)
    t1 = T()
    with contextlib.redirect_stdout(io.StringIO()):
        s.step(w['frame_dt'])
    t2 = T(); info = s.particle_info(); t3 = T()
    print(f'frame {it}: add {1e3*(t1-t0):.1f} ms  step {1e3*(t2-t1):.1f} ms  particle_info {1e3*(t3-t2):.1f} ms')
s.substep_batch = 64
for it in range(2):
    t0 = T(); s.clear_particles()
    for x, m in parts: s.add_particles(x, m)
    t1 = T()
    with contextlib.redirect_stdout(io.StringIO()):
        s.step(w['frame_dt'])
    t2 = T(); info = s.particle_info(); t3 = T()
    print(f'batched frame {it}: add {1e3*(t1-t0):.1f} ms  step {1e3*(t2-t1):.1f} ms  particle_info {1e3*(t3-t2):.1f} ms')
