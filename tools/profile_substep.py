"""Small driver for ncu: builds the bench workload and runs a few substeps.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_substep.py --substeps 3
"""
import argparse
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='cube_drop_4m')
ap.add_argument('--warm', type=int, default=2)
ap.add_argument('--substeps', type=int, default=3)
ap.add_argument('--batch', type=int, default=1, help='substeps per mpm_substeps call (> 1: G2P also emits the next keys)')
args = ap.parse_args()

from taichi_elements_b200.engine.mpm_solver import MPMSolver  # noqa: E402

w = workload(args.workload)
with contextlib.redirect_stdout(io.StringIO()):
    mpm = MPMSolver(res=w['res'])
mpm.set_gravity(w['gravity'])
for x, m in w['parts']:
    mpm.add_particles(x, m)
dt = w['frame_dt'] / (int(w['frame_dt'] / mpm.default_dt) + 1)
mpm._run_substeps(dt, 1)              # grows the block workspace (failed attempts launch no-op kernels)
for _ in range(args.warm):
    mpm._run_substeps(dt, 1)
for _ in range(args.substeps):
    mpm._run_substeps(dt, args.batch)
st = mpm.stats()
print('particles', mpm.n_particles[None], 'particle blocks', st.n_particle_blocks, 'grid blocks', st.n_grid_blocks,
      'max_blocks', st.max_blocks, 'key bits', st.key_bits)
