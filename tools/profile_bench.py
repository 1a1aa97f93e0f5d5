"""Process for ncu to attach to: builds a bench.py workload (default multimat_12m = configs[3] with one eighth of
the z extent), pre-rolls it and runs a few batches.
    ncu --profile-from-start off --set full -k regex:k_p2g3 -c 1 ... python tools/profile_bench.py
Only the final `--steps` substeps lie between cudaProfilerStart/Stop (the pre-roll includes capacity-growth retries whose
kernels are no-ops)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='multimat_12m')
ap.add_argument('--preroll', type=int, default=-1)
ap.add_argument('--steps', type=int, default=6)
ap.add_argument('--batch', type=int, default=3)
ap.add_argument('--g2p2g', action='store_true')
ap.add_argument('--quant', action='store_true')
args = ap.parse_args()
w = bench.workload(args.workload)
chunks, cuts = bench.rank_chunks(w, 0, 1)
s = bench.make_solver(w, 1, 0, 0, args, cuts)
bench.seed(s, chunks, 1)
dt = bench.substep_dt(w, s.default_dt)
pre = w['preroll'] if args.preroll < 0 else args.preroll
s._run_substeps(dt, pre)
s._run_substeps(dt, args.batch)      # one batch outside the profiled range: steady-state capacities
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
left = args.steps
while left > 0:
    st = s._run_substeps(dt, min(args.batch, left))
    left -= args.batch
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f'{s.n_particles[None]} particles, {st.n_grid_blocks} active blocks, {st.n_particle_blocks} particle blocks, '
      f'max |v| {st.max_velocity:.3f}, pre-roll {pre}, {args.steps} substeps in batches of {args.batch}')
