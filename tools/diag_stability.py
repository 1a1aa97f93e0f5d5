"""Development aid: step a bench workload in small batches and print max |v|, block counts and the fastest particle."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='multimat_100m')
ap.add_argument('--steps', type=int, default=300)
ap.add_argument('--every', type=int, default=20)
ap.add_argument('--g2p2g', action='store_true')
ap.add_argument('--quant', action='store_true')
args = ap.parse_args()
w = bench.workload(args.workload)
chunks, cuts = bench.rank_chunks(w, 0, 1)
s = bench.make_solver(w, 1, 0, 0, args, cuts)
bench.seed(s, chunks, 1)
dt = bench.substep_dt(w, s.default_dt)
done = 0
while done < args.steps:
    try:
        st = s._run_substeps(dt, args.every)
    except Exception as e:
        print('FAILED after', done, 'substeps:', e)
        st = s.stats()
        print('bbox', list(st.bbox_min), list(st.bbox_max))
        break
    done += args.every
    print(done, 'max|v| %.3f' % st.max_velocity, 'blocks', st.n_grid_blocks, st.n_particle_blocks, 'cap', st.max_blocks,
          'bbox', list(st.bbox_min), list(st.bbox_max), flush=True)
v = s.v.to_numpy()
x = s.x.to_numpy()
sp = np.abs(v).max(1)
k = np.argsort(sp)[-5:]
print('fastest:', [(int(i), x[i].tolist(), v[i].tolist(), int(s.material.to_numpy()[i])) for i in k])
print('non-finite x:', int((~np.isfinite(x)).any(1).sum()), 'non-finite v:', int((~np.isfinite(v)).any(1).sum()))
F = s.F.to_numpy()
d = np.linalg.det(F.astype(np.float64))
print('det F range', d.min(), d.max())
