"""BASELINE configs[2]-like scene: unbounded domain, six slip planes with friction,
SNOW/SAND ellipsoids seeded with add_ellipsoid (every particle takes the full SVD path),
plus one mesh through the voxelizer.  Prints particle-substeps/s and the phase times."""
import argparse, contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from taichi_elements_b200.engine.mpm_solver import MPMSolver
from oracle.seeding_oracle import icosphere

ap = argparse.ArgumentParser()
ap.add_argument('--ellipsoids', type=int, default=482)
ap.add_argument('--steps', type=int, default=50)   # the timed window covers free flight and the first floor contacts
args = ap.parse_args()
with contextlib.redirect_stdout(io.StringIO()):
    s = MPMSolver(res=(256, 256, 256), unbounded=True)
s.set_gravity((0, -25, 0))
for point, normal in (((0, 0, 0), (0, 1, 0)), ((0, 1.9, 0), (0, -1, 0)), ((-1.9, 0, 0), (1, 0, 0)),
                      ((1.9, 0, 0), (-1, 0, 0)), ((0, 0, -0.95), (0, 0, 1)), ((0, 0, 0.95), (0, 0, -1))):
    s.add_surface_collider(point, normal, s.surface_slip, friction=0.5)
rng = np.random.default_rng(3)
t0 = time.time()
k = 0
for ix in np.arange(-1.6, 1.61, 0.2):
    for iy in np.arange(0.1, 1.51, 0.2):
        for iz in np.arange(-0.8, 0.81, 0.2):
            if k >= args.ellipsoids:
                break
            c = np.array([ix, iy, iz]) + rng.uniform(-0.03, 0.03, 3)
            s.add_ellipsoid(center=list(c), radius=0.048, material=s.material_snow if k % 2 == 0 else s.material_sand,
                            velocity=[0, -5, 0])
            k += 1
s.add_mesh(icosphere((1.0, 1.7, 0.5), 0.08, 2), s.material_elastic, velocity=(0, -5, 0))
torch.cuda.synchronize()
n = s.n_particles[None]
print('seeded', n, 'particles in', round(time.time() - t0, 2), 's')
dt = s.default_dt
s._run_substeps(dt, 20)
s._lib.mpm_set_profiling(s._ctx, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); st = s._run_substeps(dt, args.steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(f'{n * 1e-6:.1f} M particles, {st.n_grid_blocks} active blocks: {ms:.3f} ms/substep = {n / ms * 1e-6:.2f} G particle-substeps/s;'
      f' binning {st.ms_sort:.3f} p2g {st.ms_p2g:.3f} grid {st.ms_grid:.3f} g2p {st.ms_g2p:.3f} ms; max |v| {st.max_velocity:.2f}')
x = s.x.to_numpy()
print('finite', bool(np.isfinite(x).all()), 'y range', float(x[:, 1].min()), float(x[:, 1].max()))
