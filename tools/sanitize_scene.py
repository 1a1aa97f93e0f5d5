"""Small multi-material scene for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from scenes import mixed_scene
from taichi_elements_b200.engine.mpm_solver import MPMSolver
for dim in (3, 2):
    for g2p2g in (False, True):
        s = MPMSolver((32, ) * dim, use_g2p2g=g2p2g)
        s.add_surface_collider((0.5, 0.2, 0.5)[:dim], (0, 1, 0)[:dim], 1, 0.3)
        for p, m, vel in mixed_scene(dim, n_per=300, seed=3):
            s.add_particles(p, m, velocity=vel)
        s._run_substeps(s.default_dt, 3)
        s.particle_info()
        print('dim', dim, 'g2p2g', g2p2g, 'ok', s.stats().n_grid_blocks)
