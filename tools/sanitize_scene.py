"""Small multi-material scenes for compute-sanitizer (memcheck / racecheck): split and fused (use_g2p2g) substeps in 2D and
3D, the bit-packed storage (quant=True), the bulk-copy G2P variant (MPM_G2P_TILE=1 set by the caller), batches of three
substeps so that the fused key pass and the tile double buffers are exercised, and particle export."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from scenes import mixed_scene
from taichi_elements_b200.engine.mpm_solver import MPMSolver
for dim, g2p2g, quant in ((3, False, False), (3, False, True), (3, True, False), (3, True, True), (2, False, False), (2, True, False)):
    s = MPMSolver((32, ) * dim, use_g2p2g=g2p2g, quant=quant)
    s.add_surface_collider((0.5, 0.2, 0.5)[:dim], (0, 1, 0)[:dim], 1, 0.3)
    for p, m, vel in mixed_scene(dim, n_per=300, seed=3):
        s.add_particles(p, m, velocity=vel)
    s._run_substeps(s.default_dt, 3)
    s._run_substeps(s.default_dt, 3)
    s.particle_info()
    print('dim', dim, 'g2p2g', g2p2g, 'quant', quant, 'ok', s.stats().n_grid_blocks, flush=True)
