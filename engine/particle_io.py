from taichi_elements_b200.engine.particle_io import *  # noqa: F401,F403
