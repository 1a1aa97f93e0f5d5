from taichi_elements_b200.engine.mesh_io import *  # noqa: F401,F403
