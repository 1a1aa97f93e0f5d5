from taichi_elements_b200.engine.voxelizer import *  # noqa: F401,F403
