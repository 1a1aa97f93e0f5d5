from taichi_elements_b200.engine.mpm_solver import *  # noqa: F401,F403
