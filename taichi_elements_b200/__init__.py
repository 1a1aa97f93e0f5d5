"""B200-native (sm_100a) MLS-MPM engine behind taichi_elements' MPMSolver API.

    from taichi_elements_b200.engine.mpm_solver import MPMSolver
or, for scripts written against the reference,
    from engine.mpm_solver import MPMSolver
"""
__version__ = '0.1.0'
