"""Import-path shim: `from engine.mpm_solver import MPMSolver`, as in the
reference tree (engine/__init__.py:1), resolves to the B200-native engine."""
from taichi_elements_b200.engine import mpm_solver  # noqa: F401
