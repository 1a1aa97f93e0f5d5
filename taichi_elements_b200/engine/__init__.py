from . import mpm_solver  # noqa: F401  (reference engine/__init__.py:1 exposes the same name)
