import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def host_math():
    """Test-only host build of csrc/mpm_math.cuh (see tests/host_harness.cpp)."""
    import ctypes
    import subprocess
    import tempfile
    src = os.path.join(ROOT, 'tests', 'host_harness.cpp')
    out = os.path.join(tempfile.gettempdir(), f'mpm_host_harness_{os.getpid()}.so')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-x', 'c++', '-o', out, src], check=True)
    return ctypes.CDLL(out)
