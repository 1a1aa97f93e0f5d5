"""The C-ABI library loads and exports every symbol include/mpm_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'mpm_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mpm_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported():
    from taichi_elements_b200 import _lib
    _lib.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(n for n, _, _ in _lib.SYMBOLS) == names      # the binding covers the whole header
    assert lib.mpm_abi_version() == _lib.ABI_VERSION


def test_sizes_and_argument_checks():
    from taichi_elements_b200 import _lib
    lib = _lib.load()
    assert lib.mpm_state_fields(3) == 26 and lib.mpm_state_fields(2) == 14 and lib.mpm_state_fields(4) == -1
    assert lib.mpm_virtual_fields(3) == 29 and lib.mpm_virtual_fields(2) == 17      # the reference's field order
    a = lib.mpm_workspace_bytes(3, 1 << 20, 1 << 12)
    b = lib.mpm_workspace_bytes(3, 1 << 21, 1 << 12)
    c = lib.mpm_workspace_bytes(3, 1 << 20, 1 << 13)
    assert 0 < a < b and a < c
    assert lib.mpm_workspace_bytes(5, 1024, 16) == 0
    assert lib.mpm_create(None, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        MPMSolver((32, 32, 32))


def test_reference_import_path_and_constants():
    from engine import mpm_solver            # the reference's import path
    M = mpm_solver.MPMSolver
    assert (M.material_water, M.material_elastic, M.material_snow, M.material_sand, M.material_stationary) == \
        (0, 1, 2, 3, 4)
    assert M.materials['SAND'] == 3 and M.surfaces == {'STICKY': 0, 'SLIP': 1, 'SEPARATE': 2}
    import inspect
    sig = inspect.signature(M.__init__)
    for kw in ('res', 'quant', 'use_voxelizer', 'size', 'max_num_particles', 'padding', 'unbounded', 'dt_scale',
               'E_scale', 'voxelizer_super_sample', 'use_g2p2g', 'v_clamp_g2p2g', 'use_bls', 'g2p2g_allowed_cfl',
               'water_density', 'support_plasticity', 'use_adaptive_dt', 'use_ggui', 'use_emitter_id'):
        assert kw in sig.parameters, kw
    for meth in ('set_gravity', 'add_cube', 'add_ellipsoid', 'add_mesh', 'add_particles', 'add_surface_collider',
                 'add_sphere_collider', 'clear_grid_postprocess', 'step', 'particle_info', 'write_particles',
                 'write_particles_ply', 'copy_ranged', 'read_restart', 'clear_particles'):
        assert callable(getattr(M, meth)), meth
    assert list(inspect.signature(M.add_mesh).parameters)[-1] == 'emmiter_id'   # the reference's spelling


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) needs no GPU and prints one
    JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1'], check=True, capture_output=True, text=True, cwd=root, timeout=600).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'particle-substeps/sec' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['unit'] == 'particle-substeps/s' and line['dtype'] == 'f32'
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
