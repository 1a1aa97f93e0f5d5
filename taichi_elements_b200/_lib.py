"""ctypes binding of libmpm_b200.so (the C ABI in include/mpm_b200.h).

There is deliberately no fallback: if the shared library is missing or CUDA is
unavailable every entry point raises, so a silent CPU path can never stand in
for the sm_100a kernels.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get('MPM_B200_LIB') or os.path.join(_HERE, 'libmpm_b200.so')   # (override: A/B builds of the kernels)
CSRC = os.path.join(_HERE, 'csrc')

MPM_OK = 0
MPM_E_BLOCK_CAPACITY = 1
MPM_E_KEY_BITS = 2
ABI_VERSION = 4

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
]


class MPMParams(ctypes.Structure):
    _fields_ = [
        ('dim', ctypes.c_int32), ('res', ctypes.c_int32 * 3),
        ('grid_size', ctypes.c_int32), ('leaf', ctypes.c_int32),
        ('padding', ctypes.c_int32), ('support_plasticity', ctypes.c_int32),
        ('device', ctypes.c_int32), ('flags', ctypes.c_int32),
        ('dx', ctypes.c_double), ('inv_dx', ctypes.c_double),
        ('p_vol', ctypes.c_double), ('p_mass', ctypes.c_double),
        ('mu_0', ctypes.c_double), ('lambda_0', ctypes.c_double),
        ('alpha', ctypes.c_double), ('water_density', ctypes.c_double),
        ('g2p2g_cfl', ctypes.c_double),
    ]


class MPMCollider(ctypes.Structure):
    _fields_ = [
        ('kind', ctypes.c_int32), ('surface', ctypes.c_int32),
        ('a', ctypes.c_double * 3), ('b', ctypes.c_double * 3),
        ('friction', ctypes.c_double),
    ]


class MPMStats(ctypes.Structure):
    _fields_ = [
        ('n_particles', ctypes.c_int64),
        ('n_particle_blocks', ctypes.c_int32), ('n_grid_blocks', ctypes.c_int32),
        ('max_blocks', ctypes.c_int32), ('key_bits', ctypes.c_int32),
        ('bbox_min', ctypes.c_int32 * 3), ('bbox_max', ctypes.c_int32 * 3),
        ('max_velocity', ctypes.c_float), ('launches', ctypes.c_int32),
        ('substeps_done', ctypes.c_int32), ('max_grid_velocity', ctypes.c_float),
        ('ms_sort', ctypes.c_float), ('ms_p2g', ctypes.c_float),
        ('ms_grid', ctypes.c_float), ('ms_g2p', ctypes.c_float),
    ]


# every symbol include/mpm_b200.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_dp = ctypes.POINTER(ctypes.c_double)
SYMBOLS = [
    ('mpm_abi_version', _i32, []),
    ('mpm_state_fields', _i32, [_i32]),
    ('mpm_workspace_bytes', ctypes.c_size_t, [_i32, _i64, _i32]),
    ('mpm_create', _i32, [ctypes.POINTER(MPMParams), ctypes.POINTER(_vp)]),
    ('mpm_destroy', _i32, [_vp]),
    ('mpm_last_error', ctypes.c_char_p, [_vp]),
    ('mpm_virtual_fields', _i32, [_i32]),
    ('mpm_ctx_state_fields', _i32, [_vp]),
    ('mpm_bind', _i32, [_vp, _vp, _vp, _vp, _i64, _vp, ctypes.c_size_t, _i32]),
    ('mpm_get_static_rows', _i32, [_vp, ctypes.POINTER(_i64)]),
    ('mpm_set_static_rows', _i32, [_vp, _i64]),
    ('mpm_compact_statics', _i32, [_vp, _vp]),
    ('mpm_get_state', _i32, [_vp, ctypes.POINTER(_i32), ctypes.POINTER(_i64)]),
    ('mpm_set_state', _i32, [_vp, _i32, _i64]),
    ('mpm_set_gravity', _i32, [_vp, _dp]),
    ('mpm_set_colliders', _i32, [_vp, ctypes.POINTER(MPMCollider), _i32]),
    ('mpm_seed_positions', _i32, [_vp, _vp, _i64, _i32, _i32, _dp, _i32, _vp]),
    ('mpm_seed_cube', _i32, [_vp, _i64, _dp, _dp, _i32, _i32, _dp, _i32, ctypes.c_uint64, _vp]),
    ('mpm_seed_ellipsoid', _i32, [_vp, _i64, _dp, _dp, _i32, _i32, _dp, _i32, ctypes.c_uint64, _vp]),
    ('mpm_seed_restart', _i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    ('mpm_substep', _i32, [_vp, _dbl, _dbl, _vp]),
    ('mpm_substeps', _i32, [_vp, _dbl, _dbl, _i32, _vp]),
    ('mpm_get_stats', _i32, [_vp, ctypes.POINTER(MPMStats)]),
    ('mpm_set_profiling', _i32, [_vp, _i32]),
    ('mpm_download', _i32, [_vp, _i32, _i64, _i64, _vp, _vp]),
    ('mpm_gather', _i32, [_vp, _i32, _i64, _i64, _vp, _vp]),
    ('mpm_gather_rows', _i32, [_vp, _i32, _i32, _i64, _i64, _vp, _vp]),
    ('mpm_particle_ranges', _i32, [_vp, _vp, _vp]),
    ('mpm_pack_particles', _i32, [_vp, _vp, _vp, _vp, _vp]),
    ('mpm_set_slab', _i32, [_vp, _i32, _i32, _i32]),
    ('mpm_comm_bytes', ctypes.c_size_t, [_i32, _i32, _i32]),
    ('mpm_bind_comm', _i32, [_vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    ('mpm_get_bbox', _i32, [_vp, _vp, _vp, _vp]),
    ('mpm_set_layout_box', _i32, [_vp, _i32, _vp, _vp]),
    ('mpm_batch_begin', _i32, [_vp, _vp]),
    ('mpm_batch_probe', _i32, [_vp, ctypes.POINTER(_i32), _vp]),
    ('mpm_phase_unpack', _i32, [_vp, _vp, _vp, _vp]),
    ('mpm_phase_p2g', _i32, [_vp, _dbl, _vp]),
    ('mpm_phase_halo_pack', _i32, [_vp, _vp]),
    ('mpm_phase_halo_add', _i32, [_vp, _vp, _vp, _vp]),
    ('mpm_phase_g2p', _i32, [_vp, _dbl, _vp]),
    ('mpm_batch_end', _i32, [_vp, _vp]),
    ('mpm_peer_alloc', _i32, [_vp, _i32, _i32]),
    ('mpm_peer_handle', _i32, [_vp, _vp]),
    ('mpm_peer_open', _i32, [_vp, _i32, _vp]),
    ('mpm_peer_substeps', _i32, [_vp, _dbl, _i32, _i32, _vp]),
    ('mpm_download_raw', _i32, [_vp, _i32, _vp, _vp]),
    ('mpm_seed_positions_slab', _i32, [_vp, _vp, _i64, _i64, _i32, _i32, _dp, _i32, ctypes.POINTER(_i64), _vp]),
    ('mpm_seed_generate', _i32, [_vp, _i32, _i64, _i64, _dp, _dp, ctypes.c_uint64, _vp, _vp]),
    ('mpm_export_local', _i32, [_vp, _vp, ctypes.POINTER(_i64), _vp]),
    ('mpm_rebalance_pack', _i32, [_vp, _vp, _vp, _i32, ctypes.POINTER(_i64), _vp]),
    ('mpm_voxelize', _i32, [_i32, _vp, _i64, _vp, _dbl, _i32, _vp, _vp, _vp, _vp]),
    ('mpm_voxel_sample', _i32, [_i32, _vp, _vp, _vp, _vp, _i32, _i32, _dbl, _dp, _i32, _i32, ctypes.c_uint64, _i32, _vp, _vp, _vp, _vp]),
    ('mpm_debug_binning', _i32, [_vp, _vp, _vp]),
    ('mpm_debug_scan', _i32, [_vp, _vp, _vp, _i64, _vp]),
    ('mpm_debug_blocks', _i32, [_vp, _vp, _vp, ctypes.POINTER(_i32), _vp, ctypes.POINTER(_i32)]),
    ('mpm_debug_grid', _i32, [_vp, _vp, _vp, _i64, ctypes.POINTER(_i64)]),
    ('mpm_debug_particle_update', _i32, [_vp, _dbl, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(_ROOT, 'include', 'mpm_b200.h')]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    """Compile the library in-tree with nvcc for sm_100a (no GPU needed)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = ['nvcc'] + NVCC_FLAGS + ['-o', LIB_PATH, os.path.join(CSRC, 'mpm_api.cu')]
    if verbose:
        print(' '.join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


_lib = None


def load():
    """Load libmpm_b200.so and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(nvcc, sm_100a). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.mpm_abi_version() != ABI_VERSION:
        raise RuntimeError('libmpm_b200.so ABI version mismatch; rebuild')
    _lib = lib
    return lib


class MPMError(RuntimeError):
    pass


def check(lib, ctx, rc, what):
    if rc < 0:
        msg = lib.mpm_last_error(ctx)
        raise MPMError(f'{what} failed ({rc}): {msg.decode() if msg else ""}')
    return rc
