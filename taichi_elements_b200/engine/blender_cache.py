"""Blender add-on particle cache, the frame format either side of `MPMSolver.particle_info()`
in the reference's Blender driver (ref blender/particles_io.py:9-99, written by
blender/operators.py:228-247).

One frame `particles_NNNNNN` is an index file `<name>.bin`
    u32 format version (1) | u32 particle count | 5 x (u32 length, utf-8 file name)
and five raw attribute files `<name>_{pos,vel,col,mat,emt}.bin`
    pos, vel  float32 (n, dim)      col, mat, emt  int32 (n,)
in the fixed attribute order POS, VEL, COL, MAT, EMT.  No Blender (`bpy`) is needed here:
relative names are resolved against the cache folder.
"""
import os
import struct

import numpy as np

PARS_FMT_VER = 1
POS, VEL, COL, MAT, EMT = range(5)
ATTR_NAMES = ('pos', 'vel', 'col', 'mat', 'emt')
ATTR_TYPES = (np.float32, np.float32, np.int32, np.int32, np.int32)
_KEYS = ('position', 'velocity', 'color', 'material', 'emitter_ids')


def frame_name(frame):
    return 'particles_{0:0>6}'.format(frame)      # ref blender/operators.py:233


def write_frame(folder, frame, info):
    """Writes one frame of `MPMSolver.particle_info()` (emitter ids default to 0, as the solver
    reports them only with use_emitter_id=True); returns the index file's path."""
    os.makedirs(folder, exist_ok=True)
    name = frame_name(frame)
    n = int(np.asarray(info['position']).shape[0])
    arrays = []
    for key, dtype in zip(_KEYS, ATTR_TYPES):
        a = info.get(key)
        arrays.append(np.zeros(n, dtype) if a is None else np.ascontiguousarray(a, dtype=dtype))
    index = bytearray(struct.pack('II', PARS_FMT_VER, n))
    for attr, a in zip(ATTR_NAMES, arrays):
        fname = '{}_{}.bin'.format(name, attr).encode('utf-8')
        index += struct.pack('I', len(fname)) + fname
        a.tofile(os.path.join(folder, '{}_{}.bin'.format(name, attr)))
    path = os.path.join(folder, name + '.bin')
    with open(path, 'wb') as f:
        f.write(bytes(index))
    return path


def read_frame(index_path, dim=3):
    """Inverse of write_frame: dict with the particle_info() keys."""
    folder = os.path.dirname(os.path.abspath(index_path))
    with open(index_path, 'rb') as f:
        data = f.read()
    ver, n = struct.unpack_from('II', data, 0)
    if ver != PARS_FMT_VER:
        raise ValueError('Unsupported particles format version: {0}'.format(ver))
    offs, out = 8, {}
    for key, dtype in zip(_KEYS, ATTR_TYPES):
        (length, ) = struct.unpack_from('I', data, offs)
        offs += 4
        fname = data[offs:offs + length].decode('utf-8')
        offs += length
        a = np.fromfile(os.path.join(folder, fname), dtype=dtype)
        out[key] = a.reshape(n, dim) if key in ('position', 'velocity') else a
        assert out[key].shape[0] == n
    return out
