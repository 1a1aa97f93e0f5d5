"""Generates tests/golden/blender_cache/ by running the REFERENCE's own cache writer
(/root/reference/blender/particles_io.py, pure struct/NumPy apart from an unused-on-write
`import bpy`, which is stubbed) exactly as blender/operators.py:228-247 drives it.
Run here (the reference tree does not exist on the GPU box):

    python tests/golden/make_blender_cache_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules['bpy'] = types.ModuleType('bpy')
spec = importlib.util.spec_from_file_location('ref_particles_io', '/root/reference/blender/particles_io.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = os.path.join(HERE, 'blender_cache')
os.makedirs(out, exist_ok=True)
rng = np.random.default_rng(7)
n, frame = 37, 12
info = {
    'position': rng.random((n, 3), dtype=np.float32),
    'velocity': (rng.random((n, 3), dtype=np.float32) - 0.5) * 4,
    'color': rng.integers(0, 1 << 24, n).astype(np.int32),
    'material': rng.integers(0, 5, n).astype(np.int32),
    'emitter_ids': rng.integers(0, 3, n).astype(np.int32),
}
np.savez(os.path.join(out, 'input.npz'), **info)
fname = 'particles_{0:0>6}'.format(frame)                      # blender/operators.py:233
fpath = os.path.join(out, fname)
par_data = {ref.POS: info['position'], ref.VEL: info['velocity'], ref.COL: info['color'],
            ref.MAT: info['material'], ref.EMT: info['emitter_ids']}
data = ref.write_pars(par_data, fpath, fname)
with open(fpath + '.bin', 'wb') as f:                           # blender/operators.py:246-247
    f.write(data)
print('wrote', sorted(os.listdir(out)))
