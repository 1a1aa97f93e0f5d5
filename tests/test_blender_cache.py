"""Blender particle-cache format (ref blender/particles_io.py) against files written by the
reference's own writer (tests/golden/make_blender_cache_golden.py)."""
import os

import numpy as np

from taichi_elements_b200.engine import blender_cache as bc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'blender_cache')


def test_write_frame_matches_reference_bytes(tmp_path):
    info = dict(np.load(os.path.join(GOLD, 'input.npz')))
    path = bc.write_frame(str(tmp_path), 12, info)
    assert os.path.basename(path) == 'particles_000012.bin'
    names = sorted(f for f in os.listdir(GOLD) if f.startswith('particles_'))
    assert names == sorted(os.listdir(str(tmp_path))) and len(names) == 6
    for f in names:
        with open(os.path.join(GOLD, f), 'rb') as a, open(os.path.join(str(tmp_path), f), 'rb') as b:
            assert a.read() == b.read(), f


def test_read_frame_of_reference_files():
    info = dict(np.load(os.path.join(GOLD, 'input.npz')))
    got = bc.read_frame(os.path.join(GOLD, 'particles_000012.bin'))
    for k, v in info.items():
        assert got[k].dtype == v.dtype and np.array_equal(got[k], v), k


def test_missing_emitter_ids_and_2d_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    info = {'position': rng.random((5, 2), dtype=np.float32), 'velocity': rng.random((5, 2), dtype=np.float32),
            'color': np.arange(5, dtype=np.int32), 'material': np.ones(5, np.int32)}
    got = bc.read_frame(bc.write_frame(str(tmp_path), 3, info), dim=2)
    assert np.array_equal(got['position'], info['position']) and np.array_equal(got['emitter_ids'], np.zeros(5, np.int32))
    bad = tmp_path / 'bad.bin'
    bad.write_bytes(b'\x02\x00\x00\x00\x00\x00\x00\x00')
    try:
        bc.read_frame(str(bad))
        raise AssertionError('version check missing')
    except ValueError:
        pass
