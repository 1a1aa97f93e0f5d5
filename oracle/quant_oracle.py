"""NumPy restatement of the quantised particle storage of the reference (TEST INFRASTRUCTURE ONLY).

engine/mpm_solver.py:106-114, 216-262 (3D): x = 3 x fixed(21 bits, max 2.0), v = 3 x float(frac 19) with one shared 7-bit
exponent, F = 9 x fixed(16 bits, max F_bound + 0.1).  The encodings themselves are Taichi's quantised types
(ti.types.quant.*), which are not in the reference tree [EXT]; restated from their definitions -- PARITY UNPINNED:
  fixed(bits, max), signed: scale = max / 2**(bits-1); q = clip(round_half_away(x / scale), +-(2**(bits-1) - 1)); x' = q * scale
  shared-exponent float:    e = floor(log2(max_d |v_d|)) in [-64, 63]; m_d = clip(round_half_away(v_d / 2**(e-17)),
                            +-(2**18 - 1)); v_d' = m_d * 2**(e-17)
A store rounds; a later load (also inside the same kernel) sees the rounded value.
"""
import numpy as np

f32 = np.float32


def _round_half_away(t):
    t = np.asarray(t, f32)
    return np.trunc(t + np.where(t < 0, f32(-0.5), f32(0.5)).astype(f32)).astype(np.int64)


def fixed_q(x, max_value, bits):
    x = np.asarray(x, f32)
    inv = f32(f32(2 ** (bits - 1)) / f32(max_value))
    lim = 2 ** (bits - 1) - 1
    t = (x * inv).astype(f32)
    q = _round_half_away(np.clip(np.nan_to_num(t, nan=-lim), -lim, lim))
    return np.clip(q, -lim, lim)


def fixed_round(x, max_value, bits):
    scale = f32(f32(max_value) / f32(2 ** (bits - 1)))
    return (fixed_q(x, max_value, bits).astype(f32) * scale).astype(f32)


def round_x(x):
    return fixed_round(x, 2.0, 21)


def round_F(F, F_bound=4.0):
    return fixed_round(F, np.float32(F_bound) + np.float32(0.1), 16)


def shared_exp_q(v):
    """(n, 3) -> integer fractions (n, 3) and exponents (n,)."""
    v = np.asarray(v, f32)
    a = np.abs(v).max(axis=1)
    with np.errstate(divide='ignore'):
        _, ex = np.frexp(a)
    e = np.where(a > 0, ex - 1, -64)
    e = np.clip(e, -64, 63)
    inv = np.ldexp(f32(1.0), (17 - e).astype(np.int32)).astype(f32)
    lim = 2 ** 18 - 1
    t = (v * inv[:, None]).astype(f32)
    q = _round_half_away(np.clip(t, -lim, lim))
    return np.clip(q, -lim, lim), e


def round_v(v):
    q, e = shared_exp_q(v)
    s = np.ldexp(f32(1.0), (e - 17).astype(np.int32)).astype(f32)
    return (q.astype(f32) * s[:, None]).astype(f32)
