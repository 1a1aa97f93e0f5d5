"""Lossy .npz particle export in the reference's format (engine/particle_io.py).

File layout (bit-compatible, reference :9-10, 50-56, 71-74):
  ranges   float32 (2, dim, 2)   [x|v][axis][min|max]
  x_and_v  uint32  (n, dim)      (xq << 8) + vq, xq 24 bit, vq 8 bit
  color    uint8   (n, 3)        R, G, B of the packed 0xRRGGBB particle colour
"""
import gc
import time

import numpy as np

from .mesh_io import write_point_cloud


class ParticleIO:
    v_bits = 8
    x_bits = 32 - v_bits

    @staticmethod
    def _quantise(a, lo, hi, bits):
        # (a - lo) * (1 / (hi - lo)) * (2^bits - 1) + 0.499, truncated (reference :50-55)
        return ((a - lo) * (1 / (hi - lo)) * (2**bits - 1) + 0.499).astype(np.uint32)

    @staticmethod
    def write_particles(solver, fn, slice_size=1000000):
        t = time.time()
        n = solver.n_particles[None]
        dim = solver.dim
        packed = solver._pack_particles() if hasattr(solver, '_pack_particles') else None
        if packed is not None:       # quantised and packed on the device, same bytes as the loop below
            ranges, x_and_v, color = packed
            np.savez(fn, ranges=ranges, x_and_v=x_and_v, color=color)
            print(f'Writing to disk: {time.time() - t:.3f} s')
            return
        x_and_v = np.ndarray((n, dim), dtype=np.uint32)
        ranges = np.ndarray((2, dim, 2), dtype=np.float32)
        num_slices = (n + slice_size - 1) // slice_size
        buf = np.ndarray((slice_size, ), dtype=np.float32)

        def fetch(field, dtype=np.float32):
            out = np.ndarray((n, ), dtype=dtype)
            for s in range(num_slices):
                begin, end = slice_size * s, min(slice_size * (s + 1), n)
                solver.copy_ranged(buf, field, begin, end)
                out[begin:end] = buf[:end - begin]
            return out

        for d in range(dim):
            np_x = fetch(solver.x.get_scalar_field(d))
            np_v = fetch(solver.v.get_scalar_field(d))
            ranges[0, d] = [np.min(np_x), np.max(np_x)]
            ranges[1, d] = [np.min(np_v), np.max(np_v)]
            for c in range(2):  # avoid degenerate ranges
                ranges[c, d, 1] = max(ranges[c, d, 0] + 1e-5, ranges[c, d, 1])
            xq = ParticleIO._quantise(np_x, ranges[0, d, 0], ranges[0, d, 1], ParticleIO.x_bits)
            vq = ParticleIO._quantise(np_v, ranges[1, d, 0], ranges[1, d, 1], ParticleIO.v_bits)
            x_and_v[:, d] = (xq << ParticleIO.v_bits) + vq
            del np_x, np_v, xq, vq
        # the reference stages colour through a float32 buffer (:62-69): exact below 2^24
        np_color = fetch(solver.color, dtype=np.uint32)
        color = np.ndarray((n, 3), dtype=np.uint8)
        for c in range(3):
            color[:, c] = (np_color >> (8 * (2 - c))) & 255
        np.savez(fn, ranges=ranges, x_and_v=x_and_v, color=color)
        print(f'Writing to disk: {time.time() - t:.3f} s')

    @staticmethod
    def write_arrays(fn, np_x, np_v, np_color):
        """The same file from host arrays x (n, dim), v (n, dim) f32 and packed colours (n,): used by the distributed
        solver, whose particles are gathered from the ranks (same arithmetic as the loop above, ref :42-76)."""
        n, dim = np_x.shape
        x_and_v = np.ndarray((n, dim), dtype=np.uint32)
        ranges = np.ndarray((2, dim, 2), dtype=np.float32)
        for d in range(dim):
            xs, vs = np.ascontiguousarray(np_x[:, d], np.float32), np.ascontiguousarray(np_v[:, d], np.float32)
            ranges[0, d] = [np.min(xs), np.max(xs)]
            ranges[1, d] = [np.min(vs), np.max(vs)]
            for c in range(2):
                ranges[c, d, 1] = max(ranges[c, d, 0] + 1e-5, ranges[c, d, 1])
            xq = ParticleIO._quantise(xs, ranges[0, d, 0], ranges[0, d, 1], ParticleIO.x_bits)
            vq = ParticleIO._quantise(vs, ranges[1, d, 0], ranges[1, d, 1], ParticleIO.v_bits)
            x_and_v[:, d] = (xq << ParticleIO.v_bits) + vq
        col = np.asarray(np_color).astype(np.uint32)
        color = np.ndarray((n, 3), dtype=np.uint8)
        for c in range(3):
            color[:, c] = (col >> (8 * (2 - c))) & 255
        np.savez(fn, ranges=ranges, x_and_v=x_and_v, color=color)

    @staticmethod
    def read_particles_3d(fn):
        return ParticleIO.read_particles(fn, 3)

    @staticmethod
    def read_particles_2d(fn):
        return ParticleIO.read_particles(fn, 2)

    @staticmethod
    def read_particles(fn, dim):
        data = np.load(fn)
        ranges, color, x_and_v = data['ranges'], data['color'], data['x_and_v']
        del data
        gc.collect()
        x = (x_and_v >> ParticleIO.v_bits).astype(np.float32) / ((2**ParticleIO.x_bits - 1))
        v = (x_and_v & (2**ParticleIO.v_bits - 1)).astype(np.float32) / (2**ParticleIO.v_bits - 1)
        for c in range(dim):
            x[:, c] = x[:, c] * (ranges[0, c, 1] - ranges[0, c, 0]) + ranges[0, c, 0]
            v[:, c] = v[:, c] * (ranges[1, c, 1] - ranges[1, c, 0]) + ranges[1, c, 0]
        return x, v, color

    @staticmethod
    def convert_particle_to_ply(fns):
        for fn in fns:
            print(f'Converting {fn}...')
            x, _, color = ParticleIO.read_particles_3d(fn)
            x = x.astype(np.float32)
            packed = (color[:, 2].astype(np.uint32) << 16) + (color[:, 1].astype(np.uint32) << 8) + color[:, 0]
            pos_color = np.hstack([x, packed[:, None].view(np.float32)])
            del x, color, packed
            gc.collect()
            write_point_cloud(fn + ".ply", pos_color)


if __name__ == '__main__':
    import sys
    ParticleIO.convert_particle_to_ply(sys.argv[1:])
