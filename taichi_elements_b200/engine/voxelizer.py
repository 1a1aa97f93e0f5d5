"""Mesh voxelizer behind `MPMSolver.add_mesh` (reference engine/voxelizer.py).

Triangles are rasterised on the device in f64 onto a super-sampled voxel grid
(signed winding count per voxel column) by `mpm_voxelize`; particle positions
are then drawn per filled voxel by `mpm_voxel_sample`.  Only the pixel box the
mesh covers is allocated (the reference allocates sparse 8^3 blocks, :32-35).
"""
import ctypes
import math

import numpy as np
import torch

from .. import _lib


class Voxelizer:
    def __init__(self, res, dx, super_sample=2, precision='f64', padding=3, device=None):
        assert len(res) == 3
        res = list(res)
        for i in range(3):           # round up to a power of two (reference :21-25)
            r = 1
            while r < res[i]:
                r *= 2
            res[i] = r
        print(f'Voxelizer resolution {res}')
        self.super_sample = super_sample
        self.res = tuple(r * super_sample for r in res)
        self.dx = dx / super_sample
        self.inv_dx = 1 / self.dx
        assert precision in ('f64', )   # the reference default; f32 mode is not offered
        self.precision = precision
        self.padding = padding
        self._lib = _lib.load()
        self._device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.voxels = None          # dense int32 box, torch tensor
        self.box_lo = (0, 0, 0)
        self.box_hi = (0, 0, 0)

    def _i3(self, v):
        return (ctypes.c_int32 * 3)(*[int(c) for c in v])

    def voxelize(self, triangles):
        assert isinstance(triangles, np.ndarray)
        triangles = np.ascontiguousarray(triangles.astype(np.float64))
        assert len(triangles.shape) == 2
        assert triangles.shape[1] == 9
        n = len(triangles)
        pad, inv = self.padding, self.inv_dx
        if n == 0:
            self.voxels, self.box_lo, self.box_hi = None, (0, 0, 0), (0, 0, 0)
            return
        pts = triangles.reshape(-1, 3)
        lo = np.floor(pts.min(axis=0) * inv).astype(np.int64) - 1
        hi = np.floor(pts.max(axis=0) * inv).astype(np.int64) + 2
        box_lo = [max(pad, int(lo[0])), max(pad, int(lo[1])), pad]
        box_hi = [min(self.res[0] - pad, int(hi[0])), min(self.res[1] - pad, int(hi[1])),
                  min(self.res[1] - pad, int(hi[2]) + 1)]
        for d in range(3):
            box_hi[d] = max(box_hi[d], box_lo[d])
        shape = tuple(box_hi[d] - box_lo[d] for d in range(3))
        with torch.cuda.device(self._device):
            self.voxels = torch.zeros(shape, dtype=torch.int32, device=self._device)
            self.box_lo, self.box_hi = tuple(box_lo), tuple(box_hi)
            if self.voxels.numel() == 0:
                return
            tris = torch.from_numpy(triangles).to(self._device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
            rc = self._lib.mpm_voxelize(self._device.index, tris.data_ptr(), n, self._i3(self.res), self.dx, pad,
                                        self._i3(box_lo), self._i3(box_hi), self.voxels.data_ptr(), stream)
            if rc != 0:
                raise _lib.MPMError(f'mpm_voxelize failed ({rc})')
            torch.cuda.current_stream(self._device).synchronize()

    def sample_particles(self, sample_density, translation, grid_size, seed):
        """seed_from_voxels (reference engine/mpm_solver.py:1017-1047): (n, 3) f32 device tensor."""
        if self.voxels is None or self.voxels.numel() == 0:
            return torch.zeros((0, 3), dtype=torch.float32, device=self._device)
        tr = (ctypes.c_double * 3)(*([0.0] * 3))
        if translation:
            for i in range(3):
                tr[i] = float(translation[i])
        cell = self.dx      # = solver dx / super_sample
        with torch.cuda.device(self._device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
            counts = torch.empty(self.voxels.numel(), dtype=torch.int32, device=self._device)
            args = (self._device.index, self.voxels.data_ptr(), self._i3(self.res), self._i3(self.box_lo),
                    self._i3(self.box_hi), int(sample_density), int(self.super_sample), cell, tr, int(grid_size),
                    int(self.padding), ctypes.c_uint64(seed))
            rc = self._lib.mpm_voxel_sample(*args, 0, counts.data_ptr(), None, None, stream)
            if rc != 0:
                raise _lib.MPMError(f'mpm_voxel_sample failed ({rc})')
            incl = torch.cumsum(counts, dim=0, dtype=torch.int64)
            total = int(incl[-1].item())
            offsets = (incl - counts).contiguous()
            out = torch.empty((total, 3), dtype=torch.float32, device=self._device)
            if total:
                rc = self._lib.mpm_voxel_sample(*args, 1, None, offsets.data_ptr(), out.data_ptr(), stream)
                if rc != 0:
                    raise _lib.MPMError(f'mpm_voxel_sample failed ({rc})')
                torch.cuda.current_stream(self._device).synchronize()
        return out

    def voxels_numpy(self):
        """Dense winding counts of the allocated box and its lower corner."""
        if self.voxels is None:
            return np.zeros((0, 0, 0), np.int32), self.box_lo
        return self.voxels.cpu().numpy(), self.box_lo
