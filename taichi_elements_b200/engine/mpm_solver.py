"""Drop-in `MPMSolver` whose substep runs as hand-written sm_100a CUDA.

Mirrors the public surface of /root/reference/engine/mpm_solver.py (class
constants, constructor keywords, add_* / collider / step / particle_info /
write_particles methods, and the attributes demos and ParticleIO touch) so that
scripts written against taichi_elements run unchanged.  All device work goes
through the C ABI in include/mpm_b200.h (libmpm_b200.so, ctypes); PyTorch only
owns the device buffers.  No Taichi, no Triton, no CPU fallback: without the
shared library or a CUDA device the constructor raises.
"""
import ctypes
import math
import numbers
import time

import numpy as np
import torch

from .. import _lib

USE_IN_BLENDER = False


class _Scalar:
    """`field[None]` access used for `n_particles[None]`
    (reference :81; demo/demo_3d_letters.py:83)."""

    def __init__(self, getter, setter=None):
        self._get, self._set = getter, setter

    def __getitem__(self, key):
        assert key is None
        return self._get()

    def __setitem__(self, key, value):
        assert key is None
        if self._set is None:
            raise AttributeError('read-only')
        self._set(value)


class _Field:
    """Read-only view of one particle attribute in insertion order; stands in
    for the Taichi fields callers pass to copy_ranged / read with to_numpy
    (engine/particle_io.py:33-38, 68)."""

    def __init__(self, solver, first_word, shape, dtype):
        self._s, self._w0, self.shape_tail, self.dtype = solver, first_word, tuple(shape), dtype
        self.n = len(shape)

    def _words(self):
        return int(np.prod(self.shape_tail)) if self.shape_tail else 1

    def get_scalar_field(self, *idx):
        if len(idx) != len(self.shape_tail):
            raise IndexError('wrong number of indices')
        off = 0
        for i, extent in zip(idx, self.shape_tail):
            off = off * extent + i
        return _Field(self._s, self._w0 + off, (), self.dtype)

    def to_numpy(self, begin=0, end=None):
        """Rows [begin, end) in insertion order: gathered on the device, one copy into pinned
        host memory; the returned array is a fresh, caller-owned buffer."""
        s = self._s
        if getattr(s, '_global_ids', False):
            raise NotImplementedError('insertion-order field read-back is single-device; use particle_info() / '
                                      'gather_particle_info() on a DistributedMPMSolver')
        n = s.n_particles[None]
        end = n if end is None else end
        cnt = max(end - begin, 0)
        tdt = torch.float32 if self.dtype == np.float32 else torch.int32
        host = torch.empty((cnt, self._words()), dtype=tdt, pin_memory=cnt > 0)
        if cnt:
            with torch.cuda.device(s._device):
                dev = torch.empty((cnt, self._words()), dtype=tdt, device=s._device)
                s._check(s._lib.mpm_gather_rows(s._ctx, self._w0, self._words(), begin, end, dev.data_ptr(),
                                                s._stream()), 'mpm_gather_rows')
                host.copy_(dev)
        return host.numpy().reshape((cnt, ) + self.shape_tail)


class _ParticleNode:
    def __init__(self, cell_bytes):
        self._cell_size_bytes = cell_bytes


class MPMSolver:
    material_water = 0
    material_elastic = 1
    material_snow = 2
    material_sand = 3
    material_stationary = 4
    materials = {
        'WATER': material_water,
        'ELASTIC': material_elastic,
        'SNOW': material_snow,
        'SAND': material_sand,
        'STATIONARY': material_stationary,
    }

    surface_sticky = 0   # velocity forced to the collider's
    surface_slip = 1     # normal component removed
    surface_separate = 2  # only the inward normal component removed
    surfaces = {'STICKY': surface_sticky, 'SLIP': surface_slip, 'SEPARATE': surface_separate}

    def __init__(self,
                 res,
                 quant=False,
                 use_voxelizer=True,
                 size=1,
                 max_num_particles=2**30,
                 padding=3,
                 unbounded=False,
                 dt_scale=1,
                 E_scale=1,
                 voxelizer_super_sample=2,
                 use_g2p2g=False,
                 v_clamp_g2p2g=True,
                 use_bls=True,
                 g2p2g_allowed_cfl=0.9,
                 water_density=1.0,
                 support_plasticity=True,
                 use_adaptive_dt=False,
                 use_ggui=False,
                 use_emitter_id=False,
                 device=None):
        self.dim = len(res)
        assert self.dim in (2, 3), "MPM solver supports only 2D and 3D simulations."
        # quant=True (:106-114, 216-262) in 3D: x, v and F are stored bit-packed (csrc/mpm_quant.cuh) -- with use_g2p2g
        # 44 B per particle and set instead of 104 (no C), with the split substep 80 B (C stays f32, as in the
        # reference).  In 2D the state stays f32: same API, unquantised (more accurate) numbers.
        self.packed_storage = bool(quant and self.dim == 3)
        if quant and not self.packed_storage:
            import warnings
            warnings.warn('quant=True in 2D: particle state is kept in f32 in this build')
        self.quant = quant
        self.use_g2p2g = use_g2p2g
        self.v_clamp_g2p2g = v_clamp_g2p2g
        self.use_bls = use_bls            # accepted; shared-memory staging is always on
        self.g2p2g_allowed_cfl = g2p2g_allowed_cfl
        self.water_density = water_density
        self.grid_size = 4096
        self.t = 0.0
        self.res = res
        self.dx = size / res[0]
        self.inv_dx = 1.0 / self.dx
        self.default_dt = 2e-2 * self.dx / size * dt_scale
        self.p_vol = self.dx**self.dim
        self.p_rho = 1000
        self.p_mass = self.p_vol * self.p_rho
        self.max_num_particles = max_num_particles
        self.input_grid = 0
        self.all_time_max_velocity = 0
        self.support_plasticity = support_plasticity
        self.use_adaptive_dt = use_adaptive_dt
        self.use_ggui = use_ggui
        self.use_emitter_id = use_emitter_id
        self.F_bound = 4.0
        if unbounded:
            # the virtual domain must exceed twice the resolution (reference :143-148)
            while self.grid_size <= 2 * max(self.res):
                self.grid_size *= 2
        self.offset = tuple(-self.grid_size // 2 for _ in range(self.dim))
        self.leaf_block_size = 16 if self.dim == 2 else 4
        self.block_offset = tuple(o // self.leaf_block_size for o in self.offset)
        self.num_grids = 2 if use_g2p2g else 1
        self.padding = padding
        self.E, self.nu = 1e6 * size * E_scale, 0.2
        self.mu_0 = self.E / (2 * (1 + self.nu))
        self.lambda_0 = self.E * self.nu / ((1 + self.nu) * (1 - 2 * self.nu))
        sin_phi = math.sin(math.radians(45))
        self.alpha = math.sqrt(2 / 3) * 2 * sin_phi / (3 - sin_phi)
        self.total_substeps = 0
        self.unbounded = unbounded
        self.voxelizer_super_sample = voxelizer_super_sample
        self.writers = []
        self.rng_seed = 0          # base of the counter-based seeding generator
        self._seed_calls = 0
        self.substep_batch = 16    # substeps enqueued per host synchronisation (1 = the reference's per-substep read-back)

        # ---- native engine -------------------------------------------------
        if not torch.cuda.is_available():
            raise RuntimeError('taichi_elements_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self._lib = _lib.load()
        self._device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        p = _lib.MPMParams()
        p.dim = self.dim
        for d in range(3):
            p.res[d] = int(res[d]) if d < self.dim else 1
        p.grid_size, p.leaf, p.padding = self.grid_size, self.leaf_block_size, int(padding)
        p.support_plasticity = int(bool(support_plasticity))
        p.flags = (1 if use_g2p2g else 0) | (2 if quant else 0)
        p.g2p2g_cfl = float(g2p2g_allowed_cfl) if (use_g2p2g and v_clamp_g2p2g and g2p2g_allowed_cfl > 0) else 0.0
        p.device = self._device.index
        p.dx, p.inv_dx, p.p_vol, p.p_mass = self.dx, self.inv_dx, self.p_vol, self.p_mass
        p.mu_0, p.lambda_0, p.alpha, p.water_density = self.mu_0, self.lambda_0, self.alpha, water_density
        ctx = ctypes.c_void_p()
        rc = self._lib.mpm_create(ctypes.byref(p), ctypes.byref(ctx))
        if rc != 0:
            raise _lib.MPMError(f'mpm_create failed ({rc})')
        self._ctx = ctx
        self._nf = self._lib.mpm_ctx_state_fields(ctx)           # physical words per particle (x v F C Jp tag; 11 packed)
        self._cap = 0
        self._max_blocks = 0
        self._state = None
        self._static = None
        self._ws = None
        self._n = 0
        self._rebind(capacity=1 << 14, max_blocks=1 << 10)

        d, dd = self.dim, self.dim * self.dim
        self.x = _Field(self, 0, (d, ), np.float32)
        self.v = _Field(self, d, (d, ), np.float32)
        self.F = _Field(self, 2 * d, (d, d), np.float32)
        self.C = _Field(self, 2 * d + dd, (d, d), np.float32)
        jp = 2 * d + 2 * dd
        if support_plasticity:
            self.Jp = _Field(self, jp, (), np.float32)
        self.material = _Field(self, jp + 1, (), np.int32)
        self.color = _Field(self, jp + 2, (), np.int32)
        if use_emitter_id:
            self.emitter_ids = _Field(self, jp + 4, (), np.int32)
        self.n_particles = _Scalar(lambda: self._n)
        self.particle = _ParticleNode((self._nf + 3) * 4)     # bytes per particle: one state set + the static row

        self.grid_postprocess = []
        if self.dim == 2:
            self.voxelizer = None
            self.set_gravity((0, -9.8))
        else:
            if use_voxelizer:
                from .voxelizer import Voxelizer
                self.voxelizer = Voxelizer(res=self.res,
                                           dx=self.dx,
                                           padding=self.padding,
                                           super_sample=voxelizer_super_sample,
                                           device=self._device)
            else:
                self.voxelizer = None
            self.set_gravity((0, -9.8, 0))
        self.add_bounding_box(self.unbounded)

    def __del__(self):
        ctx = getattr(self, '_ctx', None)
        if ctx:
            try:
                self._lib.mpm_destroy(ctx)
            except Exception:
                pass
            self._ctx = None

    # ------------------------------------------------------------ plumbing
    def _check(self, rc, what):
        return _lib.check(self._lib, self._ctx, rc, what)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _rebind(self, capacity=None, max_blocks=None):
        """(Re)allocate torch-owned buffers and bind them; live particles are kept."""
        cap = self._cap if capacity is None else (int(capacity) + 63) // 64 * 64
        mb = self._max_blocks if max_blocks is None else int(max_blocks)
        with torch.cuda.device(self._device):
            if cap != self._cap:
                # one set = tiles of 32 particles x nf words x 32 lanes (include/mpm_b200.h): rows are tile-major,
                # so the live particles are a prefix of the buffer
                new_state = torch.empty((2, cap // 32, self._nf, 32), dtype=torch.int32, device=self._device)
                new_static = torch.empty((3, cap), dtype=torch.int32, device=self._device)   # colour, id, emitter by sid
                if self._n > 0:
                    cur = ctypes.c_int32()
                    self._lib.mpm_get_state(self._ctx, ctypes.byref(cur), None)
                    tiles = (self._n + 31) // 32
                    new_state[0, :tiles] = self._state[cur.value, :tiles]
                    ns = ctypes.c_int64()
                    self._lib.mpm_get_static_rows(self._ctx, ctypes.byref(ns))
                    new_static[:, :ns.value] = self._static[:, :ns.value]
                self._state, self._static = new_state, new_static
                cur_set = 0
            else:
                cur = ctypes.c_int32()
                self._lib.mpm_get_state(self._ctx, ctypes.byref(cur), None)
                cur_set = cur.value
            nbytes = self._lib.mpm_workspace_bytes(self.dim, cap, mb)
            # use_g2p2g keeps the last substep's output grid and block table in the workspace: the old one must stay
            # alive until mpm_bind has copied them over; otherwise it is released first (less peak memory)
            old_ws = self._ws if self.use_g2p2g else None
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            torch.cuda.synchronize(self._device)
        self._cap, self._max_blocks = cap, mb
        self._check(
            self._lib.mpm_bind(self._ctx, self._state[0].data_ptr(), self._state[1].data_ptr(),
                               self._static.data_ptr(), cap, self._ws.data_ptr(), nbytes, mb), 'mpm_bind')
        del old_ws
        self._check(self._lib.mpm_set_state(self._ctx, cur_set, self._n), 'mpm_set_state')

    def _reserve(self, new_particles):
        need = self._n + new_particles
        assert need <= self.max_num_particles
        if need > self._cap:
            self._rebind(capacity=max(need, int(self._cap * 1.5)))

    def _vec(self, values, default=0.0):
        arr = (ctypes.c_double * 3)(default, default, default)
        if values is not None:
            values = list(values)
            assert len(values) == self.dim
            for i, val in enumerate(values):
                arr[i] = float(val)
        return arr

    def _next_seed(self):
        self._seed_calls += 1
        return (int(self.rng_seed) * 0x9E3779B97F4A7C15 + self._seed_calls) & 0xFFFFFFFFFFFFFFFF

    def _download_word(self, word, begin, end, out):
        if end <= begin:
            return
        assert out.flags['C_CONTIGUOUS'] and out.itemsize == 4 and out.size >= end - begin
        self._check(
            self._lib.mpm_download(self._ctx, word, begin, end, out.ctypes.data_as(ctypes.c_void_p), self._stream()),
            'mpm_download')

    def stats(self):
        s = _lib.MPMStats()
        self._check(self._lib.mpm_get_stats(self._ctx, ctypes.byref(s)), 'mpm_get_stats')
        return s

    # ------------------------------------------------------------ configuration
    def stencil_range(self):
        return np.ndindex(*((3, ) * self.dim))

    def set_gravity(self, g):
        assert isinstance(g, (tuple, list))
        assert len(g) == self.dim
        self.gravity = tuple(float(c) for c in g)
        self._check(self._lib.mpm_set_gravity(self._ctx, self._vec(g)), 'mpm_set_gravity')

    def _push_colliders(self):
        n = len(self.grid_postprocess)
        table = (_lib.MPMCollider * max(n, 1))()
        for i, c in enumerate(self.grid_postprocess):
            table[i] = c
        self._check(self._lib.mpm_set_colliders(self._ctx, table, n), 'mpm_set_colliders')

    def add_sphere_collider(self, center, radius, surface=surface_sticky):
        c = _lib.MPMCollider()
        c.kind, c.surface = 1, int(surface)
        for i, val in enumerate(list(center)):
            c.a[i] = float(val)
        c.b[0] = float(radius)
        self.grid_postprocess.append(c)
        self._push_colliders()

    def clear_grid_postprocess(self):
        self.grid_postprocess.clear()
        self._push_colliders()

    def add_surface_collider(self, point, normal, surface=surface_sticky, friction=0.0):
        point = list(point)
        inv_len = 1.0 / math.sqrt(sum(c**2 for c in normal))
        normal = [inv_len * c for c in normal]
        if surface == self.surface_sticky and friction != 0:
            raise ValueError('friction must be 0 on sticky surfaces.')
        c = _lib.MPMCollider()
        c.kind, c.surface, c.friction = 2, int(surface), float(friction)
        for i in range(self.dim):
            c.a[i], c.b[i] = float(point[i]), float(normal[i])
        self.grid_postprocess.append(c)
        self._push_colliders()

    def add_bounding_box(self, unbounded):
        c = _lib.MPMCollider()
        c.kind = 0
        c.a[0] = 1.0 if unbounded else 0.0
        self.grid_postprocess.append(c)
        self._push_colliders()

    # ------------------------------------------------------------ time stepping
    def _run_substeps(self, dt, count):
        """`count` substeps on the device; grows the block workspace on demand."""
        left = count
        while left > 0:
            rc = self._check(self._lib.mpm_substeps(self._ctx, dt, self.t, left, self._stream()), 'mpm_substeps')
            st = self.stats()
            if rc == _lib.MPM_OK:
                return st
            if rc == _lib.MPM_E_BLOCK_CAPACITY:
                # nothing past the completed substeps was modified; enlarge and go on
                left -= st.substeps_done
                self._rebind(max_blocks=max(2 * st.n_grid_blocks, 2 * self._max_blocks))
                continue
            raise _lib.MPMError(self._lib.mpm_last_error(self._ctx).decode())
        return self.stats()

    def compute_max_velocity(self):
        return float(self.stats().max_velocity) if self._n > 0 else 0.0

    def compute_max_grid_velocity(self, grid_v=None):
        """max |v|_inf over the active grid cells of the last substep (reference :737-746)."""
        return float(self.stats().max_grid_velocity) if self._n > 0 else 0.0

    def step(self, frame_dt, print_stat=False, smry_writer=None):
        begin_t = time.time()
        begin_substep = self.total_substeps
        substeps = int(frame_dt / self.default_dt) + 1
        dt = frame_dt / substeps
        frame_time_left = frame_dt
        if print_stat:
            print(f'needed substeps: {substeps}')
        # the reference's loop subtracts dt until the residue is <= 0, which
        # often yields substeps+1 iterations; replay it
        pending = 0
        while frame_time_left > 0:
            print('.', end='', flush=True)
            self.total_substeps += 1
            if self.use_adaptive_dt:
                # CFL-limited dt from the last substep's grid (reference :762-770); dt only shrinks in a frame
                max_grid_v = self.compute_max_grid_velocity()
                cfl_dt = self.g2p2g_allowed_cfl * self.dx / (max_grid_v + 1e-6)
                dt = min(dt, cfl_dt, frame_time_left)
                frame_time_left -= dt
                self._advance(dt, 1, smry_writer)
                continue
            frame_time_left -= dt
            pending += 1
            batch = self.substep_batch if smry_writer is None else 1
            if batch <= 1 or pending >= batch or frame_time_left <= 0:
                self._advance(dt, pending, smry_writer)
                pending = 0
        print()
        if print_stat:
            cur_frame_velocity = self.compute_max_velocity()
            st = self.stats()
            print(f'CFL: {cur_frame_velocity * dt / self.dx}')
            print(f'num particles={self.n_particles[None]}')
            print(f'  active blocks={st.n_grid_blocks} particle blocks={st.n_particle_blocks}')
            print(f'  frame time {time.time() - begin_t:.3f} s')
            print(f'  substep time {1000 * (time.time() - begin_t) / (self.total_substeps - begin_substep):.3f} ms')

    def _advance(self, dt, count, smry_writer):
        if self._n > 0:
            st = self._run_substeps(dt, count)
            cur_frame_velocity = float(st.max_velocity)
        else:
            cur_frame_velocity = 0.0
        self.t += dt * count
        if smry_writer is not None:
            smry_writer.add_scalar("substep_max_CFL", cur_frame_velocity * dt / self.dx, self.total_substeps)
        self.all_time_max_velocity = max(self.all_time_max_velocity, cur_frame_velocity)

    # ------------------------------------------------------------ seeding
    def set_source_velocity(self, velocity):
        if velocity is not None:
            velocity = list(velocity)
            assert len(velocity) == self.dim
        self.source_velocity = velocity

    def add_cube(self, lower_corner, cube_size, material, color=0xFFFFFF, sample_density=None, velocity=None):
        if sample_density is None:
            sample_density = 2**self.dim
        vol = 1
        for i in range(self.dim):
            vol = vol * cube_size[i]
        num_new_particles = int(sample_density * vol / self.dx**self.dim + 1)
        self._reserve(num_new_particles)
        self.set_source_velocity(velocity)
        self._check(
            self._lib.mpm_seed_cube(self._ctx, num_new_particles, self._vec(lower_corner), self._vec(cube_size),
                                    int(material), int(color), self._vec(velocity), 0, self._next_seed(),
                                    self._stream()), 'mpm_seed_cube')
        self._n += num_new_particles

    def add_ellipsoid(self, center, radius, material, color=0xFFFFFF, sample_density=None, velocity=None):
        if sample_density is None:
            sample_density = 2**self.dim
        if isinstance(radius, numbers.Number):
            radius = [radius] * self.dim
        radius = list(radius)
        num_particles = math.pi if self.dim == 2 else 4 / 3 * math.pi
        for i in range(self.dim):
            num_particles *= radius[i] * self.inv_dx
        num_particles = int(math.ceil(num_particles * sample_density))
        self._reserve(num_particles)
        self.set_source_velocity(velocity)
        self._check(
            self._lib.mpm_seed_ellipsoid(self._ctx, num_particles, self._vec(center), self._vec(radius),
                                         int(material), int(color), self._vec(velocity), 0, self._next_seed(),
                                         self._stream()), 'mpm_seed_ellipsoid')
        self._n += num_particles

    def add_ngon(self, sides, center, radius, angle, material, color=0xFFFFFF, sample_density=None, velocity=None):
        """Regular polygon by rejection sampling (reference :886-941).  Setup path: the points are
        drawn on the host with the counter-based generator and uploaded like add_particles."""
        if self.dim != 2:
            raise ValueError("Add Ngon only works for 2D simulations")
        if sample_density is None:
            sample_density = 2**self.dim
        num_particles = 0.5 * (radius * self.inv_dx)**2 * math.sin(2 * math.pi / sides) * sides
        num_particles = int(math.ceil(num_particles * sample_density))
        assert self.n_particles[None] + num_particles <= self.max_num_particles
        from ..seeding import polygon_points
        pts = polygon_points(self._next_seed(), self._n, num_particles, sides, angle)
        pts = (np.asarray(center, np.float32)[None] + pts * np.float32(radius)).astype(np.float32)
        self.add_particles(pts, material, color, velocity)

    def add_texture_2d(self, offset_x, offset_y, texture, new_material, color):
        """One particle per texel above 0.1 at (offset + index * dx) (reference :943-957); uses the
        source velocity of the last add_* call, like the reference kernel."""
        assert self.dim == 2
        idx = np.argwhere(np.asarray(texture) > 0.1)
        pts = (np.array([offset_x, offset_y], np.float32)[None] + idx.astype(np.float32) * np.float32(self.dx))
        self.add_particles(pts.astype(np.float32), new_material, color, getattr(self, 'source_velocity', None))

    def add_particles(self, particles, material, color=0xFFFFFF, velocity=None):
        particles = np.ascontiguousarray(np.asarray(particles, dtype=np.float32))
        assert particles.ndim == 2 and particles.shape[1] == self.dim
        n = len(particles)
        self._reserve(n)
        self.set_source_velocity(velocity)
        if n == 0:
            return
        self._seed_from_device(torch.from_numpy(particles).to(self._device), material, color, velocity)

    def _seed_from_device(self, dev, material, color, velocity):
        """add_particles for positions that already live on the device: (n, dim) float32, contiguous."""
        n = int(dev.shape[0])
        self._reserve(n)
        self.set_source_velocity(velocity)
        if n == 0:
            return
        self._check(
            self._lib.mpm_seed_positions(self._ctx, dev.data_ptr(), n, int(material), int(color),
                                         self._vec(velocity), 0, self._stream()), 'mpm_seed_positions')
        torch.cuda.current_stream(self._device).synchronize()
        self._n += n

    def add_mesh(self, triangles, material, color=0xFFFFFF, sample_density=None, velocity=None, translation=None,
                 emmiter_id=0):
        assert self.dim == 3
        if sample_density is None:
            sample_density = 2**self.dim
        self.set_source_velocity(velocity)
        if self.voxelizer is None:
            raise RuntimeError('add_mesh needs use_voxelizer=True')
        self.voxelizer.voxelize(triangles)
        pos = self.voxelizer.sample_particles(sample_density=sample_density,
                                              translation=translation,
                                              grid_size=self.grid_size,
                                              seed=self._next_seed())
        n = int(pos.shape[0])
        self._reserve(n)
        if n == 0:
            return
        self._check(
            self._lib.mpm_seed_positions(self._ctx, pos.data_ptr(), n, int(material), int(color),
                                         self._vec(velocity), int(emmiter_id), self._stream()), 'mpm_seed_positions')
        torch.cuda.current_stream(self._device).synchronize()
        self._n += n

    def read_restart(self, num_particles, pos, vel, material, color):
        pos = np.ascontiguousarray(np.asarray(pos, np.float32)[:num_particles])
        vel = np.ascontiguousarray(np.asarray(vel, np.float32)[:num_particles])
        material = np.ascontiguousarray(np.asarray(material, np.int32)[:num_particles])
        color = np.ascontiguousarray(np.asarray(color).astype(np.int32)[:num_particles])
        self._reserve(num_particles)
        if num_particles == 0:
            return
        t = [torch.from_numpy(a).to(self._device) for a in (pos, vel, material, color)]
        self._check(
            self._lib.mpm_seed_restart(self._ctx, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(),
                                       num_particles, self._stream()), 'mpm_seed_restart')
        torch.cuda.current_stream(self._device).synchronize()
        self._n += num_particles

    def clear_particles(self):
        self._n = 0
        self._check(self._lib.mpm_set_state(self._ctx, 0, 0), 'mpm_set_state')

    # ------------------------------------------------------------ read-back
    def copy_ranged(self, np_x, input_x, begin, end):
        """np_x[0:end-begin] = field[begin:end] for a scalar field (reference :1157-1162)."""
        assert isinstance(input_x, _Field) and input_x.shape_tail == ()
        if np_x.dtype.itemsize == 4 and np_x.flags['C_CONTIGUOUS'] and np_x.dtype == input_x.dtype:
            self._download_word(input_x._w0, begin, end, np_x)
        else:
            tmp = np.empty(end - begin, dtype=input_x.dtype)
            self._download_word(input_x._w0, begin, end, tmp)
            np_x[:end - begin] = tmp

    def copy_ranged_nd(self, np_x, input_x, begin, end):
        np_x[:end - begin] = input_x.to_numpy(begin, end)

    def copy_dynamic(self, np_x, input_x):
        self.copy_ranged(np_x, input_x, 0, self._n)

    def copy_dynamic_nd(self, np_x, input_x):
        np_x[:self._n] = input_x.to_numpy()

    def particle_info(self):
        data = {
            'position': self.x.to_numpy(),
            'velocity': self.v.to_numpy(),
            'material': self.material.to_numpy(),
            'color': self.color.to_numpy(),
        }
        if self.use_emitter_id:
            data['emitter_ids'] = self.emitter_ids.to_numpy()
        return data

    def write_particles(self, fn, slice_size=1000000):
        from .particle_io import ParticleIO
        ParticleIO.write_particles(self, fn, slice_size)

    def _pack_particles(self):
        """(ranges, x_and_v, color) of ParticleIO.write_particles (ref engine/particle_io.py:12-76),
        quantised and packed on the device: 4*dim + 3 bytes per particle cross PCIe instead of the
        reference's slice-by-slice read-back of six f32 fields and the colour."""
        n, dim = self._n, self.dim
        if n == 0:
            return None
        with torch.cuda.device(self._device):
            rdev = torch.empty((2, dim, 2), dtype=torch.float32, device=self._device)
            self._check(self._lib.mpm_particle_ranges(self._ctx, rdev.data_ptr(), self._stream()),
                        'mpm_particle_ranges')
            ranges = rdev.cpu().numpy()
            for c in range(2):          # avoid degenerate ranges (ref :46-47), in f32 like the reference
                for d in range(dim):
                    ranges[c, d, 1] = max(ranges[c, d, 0] + 1e-5, ranges[c, d, 1])
            lo_inv = np.empty((2, dim, 2), np.float32)
            lo_inv[:, :, 0] = ranges[:, :, 0]
            lo_inv[:, :, 1] = 1 / (ranges[:, :, 1] - ranges[:, :, 0])
            xv = torch.empty((n, dim), dtype=torch.int32, device=self._device)
            col = torch.empty((n, 3), dtype=torch.uint8, device=self._device)
            self._check(self._lib.mpm_pack_particles(self._ctx, lo_inv.ctypes.data_as(ctypes.c_void_p),
                                                     xv.data_ptr(), col.data_ptr(), self._stream()),
                        'mpm_pack_particles')
        return ranges, xv.cpu().numpy().view(np.uint32), col.cpu().numpy()

    def write_blender_cache(self, folder, frame):
        """One frame of the Blender add-on's particle cache (ref blender/particles_io.py:40-61 as driven by
        blender/operators.py:228-247 from particle_info())."""
        from .blender_cache import write_frame
        return write_frame(folder, frame, self.particle_info())

    def write_particles_ply(self, fn):
        np_x = self.x.to_numpy()
        np_color = self.color.to_numpy().astype(np.uint32)
        data = np.hstack([np_x, (np_color[:, None]).view(np.float32)])
        from .mesh_io import write_point_cloud
        write_point_cloud(fn, data)

    # ------------------------------------------------------------ parity getters (tests)
    def _inject_state(self, x, v, F, C, Jp, material, color):
        """Replace all particles by a full state (tests: start CUDA and oracle from the same bits)."""
        n, d = x.shape
        assert d == self.dim
        if self.packed_storage:
            raise NotImplementedError('_inject_state writes f32 words; not available with packed storage')
        self.clear_particles()
        self._reserve(n)
        rows = [np.asarray(x, np.float32).T, np.asarray(v, np.float32).T,
                np.asarray(F, np.float32).reshape(n, d * d).T, np.asarray(C, np.float32).reshape(n, d * d).T,
                np.asarray(Jp, np.float32)[None]]
        fl = np.ascontiguousarray(np.concatenate(rows, axis=0)).view(np.int32)
        tag = ((np.asarray(material, np.int64) << 29) | np.arange(n, dtype=np.int64)).astype(np.uint32).view(np.int32)
        words = np.ascontiguousarray(np.concatenate([fl, tag[None]], axis=0))       # (nf, n): x v F C Jp tag
        statics = np.stack([np.asarray(color, np.int32), np.arange(n, dtype=np.int32), np.zeros(n, np.int32)])
        tiles = (n + 31) // 32
        padded = np.zeros((self._nf, tiles * 32), np.int32)
        padded[:, :n] = words
        tiled = np.ascontiguousarray(padded.reshape(self._nf, tiles, 32).transpose(1, 0, 2))   # (tiles, nf, 32)
        self._state[0, :tiles] = torch.from_numpy(tiled).to(self._device)
        self._static[:, :n] = torch.from_numpy(statics).to(self._device)
        torch.cuda.synchronize(self._device)
        self._n = n
        self._check(self._lib.mpm_set_state(self._ctx, 0, n), 'mpm_set_state')
        self._check(self._lib.mpm_set_static_rows(self._ctx, n), 'mpm_set_static_rows')

    def debug_binning(self):
        out = np.empty((self._n, self.dim), np.int32)
        self._check(self._lib.mpm_debug_binning(self._ctx, out.ctypes.data_as(ctypes.c_void_p), self._stream()),
                    'mpm_debug_binning')
        return out

    def debug_blocks(self):
        npb, ngb = ctypes.c_int32(), ctypes.c_int32()
        self._check(self._lib.mpm_debug_blocks(self._ctx, None, None, ctypes.byref(npb), None, ctypes.byref(ngb)),
                    'mpm_debug_blocks')
        pbc = np.empty((npb.value, self.dim), np.int32)
        cnt = np.empty((npb.value, ), np.int32)
        gbc = np.empty((ngb.value, self.dim), np.int32)
        self._check(
            self._lib.mpm_debug_blocks(self._ctx, pbc.ctypes.data_as(ctypes.c_void_p),
                                       cnt.ctypes.data_as(ctypes.c_void_p), ctypes.byref(npb),
                                       gbc.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ngb)), 'mpm_debug_blocks')
        return pbc, cnt, gbc

    def debug_grid(self):
        n = ctypes.c_int64()
        self._check(self._lib.mpm_debug_grid(self._ctx, None, None, 0, ctypes.byref(n)), 'mpm_debug_grid')
        cells = np.empty((n.value, self.dim), np.int32)
        vm = np.empty((n.value, 4), np.float32)
        self._check(
            self._lib.mpm_debug_grid(self._ctx, cells.ctypes.data_as(ctypes.c_void_p),
                                     vm.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)), 'mpm_debug_grid')
        return cells, vm[:, :self.dim].copy(), vm[:, self.dim].copy()

    def debug_particle_update(self, dt, material, F, C, Jp):
        n = len(material)
        F = np.ascontiguousarray(F, np.float32).copy()
        C = np.ascontiguousarray(C, np.float32)
        Jp = np.ascontiguousarray(Jp, np.float32).copy()
        material = np.ascontiguousarray(material, np.int32)
        aff = np.empty_like(F)
        mass = np.empty(n, np.float32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self._check(
            self._lib.mpm_debug_particle_update(self._ctx, dt, n, vp(material), vp(F), vp(C), vp(Jp), vp(aff),
                                                vp(mass)), 'mpm_debug_particle_update')
        return F, Jp, aff, mass
