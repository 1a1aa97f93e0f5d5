"""Multi-GPU MLS-MPM: 1-D slab decomposition along x, one process per GPU.

The reference is single-device; this layer is new.  A rank owns
the leaf-block columns [lo, hi) of the x axis; `torch.distributed` (NCCL over
NVLink, or gloo in the CPU tests of the host logic) carries the two neighbour
exchanges of every substep -- the shared grid column after P2G and the
particles that crossed a cut after G2P -- as fixed-capacity messages, so that a
whole batch of substeps is enqueued without a host synchronisation
(include/mpm_b200.h, "multi-GPU slab decomposition").
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

INT_MIN, INT_MAX = -2**31, 2**31 - 1


class SlabDecomposition:
    """Host-side bookkeeping of the cuts: pure NumPy, no device needed."""

    def __init__(self, world, rank, cuts, leaf, grid_size, inv_dx):
        cuts = [int(c) for c in cuts]
        assert len(cuts) == world - 1 and all(b > a for a, b in zip(cuts, cuts[1:])), 'cuts must be strictly increasing'
        self.world, self.rank, self.cuts = world, rank, cuts
        self.leaf, self.grid_size, self.inv_dx = leaf, grid_size, inv_dx
        self.lo = cuts[rank - 1] if rank > 0 else INT_MIN
        self.hi = cuts[rank] if rank < world - 1 else INT_MAX
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None

    def block_x(self, x):
        """Absolute leaf-block x of particle positions (same f32 arithmetic as the binning kernel)."""
        x = np.asarray(x, np.float32)
        base = np.floor(x * np.float32(self.inv_dx) - np.float32(0.5)).astype(np.int64)
        return (base + self.grid_size // 2) // self.leaf

    def owner(self, x):
        return np.searchsorted(np.asarray(self.cuts, np.int64), self.block_x(x), side='right')

    def mine(self, x):
        return self.owner(x) == self.rank

    @staticmethod
    def balanced_cuts(x, world, leaf, grid_size, inv_dx):
        """Cut planes (block units) that give every rank about the same number of particles."""
        x = np.asarray(x, np.float32)
        base = np.floor(x * np.float32(inv_dx) - np.float32(0.5)).astype(np.int64)
        bx = np.sort((base + grid_size // 2) // leaf)
        cuts = []
        for k in range(1, world):
            target = (len(bx) * k) // world
            c = int(bx[min(len(bx) - 1, target)])
            # a cut at c leaves count(bx < c) particles on the left: take c or c + 1, whichever is closer
            below = int(np.searchsorted(bx, c, side='left'))
            below_next = int(np.searchsorted(bx, c + 1, side='left'))
            if abs(below_next - target) < abs(below - target):
                c += 1
            if cuts and c <= cuts[-1]:
                c = cuts[-1] + 1
            cuts.append(c)
        return cuts

    @staticmethod
    def cost_balanced_cuts(segments, world, fallback):
        """Cut planes that give every rank the same share of a measured cost.  `segments` = per rank, in rank order,
        (first block column, one past the last, cost): the cost of a rank is taken as spread evenly over the columns its
        particles occupy.  Ranks without particles or cost are skipped; with nothing to go by the `fallback` cuts stay."""
        segs = [(float(x0), float(x1), float(c)) for x0, x1, c in segments if x1 > x0 and c > 0]
        total = sum(c for _, _, c in segs)
        if not segs or total <= 0:
            return list(fallback)
        cuts, acc, k = [], 0.0, 1
        for x0, x1, c in segs:
            while k < world and acc + c >= total * k / world:
                frac = (total * k / world - acc) / c
                cuts.append(int(round(x0 + frac * (x1 - x0))))
                k += 1
            acc += c
        while len(cuts) < world - 1:
            cuts.append(cuts[-1] + 1 if cuts else int(segs[-1][1]))
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1] + 1)
        return cuts

    @staticmethod
    def uniform_cuts(x_lo, x_hi, world, leaf, grid_size, inv_dx):
        """Equal-width slabs over [x_lo, x_hi), snapped to leaf-block boundaries."""
        half = grid_size // 2
        b_lo = (int(np.floor(x_lo * inv_dx)) + half) // leaf
        b_hi = (int(np.ceil(x_hi * inv_dx)) + half + leaf - 1) // leaf
        cuts = [b_lo + int(round((b_hi - b_lo) * k / world)) for k in range(1, world)]
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1] + 1)
        return cuts


def neighbour_exchange(send_lo, send_hi, recv_lo, recv_hi, left, right, group=None):
    """Send `send_lo` to the left rank and `send_hi` to the right one, receive their
    counterparts.  Tensors may live on the GPU (NCCL) or the CPU (gloo).  Ends of the
    chain pass None for the missing side."""
    ops = []
    if left is not None:
        ops.append(dist.P2POp(dist.isend, send_lo, left, group))
        ops.append(dist.P2POp(dist.irecv, recv_lo, left, group))
    if right is not None:
        ops.append(dist.P2POp(dist.isend, send_hi, right, group))
        ops.append(dist.P2POp(dist.irecv, recv_hi, right, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def _make_distributed_solver():
    from .engine.mpm_solver import MPMSolver

    class DistributedMPMSolver(MPMSolver):
        """`MPMSolver` whose particles are spread over the ranks of a process group.

        Every rank constructs it with the same arguments and makes the same add_* calls
        with the same data; each keeps the particles of its slab.  `cuts` are the
        world-1 cut planes in absolute leaf-block units (see SlabDecomposition)."""

        def __init__(self, res, cuts, group=None, mig_capacity=1 << 16, halo_capacity=1 << 12, substep_batch=8,
                     world=None, rank=None, comm='auto', **kw):
            if kw.get('use_g2p2g'):
                raise NotImplementedError('use_g2p2g is single-GPU in this build')
            if kw.get('quant') and len(res) == 3:
                raise NotImplementedError('quant=True (bit-packed storage) is single-GPU in this build')
            super().__init__(res, **kw)
            self.group = group
            # 'peer': kernels write into the neighbour's buffers over NVLink (CUDA IPC), no NCCL per
            # substep; 'nccl': packed buffers travel by torch.distributed send/recv; 'auto': peer when a
            # process group with the nccl backend is up, else nccl/gloo
            self.comm = comm
            # world/rank may be given explicitly (single-process loop-back tests drive several slabs by hand)
            self.world = dist.get_world_size(group) if world is None else world
            self.rank = dist.get_rank(group) if rank is None else rank
            self.slab = SlabDecomposition(self.world, self.rank, cuts, self.leaf_block_size, self.grid_size,
                                          self.inv_dx)
            self.substep_batch = substep_batch
            self._global_n = 0
            self._global_ids = True      # `id` holds global ids: the insertion-order read-backs are disabled
            self._next_box = None
            self._mig_cap, self._halo_cap = int(mig_capacity), int(halo_capacity)
            lib, dev = self._lib, self._device
            mig_bytes = lib.mpm_comm_bytes(self.dim, 0, self._mig_cap)
            halo_bytes = lib.mpm_comm_bytes(self.dim, 1, self._halo_cap)
            z = lambda nbytes: torch.zeros(nbytes // 4, dtype=torch.int32, device=dev)
            s = self.slab
            self._mig_send = [z(mig_bytes) if s.left is not None else None, z(mig_bytes) if s.right is not None else None]
            self._mig_recv = [z(mig_bytes) if s.left is not None else None, z(mig_bytes) if s.right is not None else None]
            self._halo_send = [z(halo_bytes) if s.left is not None else None,
                               z(halo_bytes) if s.right is not None else None]
            self._halo_recv = [z(halo_bytes) if s.left is not None else None,
                               z(halo_bytes) if s.right is not None else None]
            ptr = lambda t: t.data_ptr() if t is not None else None
            self._check(lib.mpm_set_slab(self._ctx, 1, max(s.lo, INT_MIN), min(s.hi, INT_MAX)), 'mpm_set_slab')
            self._check(
                lib.mpm_bind_comm(self._ctx, ptr(self._mig_send[0]), ptr(self._mig_send[1]), self._mig_cap,
                                  ptr(self._halo_send[0]), ptr(self._halo_send[1]), self._halo_cap), 'mpm_bind_comm')
            if self.comm == 'auto':
                use_peer = (world is None and self.world > 1 and dist.is_initialized()
                            and dist.get_backend(group) == 'nccl')
                self.comm = 'peer' if use_peer else 'nccl'
            if self.comm == 'peer':
                self._setup_peer()

        def _setup_peer(self):
            """Exchange CUDA IPC handles of the receive regions and map the neighbours' ones.  The handles travel by
            any backend (NCCL, or gloo when several ranks share one GPU in the tests); the data path is peer memory."""
            lib, ctx = self._lib, self._ctx
            self._check(lib.mpm_peer_alloc(ctx, self._mig_cap, self._halo_cap), 'mpm_peer_alloc')
            h = (ctypes.c_uint8 * 64)()
            self._check(lib.mpm_peer_handle(ctx, h), 'mpm_peer_handle')
            cdev = self._coll_device()
            mine = torch.tensor(list(h), dtype=torch.uint8, device=cdev)
            allh = [torch.empty(64, dtype=torch.uint8, device=cdev) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=self.group)
            for side, peer in ((0, self.slab.left), (1, self.slab.right)):
                if peer is not None:
                    buf = (ctypes.c_uint8 * 64)(*allh[peer].cpu().tolist())
                    self._check(lib.mpm_peer_open(ctx, side, buf), 'mpm_peer_open')
            dist.barrier(group=self.group)

        # ---- collectives that work on any backend (NCCL: device tensors, gloo: host tensors) ----
        def _coll_device(self):
            if dist.is_initialized() and dist.get_backend(self.group) == 'gloo':
                return torch.device('cpu')
            return self._device

        def _allreduce_ints(self, values, op):
            t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=self._coll_device())
            if self.world > 1 and dist.is_initialized():
                dist.all_reduce(t, op=op, group=self.group)
            return [int(v) for v in t.tolist()]

        # ---- seeding: same call on every rank, each keeps its slab -------------
        _SLICE = 1 << 22     # rows selected per call of mpm_seed_positions_slab (it borrows the binning scratch)

        def _add_device_positions(self, dev, id_base, material, color, velocity, emitter=0):
            """Append the rows of `dev` ((n, dim) f32 on the device) that fall into this rank's slab; row i of `dev`
            carries the global id id_base + i.  Selection, block-sorted storage order and the count come from the
            library (mpm_seed_positions_slab: one radix sort), no eager tensor ops."""
            n_all = int(dev.shape[0])
            kept_total = 0
            self._next_box = None
            for lo in range(0, n_all, self._SLICE):
                part = dev[lo:lo + self._SLICE]
                m = int(part.shape[0])
                self._reserve(m)        # worst case: every row is mine (and the sort needs m <= capacity)
                kept = ctypes.c_int64()
                self._check(
                    self._lib.mpm_seed_positions_slab(self._ctx, part.data_ptr(), m, int(id_base) + lo, int(material),
                                                      int(color), self._vec(velocity), int(emitter),
                                                      ctypes.byref(kept), self._stream()), 'mpm_seed_positions_slab')
                self._n += int(kept.value)
                kept_total += int(kept.value)
            return kept_total

        def add_particles(self, particles, material, color=0xFFFFFF, velocity=None):
            """Same call with the same data on every rank (like every add_* of this class): each rank keeps the rows
            whose base block lies in its slab; global ids = position in the call sequence."""
            particles = np.ascontiguousarray(np.asarray(particles, dtype=np.float32))
            assert particles.ndim == 2 and particles.shape[1] == self.dim
            self.set_source_velocity(velocity)
            n_all = len(particles)
            if n_all == 0:
                return
            with torch.cuda.device(self._device):
                for lo in range(0, n_all, self._SLICE):      # bounded staging: never the whole array on the device
                    dev = torch.from_numpy(particles[lo:lo + self._SLICE]).to(self._device)
                    self._add_device_positions(dev, self._global_n + lo, material, color, velocity)
            self._global_n += n_all

        def add_local_particles(self, particles, material, color=0xFFFFFF, velocity=None, id_base=0):
            """Pre-partitioned seeding: `particles` are rows of THIS rank only (rows outside its slab are dropped),
            row i gets the global id id_base + i; the caller keeps the id ranges of the ranks disjoint.  Large runs
            use this so that no rank ever touches another rank's rows."""
            particles = np.ascontiguousarray(np.asarray(particles, dtype=np.float32))
            assert particles.ndim == 2 and particles.shape[1] == self.dim
            self.set_source_velocity(velocity)
            kept = 0
            with torch.cuda.device(self._device):
                for lo in range(0, len(particles), self._SLICE):
                    dev = torch.from_numpy(particles[lo:lo + self._SLICE]).to(self._device)
                    kept += self._add_device_positions(dev, int(id_base) + lo, material, color, velocity)
            self._global_n = max(self._global_n, int(id_base) + len(particles))
            return kept

        def _add_generated(self, mode, num, a3, b3, material, color, velocity):
            """add_cube / add_ellipsoid: every rank generates the SAME points as the single-device solver would
            (counter-based generator keyed by the global id) slice by slice on the device and keeps its slab."""
            seed = self._next_seed()
            self.set_source_velocity(velocity)
            with torch.cuda.device(self._device):
                for lo in range(0, num, self._SLICE):
                    m = min(self._SLICE, num - lo)
                    x = torch.empty((m, self.dim), dtype=torch.float32, device=self._device)
                    self._check(
                        self._lib.mpm_seed_generate(self._ctx, mode, m, self._global_n + lo, self._vec(a3), self._vec(b3),
                                                    seed, x.data_ptr(), self._stream()), 'mpm_seed_generate')
                    self._add_device_positions(x, self._global_n + lo, material, color, velocity)
            self._global_n += num

        def add_cube(self, lower_corner, cube_size, material, color=0xFFFFFF, sample_density=None, velocity=None):
            if sample_density is None:
                sample_density = 2**self.dim
            vol = 1
            for i in range(self.dim):
                vol = vol * cube_size[i]
            num = int(sample_density * vol / self.dx**self.dim + 1)              # reference :873
            assert self._global_n + num <= self.max_num_particles
            self._add_generated(1, num, lower_corner, cube_size, material, color, velocity)

        def add_ellipsoid(self, center, radius, material, color=0xFFFFFF, sample_density=None, velocity=None):
            import math
            import numbers
            if sample_density is None:
                sample_density = 2**self.dim
            if isinstance(radius, numbers.Number):
                radius = [radius] * self.dim
            radius = list(radius)
            num = math.pi if self.dim == 2 else 4 / 3 * math.pi
            for i in range(self.dim):
                num *= radius[i] * self.inv_dx
            num = int(math.ceil(num * sample_density))                            # reference :997-1005
            assert self._global_n + num <= self.max_num_particles
            self._add_generated(2, num, center, radius, material, color, velocity)

        def add_mesh(self, triangles, material, color=0xFFFFFF, sample_density=None, velocity=None, translation=None,
                     emmiter_id=0):
            assert self.dim == 3
            if sample_density is None:
                sample_density = 2**self.dim
            self.set_source_velocity(velocity)
            if self.voxelizer is None:
                raise RuntimeError('add_mesh needs use_voxelizer=True')
            self.voxelizer.voxelize(triangles)          # every rank rasterises the (small) mesh; each keeps its slab
            pos = self.voxelizer.sample_particles(sample_density=sample_density, translation=translation,
                                                  grid_size=self.grid_size, seed=self._next_seed())
            n = int(pos.shape[0])
            if n:
                with torch.cuda.device(self._device):
                    self._add_device_positions(pos, self._global_n, material, color, velocity, emitter=emmiter_id)
            self._global_n += n

        def add_ngon(self, *a, **k):
            MPMSolver.add_ngon(self, *a, **k)           # host-generated points -> add_particles (same on every rank)

        def read_restart(self, num_particles, pos, vel, material, color):
            raise NotImplementedError('restart a distributed run through add_particles per material')

        def clear_particles(self):
            if self.comm == 'peer':
                self.flush_migration()       # collective: drain what the last substep published to the neighbours
            super().clear_particles()
            self._global_n = 0
            self._next_box = None
            for t in self._mig_send:
                if t is not None:
                    t[0] = 0

        # ---- stepping -------------------------------------------------------------
        def _global_box(self):
            """Base-cell bounding box of the particles of ALL ranks: one all-reduce (MAX of (-lo, hi))."""
            lo, hi = self._local_box()
            r = self._allreduce_ints([-v for v in lo] + list(hi), dist.ReduceOp.MAX)
            return [-v for v in r[:3]], r[3:]

        def _local_box(self):
            lo = (ctypes.c_int32 * 3)()
            hi = (ctypes.c_int32 * 3)()
            self._check(self._lib.mpm_get_bbox(self._ctx, lo, hi, self._stream()), 'mpm_get_bbox')
            return list(lo), list(hi)

        def _batch_begin(self, glo, ghi):
            lib, ctx = self._lib, self._ctx
            # static rows (colour, id, emitter by sid): departures leave holes, arrivals append -- renumber them by
            # storage slot once the holes exceed one message capacity (a cheap pass between batches)
            ns = ctypes.c_int64()
            lib.mpm_get_static_rows(ctx, ctypes.byref(ns))
            if ns.value - self._n > self._mig_cap:
                self._check(lib.mpm_compact_statics(ctx, self._stream()), 'mpm_compact_statics')
                ns.value = self._n
            need = max(self._n, ns.value) + 2 * self._mig_cap     # rows for the particles that may arrive
            if need > self._cap:
                self._rebind(capacity=max(need, int(self._cap * 1.25)))
            self._check(lib.mpm_set_layout_box(ctx, 1, (ctypes.c_int32 * 3)(*glo), (ctypes.c_int32 * 3)(*ghi)),
                        'mpm_set_layout_box')
            self._check(lib.mpm_batch_begin(ctx, self._stream()), 'mpm_batch_begin')

        def _substep_pre(self, dt):
            """After the migration buffers were exchanged: append arrivals, bin, P2G, pack the shared columns."""
            lib, ctx = self._lib, self._ctx
            ptr = lambda t: t.data_ptr() if t is not None else None
            self._check(lib.mpm_phase_unpack(ctx, ptr(self._mig_recv[0]), ptr(self._mig_recv[1]), self._stream()),
                        'mpm_phase_unpack')
            self._check(lib.mpm_phase_p2g(ctx, dt, self._stream()), 'mpm_phase_p2g')
            self._check(lib.mpm_phase_halo_pack(ctx, self._stream()), 'mpm_phase_halo_pack')

        def _substep_post(self, dt):
            """After the halo buffers were exchanged: add the neighbours' sums, grid op, G2P (packs leavers)."""
            lib, ctx = self._lib, self._ctx
            ptr = lambda t: t.data_ptr() if t is not None else None
            self._check(lib.mpm_phase_halo_add(ctx, ptr(self._halo_recv[0]), ptr(self._halo_recv[1]),
                                               self._stream()), 'mpm_phase_halo_add')
            self._check(lib.mpm_phase_g2p(ctx, dt, self._stream()), 'mpm_phase_g2p')

        def _batch_end(self):
            rc = self._lib.mpm_batch_end(self._ctx, self._stream())
            n = ctypes.c_int64()
            self._lib.mpm_get_state(self._ctx, None, ctypes.byref(n))
            self._n = int(n.value)
            return rc

        def _exchange_migration(self):
            s = self.slab
            neighbour_exchange(self._mig_send[0], self._mig_send[1], self._mig_recv[0], self._mig_recv[1], s.left,
                               s.right, self.group)

        def _exchange_halo(self):
            s = self.slab
            neighbour_exchange(self._halo_send[0], self._halo_send[1], self._halo_recv[0], self._halo_recv[1], s.left,
                               s.right, self.group)

        def _run_substeps(self, dt, count):
            """`count` substeps in batches of `substep_batch`.  Per batch the host does ONE synchronisation
            (mpm_batch_end) and ONE all-reduce that carries both the agreement on success and the bounding box the
            next batch's key layout needs (G2P measured it on the device)."""
            left = count
            box = getattr(self, '_next_box', None)
            self._next_box = None
            while left > 0:
                nb = min(left, max(1, self.substep_batch))
                glo, ghi = box if box is not None else self._global_box()
                if glo[0] > ghi[0]:
                    return self.stats()          # no particles anywhere
                self._batch_begin(glo, ghi)
                if box is None:
                    # particles were added or moved since the last batch: size the block workspace from a dry run of
                    # the block discovery (a capacity miss inside a batch cannot be retried, the neighbours run ahead)
                    need = ctypes.c_int32()
                    self._check(self._lib.mpm_batch_probe(self._ctx, ctypes.byref(need), self._stream()), 'mpm_batch_probe')
                    if need.value * 3 > self._max_blocks * 2:
                        self._batch_end()
                        self._rebind(max_blocks=2 * need.value)
                        self._batch_begin(glo, ghi)
                if self.comm == 'peer':
                    self._check(self._lib.mpm_peer_substeps(self._ctx, dt, nb, 0, self._stream()),
                                'mpm_peer_substeps')
                else:
                    for _ in range(nb):
                        self._exchange_migration()
                        self._substep_pre(dt)
                        self._exchange_halo()
                        self._substep_post(dt)
                rc = self._batch_end()
                st = self.stats()
                have = self._n > 0 and st.bbox_min[0] <= st.bbox_max[0]
                lo = list(st.bbox_min) if have else [INT_MAX] * 3
                hi = list(st.bbox_max) if have else [INT_MIN] * 3
                # every rank must agree on success before the next batch is enqueued
                r = self._allreduce_ints([abs(rc)] + [-v for v in lo] + hi, dist.ReduceOp.MAX)
                worst = r[0]
                if rc != 0 or worst != 0:
                    raise _lib.MPMError(f'distributed batch failed (local rc {rc}, worst {worst}): '
                                        + self._lib.mpm_last_error(self._ctx).decode()
                                        + ' -- capacities (reserve_blocks, mig_capacity, halo_capacity) cannot be '
                                          'grown inside a distributed batch')
                box = ([-v for v in r[1:4]], r[4:7])
                left -= nb
                need = max(st.n_particle_blocks, st.n_grid_blocks)      # keep a margin of 1.5x for the next batch
                if need * 3 > self._max_blocks * 2:
                    self._rebind(max_blocks=2 * need)
            self._next_box = box                  # valid until particles are added or cleared
            return self.stats()

        def _advance(self, dt, count, smry_writer):
            """A substep is a collective (neighbour exchanges, the global box): a rank that holds no particle right
            now must still take part -- it may be about to receive some."""
            st = self._run_substeps(dt, count)
            cur_frame_velocity = self._allmax_float(float(st.max_velocity) if self._n > 0 else 0.0)
            self.t += dt * count
            if smry_writer is not None:
                smry_writer.add_scalar("substep_max_CFL", cur_frame_velocity * dt / self.dx, self.total_substeps)
            self.all_time_max_velocity = max(self.all_time_max_velocity, cur_frame_velocity)

        def _allmax_float(self, v):
            t = torch.tensor([float(v)], dtype=torch.float64, device=self._coll_device())
            if self.world > 1 and dist.is_initialized():
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            return float(t.item())

        def compute_max_velocity(self):
            return self._allmax_float(MPMSolver.compute_max_velocity(self))

        def compute_max_grid_velocity(self, grid_v=None):
            """Global maximum: with use_adaptive_dt every rank must derive the same dt (reference :762-770)."""
            return self._allmax_float(MPMSolver.compute_max_grid_velocity(self))

        # ---- re-cutting ------------------------------------------------------------------
        def _exchange_buffers(self, send, recv):
            """Neighbour exchange of whole message buffers through torch.distributed (NCCL: device tensors; gloo -- several
            ranks on one GPU in the tests -- through host copies).  Only the rare bulk moves use it."""
            s = self.slab
            if dist.get_backend(self.group) == 'gloo':
                hs = [t.cpu() if t is not None else None for t in send]
                hr = [torch.empty_like(t, device='cpu') if t is not None else None for t in recv]
                neighbour_exchange(hs[0], hs[1], hr[0], hr[1], s.left, s.right, self.group)
                for d, h in zip(recv, hr):
                    if d is not None:
                        d.copy_(h)
            else:
                torch.cuda.current_stream(self._device).synchronize()
                neighbour_exchange(send[0], send[1], recv[0], recv[1], s.left, s.right, self.group)

        def rebalance(self, cuts):
            """Collective: move the cut planes to `cuts` (world - 1 absolute leaf-block x, the same on every rank) and
            move the particles that change owner, in rounds of at most `mig_capacity` rows per side (a particle may hop
            over several ranks).  Between batches only."""
            cuts = [int(c) for c in cuts]
            self.flush_migration()                       # nothing of the last substep may still be in flight
            out3 = (ctypes.c_int64 * 3)()                # rows outside the OLD columns were delivered: retire them
            self._check(self._lib.mpm_rebalance_pack(self._ctx, None, None, 0, out3, self._stream()), 'mpm_rebalance_pack')
            self.slab = SlabDecomposition(self.world, self.rank, cuts, self.leaf_block_size, self.grid_size, self.inv_dx)
            s = self.slab
            self._check(self._lib.mpm_set_slab(self._ctx, 1, max(s.lo, INT_MIN), min(s.hi, INT_MAX)), 'mpm_set_slab')
            ptr = lambda t: t.data_ptr() if t is not None else None
            moved = 0
            for _ in range(10000):
                glo, ghi = self._global_box()
                if glo[0] > ghi[0]:
                    break
                self._batch_begin(glo, ghi)
                out3 = (ctypes.c_int64 * 3)()
                self._check(self._lib.mpm_rebalance_pack(self._ctx, ptr(self._mig_send[0]), ptr(self._mig_send[1]),
                                                         self._mig_cap, out3, self._stream()), 'mpm_rebalance_pack')
                self._exchange_buffers(self._mig_send, self._mig_recv)
                self._check(self._lib.mpm_phase_unpack(self._ctx, ptr(self._mig_recv[0]), ptr(self._mig_recv[1]),
                                                       self._stream()), 'mpm_phase_unpack')
                if self._batch_end() != 0:
                    raise _lib.MPMError('rebalance: ' + self._lib.mpm_last_error(self._ctx).decode())
                for t in self._mig_send:
                    if t is not None:
                        t[0] = 0
                sent, left = self._allreduce_ints([out3[0] + out3[1], out3[2]], dist.ReduceOp.SUM)
                moved += sent
                if sent == 0 and left == 0:
                    break
            self._next_box = None
            return moved

        def balance_cuts(self, cost):
            """New cut planes that would give every rank the same share of `cost` (this rank's measured cost of a substep,
            e.g. its P2G + G2P time), assuming the cost is spread evenly over the block columns a rank's particles occupy.
            Collective; returns the same list on every rank."""
            lo, hi = self._local_box()
            half, leaf = self.grid_size // 2, self.leaf_block_size
            have = self._n > 0 and lo[0] <= hi[0]
            b0 = (lo[0] + half) // leaf if have else 0
            b1 = (hi[0] + half) // leaf + 1 if have else 0
            b0 = max(b0, self.slab.lo) if have else 0
            b1 = min(b1, self.slab.hi) if have else 0
            t = torch.tensor([float(cost) if have else 0.0, float(b0), float(b1)], dtype=torch.float64, device=self._coll_device())
            allt = [torch.empty_like(t) for _ in range(self.world)]
            if self.world > 1 and dist.is_initialized():
                dist.all_gather(allt, t, group=self.group)
            else:
                allt = [t]
            segs = [(float(a[1]), float(a[2]), float(a[0])) for a in allt]
            return SlabDecomposition.cost_balanced_cuts(segs, self.world, list(self.slab.cuts))

        def reserve_blocks(self, max_blocks):
            """Pre-size the leaf-block workspace (a capacity miss cannot be retried inside a distributed batch)."""
            if max_blocks > self._max_blocks:
                self._rebind(max_blocks=max_blocks)

        def flush_migration(self, exchange=None):
            """Deliver the particles packed by the last substep (call before reading particles back)."""
            glo, ghi = self._global_box() if exchange is None else exchange('box', self)
            if glo[0] > ghi[0]:
                return
            self._batch_begin(glo, ghi)
            if self.comm == 'peer':
                self._check(self._lib.mpm_peer_substeps(self._ctx, 0.0, 0, 1, self._stream()), 'mpm_peer_substeps')
                if self._batch_end() != 0:
                    raise _lib.MPMError('flush_migration: ' + self._lib.mpm_last_error(self._ctx).decode())
                return
            if exchange is None:
                self._exchange_migration()
            else:
                exchange('migration', self)
            ptr = lambda t: t.data_ptr() if t is not None else None
            self._check(self._lib.mpm_phase_unpack(self._ctx, ptr(self._mig_recv[0]), ptr(self._mig_recv[1]),
                                                   self._stream()), 'mpm_phase_unpack')
            for t in self._mig_send:
                if t is not None:
                    t[0] = 0
            if self._batch_end() != 0:
                raise _lib.MPMError('flush_migration: ' + self._lib.mpm_last_error(self._ctx).decode())

        # ---- read-back ----------------------------------------------------------------
        def particle_info(self):
            """MPMSolver.particle_info() (ref engine/mpm_solver.py:1172-1180) for THIS rank's particles, in storage
            order, plus their global ids ('id').  Call flush_migration() first (arrivals must have been appended);
            rows already handed to a neighbour are skipped.  One library kernel compacts the owned rows
            (mpm_export_local), one copy into pinned host memory."""
            n, d = self._n, self.dim
            cnt = ctypes.c_int64()
            names = (('position', d, np.float32), ('velocity', d, np.float32), ('material', 1, np.int32),
                     ('color', 1, np.int32), ('id', 1, np.int32))
            out = {}
            with torch.cuda.device(self._device):
                dev = torch.empty((max(n, 1) * (2 * d + 3), ), dtype=torch.int32, device=self._device)
                self._check(self._lib.mpm_export_local(self._ctx, dev.data_ptr(), ctypes.byref(cnt), self._stream()),
                            'mpm_export_local')
                k, off = int(cnt.value), 0
                for name, width, dt in names:          # every field block goes straight into its own pinned array
                    host = torch.empty((k * width, ), dtype=torch.int32, pin_memory=k > 0)
                    if k:
                        host.copy_(dev[off:off + k * width])
                    a = host.numpy().view(dt)
                    out[name] = a.reshape(k, width) if width > 1 else a
                    off += n * width
            return out

        # The insertion-order read-back paths of MPMSolver index device buffers by `id`, which is a permutation of
        # [0, n) only on a single-device solver.  Here ids are global: gather by id across the ranks instead.
        def gather_particle_info(self):
            """particle_info() of ALL ranks ordered by global id, on every rank (collective)."""
            mine = self.particle_info()
            parts = [None] * self.world
            if self.world > 1 and dist.is_initialized():
                dist.all_gather_object(parts, mine, group=self.group)
            else:
                parts = [mine]
            merged = {k: np.concatenate([p[k] for p in parts]) for k in mine}
            order = np.argsort(merged['id'], kind='stable')
            return {k: v[order] for k, v in merged.items()}

        def _pack_particles(self):
            return None            # ParticleIO falls back to copy_ranged, which gathers (below)

        def write_particles(self, fn, slice_size=1000000):
            """Collective: the particles of all ranks in global-id order, written by rank 0 in the reference's
            .npz format (ref engine/particle_io.py:12-76)."""
            info = self.gather_particle_info()
            if self.rank == 0:
                from .engine.particle_io import ParticleIO
                ParticleIO.write_arrays(fn, info['position'], info['velocity'], info['color'])

        def write_particles_ply(self, fn):
            info = self.gather_particle_info()
            if self.rank == 0:
                data = np.hstack([info['position'], info['color'].astype(np.uint32)[:, None].view(np.float32)])
                from .engine.mesh_io import write_point_cloud
                write_point_cloud(fn, data)

        def _no_insertion_order(self, *a, **k):
            raise NotImplementedError('insertion-order field read-back is single-device; use particle_info() '
                                      '(this rank) or gather_particle_info() (all ranks, by global id)')

        copy_ranged = copy_ranged_nd = copy_dynamic = copy_dynamic_nd = debug_binning = _no_insertion_order

        def local_rows(self):
            """This rank's live particles: dict of arrays in storage order, with global ids.
            Rows of particles that have left the slab since the last substep are excluded."""
            n = self._n
            nv = self._lib.mpm_virtual_fields(self.dim)      # x v F C Jp material color id emitter
            out = np.empty((nv, n), np.int32)
            for f in range(nv):
                if n:
                    self._check(
                        self._lib.mpm_download_raw(self._ctx, f, out[f].ctypes.data_as(ctypes.c_void_p),
                                                   self._stream()), 'mpm_download_raw')
            d, dd = self.dim, self.dim * self.dim
            fl = out.view(np.float32)
            rows = {
                'x': fl[0:d].T.copy(), 'v': fl[d:2 * d].T.copy(),
                'F': fl[2 * d:2 * d + dd].T.reshape(n, d, d).copy(),
                'C': fl[2 * d + dd:2 * d + 2 * dd].T.reshape(n, d, d).copy(),
                'Jp': fl[2 * d + 2 * dd].copy(), 'material': out[2 * d + 2 * dd + 1].copy(),
                'color': out[2 * d + 2 * dd + 2].copy(), 'id': out[2 * d + 2 * dd + 3].copy(),
            }
            keep = (self.slab.mine(rows['x'][:, 0]) & (rows['material'] != 7)) if n else np.zeros(0, bool)   # 7: retired by rebalance
            return {k: v[keep] for k, v in rows.items()}

        def gather_rows(self):
            """All particles of all ranks, ordered by global id, on every rank (tests / export)."""
            mine = self.local_rows()
            parts = [None] * self.world
            dist.all_gather_object(parts, mine, group=self.group)
            merged = {k: np.concatenate([p[k] for p in parts]) for k in mine}
            order = np.argsort(merged['id'], kind='stable')
            return {k: v[order] for k, v in merged.items()}

    return DistributedMPMSolver


def __getattr__(name):
    if name == 'DistributedMPMSolver':
        cls = _make_distributed_solver()
        globals()['DistributedMPMSolver'] = cls
        return cls
    raise AttributeError(name)
