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
            """Exchange CUDA IPC handles of the receive regions and map the neighbours' ones."""
            lib, ctx = self._lib, self._ctx
            self._check(lib.mpm_peer_alloc(ctx, self._mig_cap, self._halo_cap), 'mpm_peer_alloc')
            h = (ctypes.c_uint8 * 64)()
            self._check(lib.mpm_peer_handle(ctx, h), 'mpm_peer_handle')
            mine = torch.tensor(list(h), dtype=torch.uint8, device=self._device)
            allh = [torch.empty(64, dtype=torch.uint8, device=self._device) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=self.group)
            for side, peer in ((0, self.slab.left), (1, self.slab.right)):
                if peer is not None:
                    buf = (ctypes.c_uint8 * 64)(*allh[peer].cpu().tolist())
                    self._check(lib.mpm_peer_open(ctx, side, buf), 'mpm_peer_open')
            dist.barrier(group=self.group)

        # ---- seeding: same call on every rank, each keeps its slab -------------
        def add_particles(self, particles, material, color=0xFFFFFF, velocity=None):
            """Same call on every rank; each keeps the rows whose base block lies in its slab.  The rows are
            uploaded once and selected on the device (same f32 arithmetic as the binning kernel and as
            SlabDecomposition.block_x), global ids = position in the call sequence."""
            particles = np.ascontiguousarray(np.asarray(particles, dtype=np.float32))
            assert particles.ndim == 2 and particles.shape[1] == self.dim
            n_all = len(particles)
            if n_all == 0:
                return
            with torch.cuda.device(self._device):
                dev = torch.from_numpy(particles).to(self._device)
                base = torch.floor(dev[:, 0] * np.float32(self.inv_dx) - np.float32(0.5)).to(torch.int64)
                bx = torch.div(base + self.grid_size // 2, self.leaf_block_size, rounding_mode='floor')
                keep = (bx >= self.slab.lo) & (bx < self.slab.hi)
                cnt = int(keep.sum().item())
                if cnt == n_all:
                    ids = torch.arange(self._global_n, self._global_n + n_all, dtype=torch.int32, device=self._device)
                else:
                    idx = torch.nonzero(keep).squeeze(1)
                    dev = dev.index_select(0, idx).contiguous()
                    ids = (idx + self._global_n).to(torch.int32)
                if cnt >= (1 << 15) and self.grid_size == 4096:
                    # store the rows sorted by leaf block (ids keep the call order), as the single-GPU solver does for
                    # large arrays: the first substep then reads block-local rows instead of gathering at random
                    b = torch.floor(dev * np.float32(self.inv_dx) - np.float32(0.5)).to(torch.int64)
                    b = torch.div(b + self.grid_size // 2, self.leaf_block_size, rounding_mode='floor').clamp_(0, 1023)
                    key = b[:, 0]
                    for d in range(1, self.dim):
                        key = key * 1024 + b[:, d]
                    order = torch.argsort(key)
                    dev = dev.index_select(0, order).contiguous()
                    ids = ids.index_select(0, order)
                self._global_n += n_all
                n0 = self._n
                self._seed_from_device(dev, material, color, velocity)
                if cnt:
                    cur = ctypes.c_int32()
                    self._lib.mpm_get_state(self._ctx, ctypes.byref(cur), None)
                    id_row = self._nf - 2          # x v F C Jp material color id emitter
                    self._state[cur.value, id_row, n0:n0 + cnt] = ids

        def clear_particles(self):
            if self.comm == 'peer':
                self.flush_migration()       # collective: drain what the last substep published to the neighbours
            super().clear_particles()
            self._global_n = 0
            for t in self._mig_send:
                if t is not None:
                    t[0] = 0

        def add_cube(self, *a, **k):
            raise NotImplementedError('seed through add_particles on the distributed solver')

        add_ellipsoid = add_cube
        add_mesh = add_cube

        # ---- stepping -------------------------------------------------------------
        def _global_box(self):
            lo, hi = self._local_box()
            t_lo = torch.tensor(lo, dtype=torch.int64, device=self._device)
            t_hi = torch.tensor(hi, dtype=torch.int64, device=self._device)
            dist.all_reduce(t_lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(t_hi, op=dist.ReduceOp.MAX, group=self.group)
            return [int(v) for v in t_lo.tolist()], [int(v) for v in t_hi.tolist()]

        def _local_box(self):
            lo = (ctypes.c_int32 * 3)()
            hi = (ctypes.c_int32 * 3)()
            self._check(self._lib.mpm_get_bbox(self._ctx, lo, hi, self._stream()), 'mpm_get_bbox')
            return list(lo), list(hi)

        def _batch_begin(self, glo, ghi):
            need = self._n + 2 * self._mig_cap          # rows for the particles that may arrive
            if need > self._cap:
                self._rebind(capacity=max(need, int(self._cap * 1.25)))
            lib, ctx = self._lib, self._ctx
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
            left = count
            while left > 0:
                nb = min(left, max(1, self.substep_batch))
                glo, ghi = self._global_box()
                if glo[0] > ghi[0]:
                    return self.stats()          # no particles anywhere
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
                # every rank must agree on success before the next batch is enqueued
                flag = torch.tensor([abs(rc)], dtype=torch.int32, device=self._device)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
                worst = int(flag.item())
                if rc != 0 or worst != 0:
                    raise _lib.MPMError(f'distributed batch failed (local rc {rc}, worst {worst}): '
                                        + self._lib.mpm_last_error(self._ctx).decode()
                                        + ' -- capacities (reserve_blocks, mig_capacity, halo_capacity) cannot be '
                                          'grown inside a distributed batch')
                left -= nb
            return self.stats()

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
            order, plus their global ids ('id').  Call flush_migration() first: rows of particles that have been
            handed to a neighbour are excluded here (same f32 arithmetic as the binning), arrivals must have been
            appended.  One device-side selection, one copy into pinned host memory per field."""
            n, d = self._n, self.dim
            names = ('position', 'velocity', 'material', 'color', 'id')
            if n == 0:
                out = {'position': np.zeros((0, d), np.float32), 'velocity': np.zeros((0, d), np.float32)}
                out.update({k: np.zeros(0, np.int32) for k in names[2:]})
                return out
            cur = ctypes.c_int32()
            self._lib.mpm_get_state(self._ctx, ctypes.byref(cur), None)
            with torch.cuda.device(self._device):
                torch.cuda.current_stream(self._device).synchronize()
                st = self._state[cur.value]
                x0 = st[0, :n].view(torch.float32)
                base = torch.floor(x0 * np.float32(self.inv_dx) - np.float32(0.5)).to(torch.int64)
                bx = torch.div(base + self.grid_size // 2, self.leaf_block_size, rounding_mode='floor')
                keep = (bx >= self.slab.lo) & (bx < self.slab.hi)
                idx = torch.nonzero(keep).squeeze(1)
                jp = 2 * d + 2 * d * d
                rows = [st[0:d], st[d:2 * d], st[jp + 1:jp + 2], st[jp + 2:jp + 3], st[jp + 3:jp + 4]]
                out = {}
                for name, r in zip(names, rows):
                    sel = r[:, :n].index_select(1, idx).t().contiguous()          # (count, words)
                    host = torch.empty(sel.shape, dtype=torch.int32, pin_memory=True)
                    host.copy_(sel)
                    a = host.numpy()
                    out[name] = a.view(np.float32) if name in ('position', 'velocity') else a[:, 0]
            return out

        def local_rows(self):
            """This rank's live particles: dict of arrays in storage order, with global ids.
            Rows of particles that have left the slab since the last substep are excluded."""
            n = self._n
            out = np.empty((self._nf, n), np.int32)
            for f in range(self._nf):
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
            keep = self.slab.mine(rows['x'][:, 0]) if n else np.zeros(0, bool)
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
