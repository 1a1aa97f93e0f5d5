#!/usr/bin/env python
"""Benchmark of the MLS-MPM substep (BASELINE.json metric: particle-substeps/s).

    python bench.py --gpus N --steps K --warmup W             # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...    # CPU restatement of the reference

A "step" is ONE substep (binning -> P2G -> grid op -> G2P) over the whole particle set.

Default workload = BASELINE.json configs[3] (SURVEY.md 8(d) cfg 4): res 256^3, `unbounded=True`,
100 M particles in four equal x-slabs WATER / ELASTIC / SNOW / SAND over a slip floor, the slabs
moving against each other and towards the floor so that the timed window runs on deformed material
in contact.  N = 1: the whole scene on one B200.  N > 1: the SAME particle set split into N x-slabs
(strong scaling, the "1 vs 8" of configs[3]); the line also carries `weak_cfg5`, the configs[4]
brick (125 M particles per GPU, res 512 unbounded, WATER under SAND) timed at the same N.
`--workload cube_drop_4m` is configs[1].  Prints one JSON line (contract in the task statement).
"""
import argparse
import contextlib
import io
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PARTICLE = 220      # algorithmic bytes per particle-substep, 3D f32
B_CELL = 88           # algorithmic bytes per active grid cell
B_P2G_PARTICLE = 144  # P2G share: reads x,v,F,C,Jp,material (104) + writes F,Jp (40)
B_P2G_CELL = 32       # P2G grid read-modify-write
B_G2P2G_PARTICLE = 132  # use_g2p2g: reads x,v,F,Jp,material (68) + writes x,v,F,Jp (64)

WATER, ELASTIC, SNOW, SAND = 0, 1, 2, 3


# ------------------------------------------------------------------ workloads
class Chunk:
    """`n` particles uniform in a box given in CELL units (so that with dx a power of two the base cell
    floor(x/dx - 0.5) of every particle lies in [lo - 0.5, hi - 0.5) exactly): rank-independent, seeded per chunk."""

    def __init__(self, lo, hi, n, material, velocity, seed):
        self.lo, self.hi, self.n, self.material, self.velocity, self.seed = lo, hi, int(n), material, velocity, seed

    def positions(self, dx):
        rng = np.random.default_rng(self.seed)
        lo, hi = np.asarray(self.lo, np.float32), np.asarray(self.hi, np.float32)
        x = rng.random((self.n, 3), dtype=np.float32)
        x *= (hi - lo)
        x += lo
        np.minimum(x, np.nextafter(hi, -np.inf, dtype=np.float32), out=x)
        x *= np.float32(dx)
        return x


def multimat(name, z_cells=232, chunks_per_material=4, y_cells=120.27, side_walls=False):
    """configs[3] / SURVEY 8(d) cfg 4: four equal x-slabs WATER, ELASTIC, SNOW, SAND at 8 particles per cell over a
    slip floor (friction 0.5), res 256 unbounded.  Every slab is `chunks_per_material` chunks of 28 x-cells (7 leaf
    blocks); consecutive chunks move against each other (+-0.25 m/s in x) and everything falls at 1 m/s.  (Gentle on
    purpose: the reference's water, p = lambda J (J - 1), cannot stop an impact faster than ~4 m/s (|p| <= lambda / 4),
    and its snow hardens as exp(10 (1 - Jp)), which breaks the CFL bound of the default dt after ~25 % compaction.)  The
    materials are in contact with the floor and with each other and F is away from the identity everywhere after
    the pre-roll.  The full scene: x in [-0.875, 0.875], y in [0.008, 0.478], z in [-0.453, 0.453], 100 000 000
    particles.  `side_walls`: slip planes on the two z faces (the CPU sample is a thin z-slice of the scene)."""
    res, dx = 256, 1.0 / 256
    nchunk = 4 * chunks_per_material
    x0 = -28 * nchunk // 2
    y0 = 2.0                     # cells above the floor plane y = 0
    per_chunk = int(round(28 * y_cells * z_cells * 8))
    if name == 'multimat_100m':
        per_chunk = 6_250_000
    chunks = []
    for k in range(nchunk):
        lo = (x0 + 28 * k + 0.5, y0, -z_cells / 2 + 0.5)
        hi = (x0 + 28 * (k + 1) + 0.5, y0 + y_cells, z_cells / 2 + 0.5)
        mat = k // chunks_per_material
        if os.environ.get('MPM_BENCH_MATERIAL'):         # development: the whole scene of ONE material (cost per material)
            mat = int(os.environ['MPM_BENCH_MATERIAL'])
        chunks.append(Chunk(lo, hi, per_chunk, mat, (0.25 if k % 2 == 0 else -0.25, -1.0, 0.0), 4000 + k))
    colliders = [((0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 1, 0.5)]      # point, normal, surface (slip), friction
    if side_walls:
        colliders.append(((0.0, 0.0, (-z_cells / 2 + 0.5) * dx), (0.0, 0.0, 1.0), 1, 0.0))
        colliders.append(((0.0, 0.0, (z_cells / 2 + 0.5) * dx), (0.0, 0.0, -1.0), 1, 0.0))
    return dict(name=name, res=(res, ) * 3, unbounded=True, gravity=(0.0, -9.8, 0.0), frame_dt=3e-3, chunks=chunks,
                colliders=colliders, preroll=100, n=per_chunk * nchunk,
                cut_cells=[x0 + 28 * k for k in range(1, nchunk)],   # admissible cut planes (leaf-block aligned)
                label=f'configs[3]: 3D unbounded res 256^3, {per_chunk * nchunk} particles in four x-slabs '
                      f'WATER/ELASTIC/SNOW/SAND (8 per cell), slip floor mu=0.5, g=(0,-9.8,0), chunks of 28 cells '
                      f'colliding at +-0.25 m/s and falling at 1 m/s from 2 cells above the floor, dt=2e-2*dx')


def brick(name, world=1, cells=(248, 252, 250)):
    """configs[4] / SURVEY 8(d) cfg 5: res 512 unbounded, one brick per GPU (248 x 252 x 250 cells x 8 = 124 992 000
    particles), WATER below SAND, bricks tiled along x (cuts on leaf-block boundaries carry halo and migration),
    all fall at 1 m/s onto a slip floor 2 cells below (every brick does the same work as the single one at N = 1)."""
    res, dx = 512, 1.0 / 512
    cx, cy, cz = cells
    x0 = -cx * world // 2
    x0 -= x0 % 4
    y0 = 2.0
    chunks = []
    for r in range(world):
        vel = (0.0, -1.0, 0.0)
        for h, mat in enumerate((WATER, SAND)):
            lo = (x0 + cx * r + 0.5, y0 + h * cy / 2, -cz / 2 + 0.5)
            hi = (x0 + cx * (r + 1) + 0.5, y0 + (h + 1) * cy / 2, cz / 2 + 0.5)
            c = Chunk(lo, hi, cx * cy * cz * 4, mat, vel, 5000 + 2 * r + h)
            c.rank = r
            chunks.append(c)
    return dict(name=name, res=(res, ) * 3, unbounded=True, gravity=(0.0, -9.8, 0.0), frame_dt=1.5e-3, chunks=chunks,
                colliders=[((0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 1, 0.5)], preroll=40, n=cx * cy * cz * 8 * world,
                cut_cells=[x0 + cx * r for r in range(1, world)],
                label=f'configs[4]: 3D unbounded res 512^3, {cx * cy * cz * 8} particles per GPU (brick {cx}x{cy}x{cz} '
                      f'cells, 8 per cell, WATER below SAND), bricks tiled along x, slip floor, dt=2e-2*dx')


def cube_drop(name):
    """configs[1]: 3D cube drop, res 256^3 bounded, ELASTIC cube over WATER cube, 8 per cell, at rest."""
    side = {'cube_drop_4m': 0.25, 'cube_drop_sample': 0.16, 'cube_drop_small': 0.0625}[name]
    res = 256
    ns = int(round(side * res))
    lo_c = res // 2 - ns // 2
    chunks = []
    for k, (y, mat) in enumerate(((0.55, ELASTIC), (0.15, WATER))):
        lo = (lo_c, y * res, lo_c)
        hi = (lo_c + ns, y * res + ns, lo_c + ns)
        chunks.append(Chunk(lo, hi, ns**3 * 8, mat, (0.0, 0.0, 0.0), 2000 + k))
    return dict(name=name, res=(res, ) * 3, unbounded=False, gravity=(0.0, -20.0, 0.0), frame_dt=3e-3, chunks=chunks,
                colliders=[], preroll=0, n=2 * ns**3 * 8, cut_cells=[], dt=3e-3 / 39,
                label=f'configs[1]: 3D cube drop, res 256^3 bounded, {2 * ns**3 * 8} particles (ELASTIC cube over '
                      f'WATER cube, 8 per cell), g=(0,-20,0), dt=3e-3/39')


def bunnies(name, n_ellipsoids=482):
    """configs[2] / SURVEY 8(d) cfg 3: res 256 unbounded, g = (0, -25, 0), the six slip planes (mu = 0.5) of ref
    demo/demo_3d_bunnies.py:76-107, 482 ellipsoids r = 0.048 (62 175 particles each = 29.97 M) alternating SNOW / SAND on
    a jittered 0.2 lattice, launched downwards at 5 m/s; seeded on the device with add_ellipsoid (every particle takes
    the SVD path once the material deforms)."""
    rng = np.random.default_rng(3)
    ell = []
    for ix in np.arange(-1.6, 1.61, 0.2):
        for iy in np.arange(0.1, 1.51, 0.2):
            for iz in np.arange(-0.8, 0.81, 0.2):
                if len(ell) < n_ellipsoids:
                    c = np.array([ix, iy, iz]) + rng.uniform(-0.03, 0.03, 3)
                    ell.append((list(map(float, c)), 0.048, SNOW if len(ell) % 2 == 0 else SAND, (0.0, -5.0, 0.0)))
    planes = [((0, 0, 0), (0, 1, 0)), ((0, 1.9, 0), (0, -1, 0)), ((-1.9, 0, 0), (1, 0, 0)), ((1.9, 0, 0), (-1, 0, 0)),
              ((0, 0, -0.95), (0, 0, 1)), ((0, 0, 0.95), (0, 0, -1))]
    return dict(name=name, res=(256, ) * 3, unbounded=True, gravity=(0.0, -25.0, 0.0), frame_dt=3e-3, chunks=[],
                ellipsoids=ell, colliders=[(p, n, 1, 0.5) for p, n in planes], preroll=20, n=None, cut_cells=[],
                label=f'configs[2]: 3D unbounded res 256^3, {len(ell)} SNOW/SAND ellipsoids r=0.048 (add_ellipsoid, 8 per cell) '
                      f'falling at 5 m/s between six slip planes mu=0.5, g=(0,-25,0), dt=2e-2*dx')


def workload(name, world=1):
    w = _workload(name, world)
    base = 0
    for c in w['chunks']:          # global particle ids = position in the chunk sequence
        c.id_base = base
        base += c.n
    return w


def _workload(name, world=1):
    if name == 'multimat_100m':
        return multimat(name)
    if name == 'multimat_12m':          # same scene, one eighth of the z extent (development / tests)
        return multimat(name, z_cells=29)
    if name == 'multimat_sample':       # bounded CPU sample: a 4-cell z-slice between two slip planes, 1 chunk per material
        return multimat(name, z_cells=4, chunks_per_material=1, side_walls=True)
    if name == 'multimat_tiny':
        return multimat(name, z_cells=4, chunks_per_material=1, y_cells=24, side_walls=True)
    if name.startswith('brick_125m'):
        return brick(name, world)
    if name.startswith('brick_2m'):
        return brick(name, world, cells=(64, 64, 64))
    if name.startswith('cube_drop'):
        return cube_drop(name)
    if name == 'bunnies_30m':
        return bunnies(name)
    if name == 'bunnies_2m':
        return bunnies(name, 32)
    raise ValueError(name)


def substep_dt(w, default_dt):
    return w.get('dt', default_dt)


# Relative cost of a particle-substep by material (WATER, ELASTIC, SNOW, SAND) in the timed window of this scene, from
# the per-rank phase times of a 4-GPU run with one material per rank (roofline.kernel_ms_per_rank: busy time per
# particle 80 / 93 / 117 / 113 ps; profiles/README.md): the strong-scaling cuts give every rank the same COST, not
# the same count (with equal counts the ranks that hold snow and sand set the pace).  The cuts are static.
MATERIAL_COST = (1.0, 1.16, 1.46, 1.40)


def rank_chunks(w, rank, world):
    """Chunks whose particles rank `rank` needs to generate, and the cut planes (absolute leaf-block x) of `world`
    slabs.  Bricks (weak scaling): one brick per rank.  Otherwise (strong scaling): cost-balanced cuts on leaf-block
    boundaries; a rank generates every chunk that overlaps its slab and the solver keeps the rows inside it."""
    ch = w['chunks']
    if world == 1:
        return ch, []
    if hasattr(ch[0], 'rank'):
        mine = [c for c in ch if c.rank == rank]
        return mine, [(c + 2048) // 4 for c in w['cut_cells']]
    cost = np.array([MATERIAL_COST[c.material] * c.n for c in ch], np.float64)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    edges = [c.lo[0] - 0.5 for c in ch] + [ch[-1].hi[0] - 0.5]        # chunk boundaries in base-cell units
    cuts_cells = []
    for k in range(1, world):
        target = cum[-1] * k / world
        i = int(np.searchsorted(cum, target, side='right')) - 1
        i = min(max(i, 0), len(ch) - 1)
        frac = (target - cum[i]) / cost[i]
        cell = edges[i] + frac * (edges[i + 1] - edges[i])
        cell = int(round(cell / 4.0)) * 4                               # leaf-block boundary
        if cuts_cells and cell <= cuts_cells[-1]:
            cell = cuts_cells[-1] + 4
        cuts_cells.append(cell)
    lo = cuts_cells[rank - 1] if rank > 0 else -10**9
    hi = cuts_cells[rank] if rank < world - 1 else 10**9
    mine = [c for c in ch if c.hi[0] - 0.5 > lo and c.lo[0] - 0.5 < hi]
    return mine, [(c + 2048) // 4 for c in cuts_cells]


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['nvml_unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons)}


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def measured_traffic(workload_name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `workload_name`, from the committed
    ncu --set full capture (profiles/traffic.json, written by tools/ncu_summary.py); None if not captured."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            t = json.load(f)
        e = t.get(workload_name, {}).get(kernel)
        return None if e is None else {'bytes_per_launch': float(e['dram_bytes']), 'source': e.get('source')}
    except Exception:
        return None


# ------------------------------------------------------------------ CPU arm
def taichi_probe():
    """if Taichi is ever importable the real MPMSolver on ti.cpu is the reference arm."""
    try:
        import taichi  # noqa: F401
        return 'importable'
    except Exception as e:
        return f'absent ({type(e).__name__})'


def run_cpu_baseline(sample_name, steps, warmup, preroll=None):
    """Times the C/OpenMP restatement (oracle/mpm_oracle.c) on ALL host cores on a bounded sample of the workload.
    The thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    from oracle.c_oracle import COracle, build, load
    build()
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    load().oracle_set_threads(int(cores))
    w = workload(sample_name)
    o = COracle(w['res'], unbounded=w['unbounded'])
    o.set_gravity(list(w['gravity']))
    for point, normal, surface, friction in w['colliders']:
        o.add_surface_collider(point, normal, surface, friction)
    for c in w['chunks']:
        o.add_particles(c.positions(o.dx), c.material, velocity=c.velocity)
    dt = substep_dt(w, o.default_dt)
    pre = w['preroll'] if preroll is None else preroll
    t0 = time.perf_counter()
    for _ in range(pre + warmup):
        o.substep(dt)
    t_pre = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(steps):
        o.substep(dt)
    el = time.perf_counter() - t0
    return dict(value=w['n'] * steps / el, unit='particle-substeps/s', cores=COracle.num_threads(), kind='port',
                sample=f"{sample_name}: {w['n']} particles of the workload's scene (same materials, density, velocities, "
                       f"floor; a thin z-slice between two slip planes for the multi-material scene), {steps} timed "
                       f"substeps after {pre} pre-roll + {warmup} warm-up substeps ({t_pre:.1f} s), "
                       f"oracle/mpm_oracle.c with OpenMP on {COracle.num_threads()} threads; CPU restatement of the "
                       f"reference, not Taichi ({taichi_probe()})"), el / steps * 1e3, w, o


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    name = args.workload
    sample = {'multimat_100m': 'multimat_sample', 'multimat_12m': 'multimat_sample', 'cube_drop_4m': 'cube_drop_4m',
              'multimat_tiny': 'multimat_tiny'}.get(name, name)
    base, ms, w, _ = run_cpu_baseline(sample, max(args.steps, 1), max(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': 'particle-substeps/sec', 'value': base['value'],
        'unit': 'particle-substeps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong' if args.gpus > 1 else 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload(name)['label']},
        'cpu_baseline': base, 'taichi': taichi_probe(),
        'e2e': {'value': base['value'], 'unit': 'particle-substeps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
quiet = contextlib.redirect_stdout(io.StringIO())


def make_solver(w, world, rank, local, args, cuts):
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    kw = dict(res=w['res'], size=1, unbounded=w['unbounded'], device=local, use_g2p2g=args.g2p2g, quant=args.quant)
    with quiet:
        if world == 1:
            s = MPMSolver(**kw)
        else:
            from taichi_elements_b200.distributed import DistributedMPMSolver
            per_rank = w['n'] // world
            # a face of the slab in leaf blocks bounds the shared column; leavers per substep are a small
            # fraction of a face layer of particles
            face_blocks = 1 << 15
            s = DistributedMPMSolver(cuts=cuts, mig_capacity=max(1 << 15, per_rank // 256),
                                     halo_capacity=face_blocks, substep_batch=args.batch,
                                     comm=os.environ.get('MPM_COMM', 'auto'), **kw)
    s.set_gravity(w['gravity'])
    for point, normal, surface, friction in w['colliders']:
        s.add_surface_collider(point, normal, surface, friction)
    return s


def seed(s, chunks, world, host_cache=None):
    """add_particles for every chunk of this rank (host arrays in); returns the host arrays."""
    out = []
    for i, c in enumerate(chunks):
        x = host_cache[i] if host_cache is not None else c.positions(s.dx)
        if world == 1:
            s.add_particles(x, c.material, velocity=c.velocity)
        else:
            s.add_local_particles(x, c.material, velocity=c.velocity, id_base=c.id_base)
        out.append(x)
    return out


def timed_run(s, dt, args, dev, world, local, min_repeats=1):
    """W warm-up substeps, then `repeats` x [exactly K substeps between two CUDA events on the solver's stream,
    bracketed by barrier + synchronize]; the reported time is the MEDIAN repeat, each repeat the max over ranks."""
    import torch
    import torch.distributed as dist
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    s._lib.mpm_set_profiling(s._ctx, 0)
    s._run_substeps(dt, args.warmup)
    barrier()
    times, launches, st = [], 0, None
    with ClockSampler(local) as clk:
        t_begin = time.perf_counter()
        while True:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(stream)
            st = s._run_substeps(dt, args.steps)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            t = torch.tensor([ms, time.perf_counter() - t_begin], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t[0].item()))
            launches = int(st.launches)
            if args.repeats > 0:
                if len(times) >= args.repeats:
                    break
            elif (float(t[1].item()) >= args.min_seconds and len(times) >= 3) or len(times) >= 12:
                break
    return times, launches, st, clk.summary()


def phase_times(s, dt, steps):
    """Per-phase CUDA-event times (five records per substep on the kernels' stream; they serialise the PDL
    chain, so this is a separate pass after the timed one)."""
    s._lib.mpm_set_profiling(s._ctx, 1)
    st = s._run_substeps(dt, steps)
    s._lib.mpm_set_profiling(s._ctx, 0)
    return {'sort+structure': st.ms_sort, 'p2g': st.ms_p2g, 'grid_op': st.ms_grid, 'g2p': st.ms_g2p}


def parity_check(local, args):
    """Untimed: ONE substep of a sub-box of the benchmarked scene (multimat_tiny: same materials, velocities and
    floor) on the GPU against oracle/mpm_oracle.c, after a short pre-roll on both; per-particle relative errors."""
    from oracle.c_oracle import COracle, build
    from taichi_elements_b200.engine.mpm_solver import MPMSolver
    build()
    w = workload('multimat_tiny')
    with quiet:
        s = MPMSolver(res=w['res'], unbounded=True, device=local, use_g2p2g=False)
    o = COracle(w['res'], unbounded=True)
    for obj in (s, o):
        obj.set_gravity(list(w['gravity']))
        for point, normal, surface, friction in w['colliders']:
            obj.add_surface_collider(point, normal, surface, friction)
        for c in w['chunks']:
            obj.add_particles(c.positions(1.0 / 256), c.material, velocity=c.velocity)
    dt = o.default_dt
    pre = 12
    for _ in range(pre):
        o.substep(dt)
    s._run_substeps(dt, pre)
    # one substep from the SAME state: inject the oracle's state into the solver
    s._inject_state(o.x, o.v, o.F, o.C, o.Jp, o.material, o.color)
    o.substep(dt)
    s._run_substeps(dt, 1)

    def rel(a, b, floor):
        a = np.asarray(a, np.float64).reshape(len(a), -1)
        b = np.asarray(b, np.float64).reshape(len(b), -1)
        return float((np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), floor)).max())

    vs = float(np.abs(o.v).max())
    return {'scene': 'multimat_tiny (sub-box of the workload)', 'particles': int(o.n_particles), 'pre_roll_substeps': pre,
            'one_substep_rel_err': {'x': rel(s.x.to_numpy(), o.x, 1e-3), 'v': rel(s.v.to_numpy(), o.v, 1e-3 * vs),
                                    'F': rel(s.F.to_numpy(), o.F, 1e-3)},
            'tolerance': 1e-4, 'oracle': 'oracle/mpm_oracle.c (parity unpinned: no Taichi)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='multimat_100m')
    ap.add_argument('--repeats', type=int, default=0, help='timed repeats of K steps (0: until --min-seconds, >= 3)')
    ap.add_argument('--min-seconds', type=float, default=1.0, help='the K-step timing is repeated until this much wall time is covered')
    ap.add_argument('--batch', type=int, default=20, help='N > 1: substeps enqueued per host synchronisation')
    ap.add_argument('--preroll', type=int, default=-1, help='untimed substeps before the warm-up (-1: the workload\'s)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-weak', action='store_true', help='N > 1: skip the configs[4] weak-scaling brick')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--g2p2g', action='store_true', help='time the use_g2p2g=True mode')
    ap.add_argument('--quant', action='store_true', help='time the quant=True storage')
    args = ap.parse_args()
    if args.impl == 'reference':
        return main_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- the headline workload: device-resident throughput (`value`) ----------------
    w = workload(args.workload, world)
    chunks, cuts = rank_chunks(w, rank, world)
    mpm = make_solver(w, world, rank, local, args, cuts)
    if world > 1:
        mpm.reserve_blocks(max(1 << 16, int(sum(c.n for c in chunks) // 160)))
    host_parts = seed(mpm, chunks, world)
    for centre, radius, mat, vel in w.get('ellipsoids', []):      # configs[2]: seeded on the device
        with quiet:
            mpm.add_ellipsoid(center=centre, radius=radius, material=mat, velocity=vel)
    n_local = mpm.n_particles[None]
    n_total = int(allsum(n_local))
    if w['n'] is None:
        w['n'] = n_total
        args.no_e2e = True                                          # (no host chunks to upload for this scene)
    assert n_total == w['n'], (n_total, w['n'])
    dt = substep_dt(w, mpm.default_dt)
    preroll = w['preroll'] if args.preroll < 0 else args.preroll
    mpm._run_substeps(dt, preroll)
    times, launches, st, clocks = timed_run(mpm, dt, args, dev, world, local)
    ms_total = float(np.median(times))
    ms_per_step = ms_total / args.steps
    value = n_total * args.steps / (ms_total * 1e-3)
    n_after = int(allsum(mpm.n_particles[None]))

    # ---------------- roofline of the dominant kernel ----------------
    phases = phase_times(mpm, dt, args.steps) if world == 1 else None
    cells_active = int(st.n_grid_blocks) * 64
    n_now = mpm.n_particles[None]
    bpp = B_G2P2G_PARTICLE if args.g2p2g else B_PARTICLE
    bp2g = B_P2G_PARTICLE
    if args.quant:       # bit-packed storage moves fewer bytes: count what THAT layout needs, not the f32 figure
        # split: P2G reads xq vq Fq C Jp tag (80) + writes Fq Jp (24); G2P reads xq tag (12) + writes xq vq C tag (56)
        # fused: reads and writes xq vq Fq Jp tag (44 + 44)
        bpp, bp2g = (88, 88) if args.g2p2g else (172, 104)
    b_alg_p2g = bp2g * n_now + B_P2G_CELL * cells_active
    b_alg_step = bpp * n_now + B_CELL * cells_active
    if world == 1:
        dom = max(phases, key=phases.get)
        kname = 'k_p2g3<640,4>'
        achieved = b_alg_p2g / (phases['p2g'] * 1e-3) / 1e9 if phases['p2g'] > 0 else 0.0
        traffic = measured_traffic(args.workload, 'k_p2g3')
        roofline = {
            'bound': 'hbm', 'kernel': kname, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': traffic['bytes_per_launch'] if traffic else None,
            'traffic_source': traffic['source'] if traffic else 'no ncu capture committed for this workload',
            'traffic_note': 'includes 16 B per particle that P2G leaves at the sorted slot (x, tag) so that G2P streams them; '
                            'G2P moves 8 B per particle less than before for it',
            'peak_source': peak_src, 'algorithmic_bytes_per_launch': b_alg_p2g,
            'kernel_ms': {k: round(float(v), 4) for k, v in phases.items()}, 'dominant_phase': dom,
            'kernel_ms_note': 'CUDA events on the kernels\' stream, separate pass of K steps right after the timed ones '
                              '(five event records per substep serialise the PDL chain, so they are kept out of `value`)',
            'substep': {'algorithmic_bytes': b_alg_step, 'achieved': b_alg_step / (ms_per_step * 1e-3) / 1e9,
                        'frac': b_alg_step / (ms_per_step * 1e-3) / 1e9 / peak,
                        'frac_of_nominal_8TBs': b_alg_step / (ms_per_step * 1e-3) / 1e9 / 8000.0},
        }
    else:
        # per-rank phase times (separate pass with CUDA events on each rank's stream): "sort+structure" includes the
        # wait for the neighbours' migration message, "grid_op" the wait for their halo -- the skew between ranks
        ph = phase_times(mpm, dt, args.steps)
        tp = torch.tensor([ph['sort+structure'], ph['p2g'], ph['grid_op'], ph['g2p'], float(mpm.n_particles[None])],
                          dtype=torch.float64, device=dev)
        allp = [torch.empty_like(tp) for _ in range(world)]
        dist.all_gather(allp, tp)
        per_rank = [{'sort+structure': round(float(t[0]), 4), 'p2g': round(float(t[1]), 4), 'grid_op': round(float(t[2]), 4),
                     'g2p': round(float(t[3]), 4), 'particles': int(t[4])} for t in allp]
        # whole job: sum over ranks of the algorithmic bytes / max-over-ranks time, against N x peak
        b_job = allsum(b_alg_step)
        achieved = b_job / (ms_per_step * 1e-3) / 1e9
        roofline = {
            'bound': 'hbm', 'kernel': 'whole substep, all ranks (per-kernel events are a single-GPU pass)',
            'achieved': achieved, 'peak': peak * world, 'unit': 'GB/s', 'frac': achieved / (peak * world),
            'traffic': None, 'peak_source': peak_src + f' x {world} GPUs', 'algorithmic_bytes_per_launch': b_job,
            'kernel_ms_per_rank': per_rank,
        }

    # ---------------- end to end through the public API, host buffers ----------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(mpm, w, chunks, host_parts, world, dev, args, n_total)
    del host_parts

    # ---------------- N > 1: the configs[4] weak-scaling brick at the same N ----------------
    weak = None
    if not args.no_weak and args.workload == 'multimat_100m':
        del mpm
        torch.cuda.empty_cache()
        weak = run_weak(world, rank, local, dev, args, peak)

    cpu = None
    parity = None
    if rank == 0 and world == 1:
        if not args.no_parity:
            parity = parity_check(local, args)
        if not args.no_cpu_baseline:
            sample = {'multimat_100m': 'multimat_sample', 'multimat_12m': 'multimat_sample'}.get(args.workload, args.workload)
            if sample == 'cube_drop_4m':
                sample = 'cube_drop_sample'
            cpu, _, _, _ = run_cpu_baseline(sample, 8, 2)

    if rank == 0:
        mode = 'use_g2p2g=True (fused g2p2g kernel, SURVEY 8(f)1)' if args.g2p2g else 'default (split p2g / g2p)'
        if args.quant:
            mode += ', quant=True (bit-packed particle storage, SURVEY 8(f)2)'
        line = {
            'metric': 'particle-substeps/sec', 'value': value, 'unit': 'particle-substeps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong' if world > 1 and not hasattr(w['chunks'][0], 'rank') else 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': w['label'], 'mode': mode, 'particles': n_total, 'particles_per_gpu': n_local,
                       'pre_roll_substeps': preroll, 'particles_after': n_after,
                       'l2': 'inputs (2 x 116 B x N particle state) exceed L2; no flush needed',
                       'active_blocks': int(st.n_grid_blocks), 'particle_blocks': int(st.n_particle_blocks),
                       'max_velocity': float(st.max_velocity),
                       'timing': f'median of {len(times)} repeats of exactly {args.steps} substeps (CUDA events, max over '
                                 f'ranks per repeat)', 'ms_per_step_repeats': [round(t / args.steps, 5) for t in times]},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
            'gpu_launches_note': 'own kernels enqueued for the K timed substeps of one repeat (9 per substep inside a batch, '
                                 '+1 for its first, +2 per batch; one more per substep with slabs)',
            'roofline': roofline, 'cpu_baseline': cpu, 'parity_check': parity, 'weak_cfg5': weak,
            'taichi': taichi_probe(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(mpm, w, chunks, host_parts, world, dev, args, n_total):
    """The same metric through the public API with HOST buffers: every frame = clear_particles, add_particles of this
    rank's chunks from pinned host arrays (H2D), step(frame_dt) (the reference's host loop), particle_info() (D2H
    into fresh arrays).  Wall clock, max over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    pinned = [torch.from_numpy(x).pin_memory().numpy() for x in host_parts]
    frames = 2
    sub_per_frame, d2h = 0, 0
    with quiet:
        for it in range(1 + frames):          # first frame: warm-up of the same path (pinned-buffer cache, growth)
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            mpm.clear_particles()
            seed(mpm, chunks, world, host_cache=pinned)
            before = mpm.total_substeps
            mpm.step(w['frame_dt'])
            sub_per_frame = mpm.total_substeps - before
            if world > 1:
                mpm.flush_migration()
            info = mpm.particle_info()
            d2h = sum(a.nbytes for a in info.values())
            del info
    barrier()
    el = time.perf_counter() - t0
    te = torch.tensor([el, float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    el, d2h = float(te[0].item()), float(te[1].item())
    h2d = sum(x.nbytes for x in pinned)
    return {'value': n_total * sub_per_frame * frames / el, 'unit': 'particle-substeps/s',
            'h2d_bytes_per_step': h2d / sub_per_frame, 'd2h_bytes_per_step': d2h / sub_per_frame,
            'what': f'{frames} x [clear_particles, add_particles(host chunks), step({w["frame_dt"]}) = {sub_per_frame} '
                    f'substeps from the seeded state, particle_info()] through '
                    f'{"MPMSolver" if world == 1 else "DistributedMPMSolver (each rank its slab)"}; wall clock incl. copies, '
                    f'bytes per rank'}


def run_weak(world, rank, local, dev, args, peak):
    import torch
    import torch.distributed as dist
    w = workload('brick_125m', world)
    chunks, cuts = rank_chunks(w, rank, world)
    s = make_solver(w, world, rank, local, args, cuts)
    if world > 1:
        s.reserve_blocks(1 << 19)
    seed(s, chunks, world)
    n_local = s.n_particles[None]
    t = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    n_total = int(t.item())
    dt = substep_dt(w, s.default_dt)
    s._run_substeps(dt, w['preroll'])
    times, launches, st, clocks = timed_run(s, dt, args, dev, world, local)
    ms = float(np.median(times)) / args.steps
    value = n_total / (ms * 1e-3)
    b = B_PARTICLE * n_total + B_CELL * int(st.n_grid_blocks) * 64 * world
    return {'workload': w['label'], 'scaling': 'weak', 'value': value, 'unit': 'particle-substeps/s', 'ms_per_step': ms,
            'particles': n_total, 'particles_per_gpu': n_local, 'pre_roll_substeps': w['preroll'],
            'roofline_frac_substep': b / (ms * 1e-3) / 1e9 / (peak * world), 'clocks': clocks,
            'ms_per_step_repeats': [round(x / args.steps, 5) for x in times]}


if __name__ == '__main__':
    main()
