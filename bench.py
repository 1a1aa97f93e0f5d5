#!/usr/bin/env python
"""Benchmark of the MLS-MPM substep (BASELINE.json metric: particle-substeps/s).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference

A "step" is ONE substep (binning -> P2G -> grid op -> G2P) over the whole
particle set.  Workload at N=1 = BASELINE.json configs[1]: 3D cube drop,
res 256^3, 2 x 2^21 particles (ELASTIC over WATER), g=(0,-20,0), dt = 3e-3/39.
Prints one JSON line (see the contract in the task statement).
"""
import argparse
import json
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


def workload(name, rank=0, world=1, seed=2):
    """Synthetic particle blocks of the named shapes; returns dict."""
    rng = np.random.default_rng(seed + 1000 * rank)
    if name == 'cube_drop_4m':          # configs[1]
        side, res = 0.25, 256
    elif name == 'cube_drop_sample':     # bounded CPU sample of configs[1]: same density, smaller cubes
        side, res = 0.16, 256
    elif name == 'cube_drop_small':      # smoke-sized
        side, res = 0.0625, 256
    elif name == 'cube_drop_32m':        # 2 x 128^3 cells x 8 at res 512
        side, res = 0.25, 512
    elif name == 'cube_drop_100m':       # 2 x 184^3 cells x 8 = 99.7 M at res 512
        side, res = 0.359375, 512
    else:
        raise ValueError(name)
    n_side = int(round(side * res))
    n_each = n_side**3 * 8               # 8 particles per cell
    lo_e = np.array([0.5 - side / 2, 0.55, 0.5 - side / 2], np.float32)
    lo_w = np.array([0.5 - side / 2, 0.15, 0.5 - side / 2], np.float32)
    xe = rng.random((n_each, 3), dtype=np.float32)
    xe *= np.float32(side)
    xe += lo_e
    xw = rng.random((n_each, 3), dtype=np.float32)
    xw *= np.float32(side)
    xw += lo_w
    return dict(res=(res, ) * 3, gravity=(0, -20, 0), frame_dt=3e-3, parts=[(xe, 1), (xw, 0)],
                n=2 * n_each, name=name)


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
            self._stop.wait(0.05)

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


def run_cpu_baseline(sample_name, steps, warmup):
    """Times the C/OpenMP restatement (oracle/mpm_oracle.c) on the host cores."""
    from oracle.c_oracle import COracle, build
    build()
    w = workload(sample_name)
    o = COracle(w['res'])
    o.set_gravity(w['gravity'])
    for x, m in w['parts']:
        o.add_particles(x, m)
    dt = w['frame_dt'] / (int(w['frame_dt'] / o.default_dt) + 1)
    for _ in range(warmup):
        o.substep(dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.substep(dt)
    el = time.perf_counter() - t0
    return dict(value=w['n'] * steps / el, unit='particle-substeps/s', cores=COracle.num_threads(), kind='port',
                sample=f"{sample_name}: {w['n']} particles (same scene and 8 particles/cell as the workload, "
                       f"cube side 0.16), {steps} substeps after {warmup} warm-up, oracle/mpm_oracle.c with "
                       f"OpenMP; CPU restatement of the reference, not Taichi"), el / steps * 1e3, w


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base, ms, w = run_cpu_baseline('cube_drop_sample', max(args.steps, 1), max(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': 'particle-substeps/sec', 'value': base['value'],
        'unit': 'particle-substeps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'configs[1] 3D cube drop res 256^3 WATER+ELASTIC, bounded sample '
                               f"({w['n']} particles)"},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': 'particle-substeps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cube_drop_4m')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--g2p2g', action='store_true', help='time the use_g2p2g=True mode (N = 1)')
    ap.add_argument('--strong', action='store_true', help='N > 1: split the N=1 scene over the ranks (strong scaling)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return main_reference(args)

    import torch
    import torch.distributed as dist
    from taichi_elements_b200.engine.mpm_solver import MPMSolver

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    w = workload(args.workload, rank=rank, world=world)
    import contextlib
    import io
    quiet = contextlib.redirect_stdout(io.StringIO())

    def make_solver():
        with quiet:
            if world == 1:
                s = MPMSolver(res=w['res'], size=1, unbounded=False, device=local, use_g2p2g=args.g2p2g)
            else:
                from taichi_elements_b200.distributed import DistributedMPMSolver, SlabDecomposition
                res = w['res'][0]
                if args.strong:
                    allx = np.concatenate([x[:, 0] for x, _ in w['parts']])
                    cuts = SlabDecomposition.balanced_cuts(allx, world, 4, 4096, float(res))
                else:
                    cuts = [int((math.floor((x0 + side * k) * res) + 2048) // 4) for k in range(1, world)]
                s = DistributedMPMSolver(res=w['res'], cuts=cuts, size=1, unbounded=False, device=local,
                                         mig_capacity=1 << 14 if not args.strong else 1 << 16,
                                         halo_capacity=1 << 11 if not args.strong else 1 << 13, substep_batch=20,
                                         comm=os.environ.get('MPM_COMM', 'auto'))
                s.reserve_blocks(1 << 16)
        s.set_gravity(w['gravity'])
        return s

    # N > 1: weak scaling -- one brick of the N=1 scene per rank, bricks contiguous along x so that
    # every cut carries a shared grid column and migrating particles; every rank builds the same
    # particle list and keeps its slab
    import math
    side = 0.25
    x0 = 0.5 - side * world / 2
    if world == 1 or args.strong:
        bricks = [w['parts']]
    else:
        bricks = []
        for r in range(world):
            wr = workload(args.workload, rank=r, world=world)
            shift = np.array([x0 + side * r - (0.5 - side / 2), 0, 0], np.float32)
            bricks.append([(x + shift, m) for x, m in wr['parts']])
    mpm = make_solver()
    for parts in bricks:
        for x, m in parts:
            mpm.add_particles(x, m)
    n_local = mpm.n_particles[None]
    if world > 1:
        mpm.reserve_blocks(max(1 << 16, n_local // 128))   # block capacity cannot grow inside a distributed batch
    dt = w['frame_dt'] / (int(w['frame_dt'] / mpm.default_dt) + 1)
    if w['res'][0] != 256:
        dt = mpm.default_dt
    lib, ctx = mpm._lib, mpm._ctx
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput (`value`) ----------------
    lib.mpm_set_profiling(ctx, 0)
    mpm._run_substeps(dt, args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        st = mpm._run_substeps(dt, args.steps)
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = int(st.launches)
    # per-phase CUDA events (five records per substep on the kernels' stream) perturb the stream, so the
    # phase times come from a second, untimed-for-`value` pass of the same length right after
    if world == 1:
        lib.mpm_set_profiling(ctx, 1)
        st = mpm._run_substeps(dt, args.steps)
        lib.mpm_set_profiling(ctx, 0)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    n_total = int(cnt.item())
    value = n_total * args.steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps
    lib.mpm_set_profiling(ctx, 0)

    # ---------------- roofline of the dominant kernel (P2G) ----------------
    peak, peak_src = peaks()
    cells_active = int(st.n_grid_blocks) * 64
    b_alg_p2g = B_P2G_PARTICLE * n_local + B_P2G_CELL * cells_active
    b_alg_step = B_PARTICLE * n_local + B_CELL * cells_active
    phases = {'sort+structure': st.ms_sort, 'p2g': st.ms_p2g, 'grid_op': st.ms_grid, 'g2p': st.ms_g2p}
    dom = max(phases, key=phases.get)
    achieved = b_alg_p2g / (st.ms_p2g * 1e-3) / 1e9 if st.ms_p2g > 0 else 0.0
    if world > 1:   # phase API: no per-kernel events; report the whole substep per GPU
        achieved = b_alg_step / (ms_per_step * 1e-3) / 1e9
    roofline = {
        'bound': 'hbm', 'kernel': 'k_p2g3<640,4>' if world == 1 else 'whole substep, per GPU (no per-kernel events in the multi-GPU phase path)',
        'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
        'frac': achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of k_p2g3<640,4> on this workload, one ncu --set full
        # capture (profiles/r01_ncu_final_v6.md): 471.3 MB + 157.9 MB per launch
        'traffic': 629.2e6 if (args.workload == 'cube_drop_4m' and world == 1) else None,
        'peak_source': peak_src,
        'algorithmic_bytes_per_launch': b_alg_p2g,
        'kernel_ms': {k: round(float(v), 4) for k, v in phases.items()}, 'dominant_phase': dom,
        'kernel_ms_note': 'CUDA events on the kernels\' stream, second pass of the same K steps right after the timed one '
                          '(five event records per substep serialise the PDL chain, so they are kept out of `value`)',
        'substep': {'algorithmic_bytes': b_alg_step,
                    'achieved': b_alg_step / (ms_per_step * 1e-3) / 1e9 if world == 1 else None,
                    'frac': b_alg_step / (ms_per_step * 1e-3) / 1e9 / peak if world == 1 else None},
    }

    # ---------------- end to end through the public API, host buffers ----------------
    e2e = None
    if not args.no_e2e:
        frames = max(1, min(3, args.steps // 40))
        # N > 1: the inputs of a rank are the rows of its slab (own brick plus the few boundary rows of the
        # neighbouring bricks whose base block falls on this side of the cut), selected once, outside the timed region
        if world == 1:
            host_parts = [(torch.from_numpy(x).pin_memory().numpy(), m) for parts in bricks for x, m in parts]
        else:
            host_parts = []
            for parts in bricks:
                for x, m in parts:
                    rows = x[mpm.slab.mine(x[:, 0])]
                    if len(rows):
                        host_parts.append((torch.from_numpy(np.ascontiguousarray(rows)).pin_memory().numpy(), m))
        sub_per_frame = 0
        with quiet:
            mpm2 = mpm
            mpm2.clear_particles()
            keep = []
            for _ in range(2):             # warm-up of the same path (also warms the pinned-buffer cache)
                mpm2.clear_particles()
                for x, m in host_parts:
                    mpm2.add_particles(x, m)
                mpm2.step(w['frame_dt'])
                if world > 1:
                    mpm2.flush_migration()
                keep.append(mpm2.particle_info())
            del keep
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        with quiet:
            for _ in range(frames):
                mpm2.clear_particles()
                for x, m in host_parts:
                    mpm2.add_particles(x, m)
                before = mpm2.total_substeps
                mpm2.step(w['frame_dt'])
                sub_per_frame = mpm2.total_substeps - before
                if world > 1:
                    mpm2.flush_migration()
                info = mpm2.particle_info()      # N > 1: this rank's particles (position, velocity, material, color, id)
                d2h = sum(a.nbytes for a in info.values())
        barrier()
        el = time.perf_counter() - t0
        te = torch.tensor([el, float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el, d2h = float(te[0].item()), float(te[1].item())
        h2d = sum(x.nbytes for x, _ in host_parts)                  # this rank's rows
        e2e = {'value': n_total * sub_per_frame * frames / el, 'unit': 'particle-substeps/s',
               'h2d_bytes_per_step': h2d / sub_per_frame, 'd2h_bytes_per_step': d2h / sub_per_frame,
               'what': f'{frames} x [clear_particles, add_particles(host arrays), step({w["frame_dt"]}) = '
                       f'{sub_per_frame} substeps, particle_info()] through MPMSolver (N > 1: DistributedMPMSolver, each rank its slab); wall clock incl. copies'}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = run_cpu_baseline('cube_drop_sample', 8, 2)

    if rank == 0:
        line = {
            'metric': 'particle-substeps/sec', 'value': value, 'unit': 'particle-substeps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f"configs[1] 3D cube drop, res 256^3 bounded, {n_total} particles "
                                   f"(ELASTIC cube over WATER cube, 8 per cell), g=(0,-20,0), dt=3e-3/39"
                                   if args.workload == 'cube_drop_4m' else args.workload,
                       'mode': 'use_g2p2g=True (fused order, SURVEY 8(f)1)' if args.g2p2g else 'default (split p2g / g2p)',
                       'particles_per_gpu': n_local, 'l2': 'inputs (2 x 116 B x N particle state) exceed L2',
                       'active_blocks': int(st.n_grid_blocks), 'particle_blocks': int(st.n_particle_blocks)},
            'clocks': clk.summary(), 'e2e': e2e, 'gpu_launches': launches,
            'gpu_launches_note': 'own kernels: 9 per substep inside a batch (scan, rank, scan, scatter, finish, clear, p2g, grid op, g2p), 10 for its first substep, + 2 per batch; more with slabs',
            'roofline': roofline, 'cpu_baseline': cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
