#!/usr/bin/env python
"""Benchmark of the numpad hot path on B200: Newton iterations per second.

One "step" = one full Newton iteration of the lid-driven cavity (reference
test/cavity.py, BASELINE.json configs[1]/metric) at the fixed benchmark state
of SURVEY.md section 8(d) config 2(ii):

    residual evaluation (tape recording)          adsolve.py:176-179
  + Jacobian assembly  J = res.diff(u)            adsolve.py:192
  + linear solve       J du = res  to 1e-10       adsolve.py:193-195

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--impl reference]

`value`  : steps/s with the state resident in HBM.
`e2e`    : the same through host buffers (pinned H2D of u, u_old; D2H of du).
`roofline`: the CSR SpMV of the assembled Jacobian (the kernel the Krylov solve
            is built from), algorithmic bytes 12*nnz + 4*(n+1) + 16*n per launch,
            timed with CUDA events; peak from MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU oracle (numpy + scipy SuperLU, the
            reference's own algorithm) on a bounded sub-grid sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--grid', type=int, default=int(os.environ.get('NPB_BENCH_GRID', '2048')))
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--sample-grid', type=int, default=192)
    ap.add_argument('--rtol', type=float, default=1e-10)
    ap.add_argument('--restart', type=int, default=100)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parallel', default='sharded', choices=['sharded', 'replicas'],
                    help='N > 1: one problem row-sharded over the GPUs (strong scaling) or N independent replicas')
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# CPU arm: the oracle (a port of the reference's algorithm on scipy)
# --------------------------------------------------------------------------- #
def cpu_newton_iteration(sample_grid):
    """time one Newton iteration of the oracle at `sample_grid`; returns phases"""
    import scipy.sparse.linalg as spla
    from oracle import numpad_oracle as orc
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(orc, sample_grid)
    u0 = P.fixed_state()
    t0 = time.perf_counter()
    u = orc.array(u0.copy())
    res = P.residual(u, orc.array(u0), 1.0, 0)
    t1 = time.perf_counter()
    J = res.diff(u).tocsr()
    t2 = time.perf_counter()
    du = spla.spsolve(J, np.ravel(res._value), use_umfpack=False)
    t3 = time.perf_counter()
    x = np.random.RandomState(1).standard_normal(P.n)
    ts = time.perf_counter()
    for _ in range(5):
        J @ x
    spmv = (time.perf_counter() - ts) / 5
    relres = np.linalg.norm(J @ du - res._value) / np.linalg.norm(res._value)
    return dict(n=P.n, nnz=int(J.nnz), residual_s=t1 - t0, assembly_s=t2 - t1,
                spsolve_s=t3 - t2, total_s=t3 - t0, spmv_s=spmv, relres=float(relres))


def cpu_baseline(full_grid, sample_grid, steps=1, warmup=0):
    n_full = full_grid * (3 * full_grid - 2)
    runs = [cpu_newton_iteration(sample_grid) for _ in range(warmup + steps)][warmup:]
    t = float(np.mean([r['total_s'] for r in runs]))
    r = runs[-1]
    frac = r['n'] / n_full
    return dict(
        value=frac / t, unit='Newton iters/sec', cores=1, kind='port',
        sample=('one Newton iteration (residual %.2fs + scipy chain-rule assembly %.2fs + '
                'SuperLU spsolve %.2fs) of the same cavity residual on a %d^2 sub-grid '
                '(%d of %d unknowns = %.3g of the workload); value = that fraction / time, '
                'i.e. time scaled LINEARLY in unknowns -- an upper bound on the reference '
                "throughput at the full grid since SuperLU grows ~n^1.5; scipy's sparse "
                'kernels and SuperLU are single-threaded (1 core of %d)'
                % (r['residual_s'], r['assembly_s'], r['spsolve_s'], sample_grid, r['n'],
                   n_full, frac, os.cpu_count())),
        sample_iters_per_sec=1.0 / t, sample_grid=sample_grid, sample_seconds=t,
        spmv_gbs=(12.0 * r['nnz'] + 4.0 * (r['n'] + 1) + 16.0 * r['n']) / r['spmv_s'] / 1e9)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_baseline(args.grid, args.sample_grid, steps=args.steps, warmup=args.warmup)
    n = args.grid * (3 * args.grid - 2)
    line = {
        'impl': 'reference', 'metric': 'Newton iters/sec (Jacobian assembly+solve), cavity',
        'value': cb['value'], 'unit': 'Newton iters/sec', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * cb['sample_seconds'], 'higher_is_better': True,
        'scaling': 'strong' if (args.gpus > 1 and args.parallel == 'sharded') else 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.grid, n, args.gpus, args.parallel),
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': 'Newton iters/sec', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t0,
    }
    print(json.dumps(line))


def workload_config(grid, n, gpus, parallel='sharded'):
    return {'workload': 'lid-driven cavity %d^2 (test/cavity.py residual), n=%d unknowns, '
                        'Re=1e4, dt=1, fixed state 1e-2*RandomState(0).standard_normal(n); '
                        'one full Newton iteration per step' % (grid, n),
            'grid': grid, 'unknowns': n, 'linear_rtol': None,
            'l2_policy': 'inputs larger than L2 (Jacobian alone > 126 MB)',
            'parallelism': 'single GPU' if gpus <= 1 else (
                ('%d independent replicas of the single-GPU problem, one per GPU, no data-path '
                 'collective' % gpus) if parallel == 'replicas' else
                ('ONE problem, linear solve row-sharded over %d GPUs in grid strips (halo all-gather '
                 'before every SpMV, all-reduce after every inner product, NCCL over NVLink); residual, '
                 'Jacobian assembly and multigrid setup replicated on every rank (DESIGN.md section 7)'
                 % gpus))}


# --------------------------------------------------------------------------- #
# clocks sampling
# --------------------------------------------------------------------------- #
class Clocks:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.samples, self.index, self.stop = [], index, False
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index),
                                      '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                self.samples.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=3)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6
                          for i in range(4) if s[2 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.samples)}


# --------------------------------------------------------------------------- #
# GPU arm
# --------------------------------------------------------------------------- #
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    import numpad_b200 as npd
    from numpad_b200 import linsolve, _lib, precond, shard
    from numpad_b200.workloads.cavity import CavityProblem
    dev = torch.device('cuda', local)
    sharded = world > 1 and args.parallel == 'sharded'
    P = CavityProblem(npd, args.grid)
    owner = P.partition(world) if sharded else None
    n = P.n
    u0 = P.fixed_state()
    u_host = torch.from_numpy(u0.copy()).pin_memory()
    uold_host = torch.from_numpy(u0.copy()).pin_memory()
    du_host = torch.empty(n, dtype=torch.float64).pin_memory()
    u_dev = u_host.to(dev)
    uold_dev = uold_host.to(dev)
    phases = {}

    def newton_iteration(ud, uoldd, record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            ev[0].record()
        u = npd.adarray(ud)
        res = P.residual(u, npd.adarray(uoldd), 1.0, 0)
        if record:
            ev[1].record()
        J = npd.diff_dev(res, u)
        if record:
            ev[2].record()
        # every step pays the full preconditioner setup (no hierarchy is carried
        # over from the previous step, although a real Newton march could)
        precond._last['lsc'] = None
        if sharded:
            du = shard.solve(J, res._dev.reshape(-1), owner=owner, rtol=args.rtol, restart=args.restart)
        else:
            du = linsolve.solve(J, res._dev.reshape(-1), rtol=args.rtol, restart=args.restart)
        if record:
            ev[3].record()
            torch.cuda.synchronize()
            li = shard.last_info if sharded else linsolve.last_info
            phases.update(residual_ms=ev[0].elapsed_time(ev[1]), assembly_ms=ev[1].elapsed_time(ev[2]),
                          linear_solve_ms=ev[2].elapsed_time(ev[3]), precond_setup_ms=li.get('setup_ms'),
                          krylov_ms=li.get('krylov_ms'), krylov_iters=li.get('iters'),
                          method=li.get('method'), nnz=J.nnz, relres=li['rnorm'] / li['bnorm'],
                          precond=li.get('precond'))
            if sharded:
                phases.update(permute_ms=li.get('permute_ms'), localize_ms=li.get('localize_ms'),
                              halo_values_per_exchange=li.get('halo'))
        return J, du

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step_resident():
        newton_iteration(u_dev, uold_dev)

    def step_e2e():
        ud = u_host.to(dev, non_blocking=True)
        uo = uold_host.to(dev, non_blocking=True)
        _, du = newton_iteration(ud, uo)
        du_host.copy_(du, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    J, du = newton_iteration(u_dev, uold_dev, record=True)
    L = _lib.lib().cdll
    with Clocks(local) as clk:
        L.npb_launch_count_reset()
        L.npb_profile_matvec(1)            # events around the Jacobian SpMV inside the timed steps
        ms = timed(step_resident, args.steps)
        launches = int(L.npb_launch_count())
        L.npb_profile_matvec(0)
        import ctypes
        tot_ms, cnt = ctypes.c_double(0.0), ctypes.c_int64(0)
        L.npb_profile_matvec_read(ctypes.byref(tot_ms), ctypes.byref(cnt))
        ms_spmv_insitu = tot_ms.value / cnt.value if cnt.value else None
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        # ---- roofline of the dominant Krylov building block: CSR SpMV of J
        Jm = J.materialized()
        x = torch.randn(n, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        for _ in range(5):
            Jm.matvec(x, out=y)
        reps = 50
        ms_spmv = timed(lambda: Jm.matvec(x, out=y), reps) / reps
    clocks = clk.summary()
    spmv_bytes = 12.0 * Jm.nnz + 4.0 * (n + 1) + 16.0 * n
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    standalone = spmv_bytes / (ms_spmv * 1e-3) / 1e9
    # headline: the launch as it ran inside the timed Newton steps; the back-to-back loop
    # after the region (same kernel, same matrix) is reported beside it
    if ms_spmv_insitu:
        achieved, ms_roof, how = spmv_bytes / (ms_spmv_insitu * 1e-3) / 1e9, ms_spmv_insitu, \
            'CUDA events around each Jacobian SpMV launch of the Krylov loop inside the timed steps (%d launches)' % cnt.value
    else:
        achieved, ms_roof, how = standalone, ms_spmv, 'back-to-back loop of 50 launches after the timed steps'
    # DRAM traffic of this kernel from the committed `ncu --set full` capture
    # (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(str(args.grid))
    except Exception:
        pass
    # replicas: N problems per step; sharded: one problem per step (strong scaling)
    per_step = 1 if (sharded or world == 1) else world
    value = per_step * args.steps / (ms * 1e-3)
    e2e = per_step * args.steps / (ms_e2e * 1e-3)
    if rank == 0:
        cfg = workload_config(args.grid, n, world, args.parallel)
        cfg['linear_rtol'] = args.rtol
        line = {
            'metric': 'Newton iters/sec (Jacobian assembly+solve), cavity',
            'value': value, 'unit': 'Newton iters/sec', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if sharded else 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
            'e2e': {'value': e2e, 'unit': 'Newton iters/sec', 'h2d_bytes_per_step': 16 * n,
                    'd2h_bytes_per_step': 8 * n, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'kernel': 'k_spmv_tma (persistent TMA-pipelined CSR SpMV, here of the Jacobian; all its instances together: 56 % of the device time of a step, profiles/r1_launches_cavity2048_newton_step.txt)',
                         'achieved': achieved, 'peak': peak,
                         'peak_source': 'measured' if peaks else 'fallback',
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                         'bytes_per_launch': spmv_bytes, 'us_per_launch': 1e3 * ms_roof,
                         'measured': how, 'standalone_gbs': standalone,
                         'standalone_us_per_launch': 1e3 * ms_spmv,
                         'frac_of_nominal_8TBs': achieved / 8000.0},
            'phases': phases, 'clocks': clocks,
        }
        if not args.no_cpu_baseline and world == 1:       # rank 0 at N = 1 only
            line['cpu_baseline'] = cpu_baseline(args.grid, args.sample_grid)
        print(json.dumps(line))
    if world > 1:
        shard.shutdown()
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
