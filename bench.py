#!/usr/bin/env python
"""Benchmark of the numpad hot path on B200: Newton iterations per second.

One "step" = one Newton iteration of the lid-driven cavity (reference
test/cavity.py, BASELINE.json configs[1]/metric) at the fixed benchmark state
of SURVEY.md section 8(d) config 2(ii), made THROUGH THE DROP-IN API at the
product's default linear tolerance (linsolve.DEFAULTS):

    solve_newton_with_dt(residual, u, (u_old, dt=1, f=0), dt=inf, max_iter=2, rel_tol=1)
      = residual evaluation (tape recording)          adsolve.py:176-179
      + Jacobian assembly  J = res.diff(u)            adsolve.py:192
      + linear solve       J du = res  to 1e-12       adsolve.py:193-195
      + update, second residual evaluation + convergence test, adsolution    :198, :176-188

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--impl reference]

`value`  : steps/s with the state resident in HBM.
`e2e`    : the same through host buffers (pinned H2D of u, u_old; D2H of the new state).
`roofline`: the CSR SpMV of the assembled Jacobian (the kernel the Krylov solve
            is built from), algorithmic bytes 12*nnz + 4*(n+1) + 16*n per launch,
            timed in situ with CUDA events; peak from MEASURED_PEAKS.json.  Beside it
            `krylov_phase`: the algorithmic bytes of EVERY operator application and
            vector kernel of the Krylov phase over the phase time.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (baseline/_ref/numpad, pure
            Python on scipy) under the numpy-2 import shim, on a bounded sub-grid sample;
            no part of this repository's engine is imported in that arm.
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

METRIC = 'Newton iters/sec (Jacobian assembly+solve), cavity'
UNIT = 'Newton iters/sec'
LINEAR_RTOL = 1e-12          # = numpad_b200.linsolve.DEFAULTS['rtol'] (asserted in run_ours)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--grid', type=int, default=int(os.environ.get('NPB_BENCH_GRID', '2048')))
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--workload', default='cavity', choices=['cavity', 'euler', 'ns'],
                    help='cavity (the headline metric, default grid 2048) or one implicit step of '
                         'test/euler_kec.py / test/navierstokes.py on the synthetic duct mesh at the '
                         "scripts' time step (configs 4 / 5; pass a --grid the direct solver can hold, e.g. 256)")
    ap.add_argument('--sample-grid', type=int, default=0,
                    help='reference arm: grid of the per-step sample (0: from --ref-budget)')
    ap.add_argument('--ref-budget', type=float, default=200.0,
                    help='reference arm: seconds of CPU work for the K + W sampled steps')
    ap.add_argument('--trend-grid', type=int, default=400,
                    help='reference arm: one extra iteration at this grid for the n^1.5 trend (0: off)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parallel', default='sharded', choices=['sharded', 'replicas'],
                    help='N > 1: one problem row-sharded over the GPUs (strong scaling) or N independent replicas')
    return ap.parse_args()


def workload_config(grid, gpus, parallel='sharded', workload='cavity'):
    """identical in both arms (the driver compares them)"""
    if workload != 'cavity':
        n = 4 * grid * grid
        return {'workload': '%s on the synthetic %d x %d duct mesh of SURVEY 8(d) config %d, n=%d unknowns, '
                            'dt=%s (the script\'s), state = uniform subsonic flow * (1 + 1e-3 '
                            'RandomState(0).standard_normal); one Newton iteration (residual + Jacobian '
                            'assembly + linear solve) per step, made through solve_newton_with_dt'
                            % ('test/euler_kec.py' if workload == 'euler' else 'test/navierstokes.py',
                               grid, grid, 4 if workload == 'euler' else 5, n,
                               '0.1' if workload == 'euler' else '1/%d' % grid),
                'grid': grid, 'unknowns': n, 'linear_rtol': LINEAR_RTOL,
                'l2_policy': 'inputs larger than L2 above 256^2; smaller grids fit (stated)',
                'parallelism': 'single GPU' if gpus <= 1 else '%d independent replicas' % gpus}
    n = grid * (3 * grid - 2)
    return {'workload': 'lid-driven cavity %d^2 (test/cavity.py residual), n=%d unknowns, '
                        'Re=1e4, dt=1, fixed state 1e-2*RandomState(0).standard_normal(n); '
                        'one Newton iteration (residual + Jacobian assembly + linear solve) per step, '
                        'made through solve_newton_with_dt' % (grid, n),
            'grid': grid, 'unknowns': n, 'linear_rtol': LINEAR_RTOL,
            'l2_policy': 'inputs larger than L2 (Jacobian alone > 126 MB)',
            'parallelism': 'single GPU' if gpus <= 1 else (
                ('%d independent replicas of the single-GPU problem, one per GPU, no data-path '
                 'collective' % gpus) if parallel == 'replicas' else
                ('ONE problem, linear solve row-sharded over %d GPUs in grid strips (halo exchange '
                 'before every SpMV, all-reduce after every inner product, over NVLink); residual, '
                 'Jacobian assembly and multigrid setup replicated on every rank (DESIGN.md section 7)'
                 % gpus))}


# --------------------------------------------------------------------------- #
# CPU arm: the unmodified reference (baseline/_ref/numpad) under the import shim of
# SURVEY Appendix C; falls back to the oracle port when the copy is absent.
# Nothing of numpad_b200 is imported here.
# --------------------------------------------------------------------------- #
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def import_reference():
    """-> (module, kind).  The reference needs: a no-op np.set_numeric_ops (removed in
    NumPy 2), adarray.__array_ufunc__ = None, time.clock."""
    if os.path.exists(os.path.join(REF_DIR, 'numpad', 'adarray.py')):
        ops = {k: getattr(np, k) for k in ('add', 'subtract', 'multiply', 'true_divide', 'divide')}
        np.set_numeric_ops = lambda **kw: dict(ops)
        time.clock = time.perf_counter
        import warnings
        warnings.filterwarnings('ignore')
        sys.path.insert(0, REF_DIR)
        import numpad
        numpad.adarray.__array_ufunc__ = None
        return numpad, 'reference'
    from oracle import numpad_oracle
    return numpad_oracle, 'port'


def kec_reference_problem(numpad, kind, N, viscous):
    """the Euler / NS residual at grid N for the CPU arm: the restated workload
    (numpad_b200/workloads/kec.py, loaded by path: no engine import) on the reference's or the
    oracle's array namespace -- the scripts themselves hard-wire a 90 x 20 / 80 x 20 mesh file-free
    geometry and pylab, SURVEY 8c"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        '_kec_workload', os.path.join(ROOT, 'numpad_b200', 'workloads', 'kec.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.KecProblem(numpad, N, N, viscous=viscous)


def reference_residual(numpad, kind, N):
    """the cavity residual function bound to grid N: the reference's own test/cavity.py source
    (everything above its time-integration driver) / the restated workload for the port"""
    if kind == 'reference':
        src = open(os.path.join(REF_DIR, 'numpad', 'test', 'cavity.py')).read()
        src = src.split('# ---------------------- time integration')[0]
        g = {}
        exec(compile(src, 'baseline/_ref/numpad/test/cavity.py', 'exec'), g)
        g.update(N=N, dx=1. / N, dy=1. / N, Re=10000)
        return g['cavity']
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        '_cavity_workload', os.path.join(ROOT, 'numpad_b200', 'workloads', 'cavity.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CavityProblem(numpad, N).residual


def cpu_newton_iteration(numpad, kind, N, workload='cavity'):
    """one Newton iteration of the reference at grid N, phases timed separately"""
    import scipy.sparse.linalg as spla
    if workload == 'cavity':
        func = reference_residual(numpad, kind, N)
        n = N * (3 * N - 2)
        u0 = 1e-2 * np.random.RandomState(0).standard_normal(n)
        call = lambda u: func(u, numpad.array(u0), 1.0, 0)
    else:
        P = kec_reference_problem(numpad, kind, N, workload == 'ns')
        n, u0 = P.n, P.fixed_state()
        dt = 1.0 / N if workload == 'ns' else 0.1
        call = lambda u: P.residual(u, numpad.array(u0), dt)
    t0 = time.perf_counter()
    u = numpad.array(u0.copy())
    res = call(u)
    t1 = time.perf_counter()
    J = res.diff(u).tocsr()
    t2 = time.perf_counter()
    du = spla.spsolve(J, np.ravel(res._value), use_umfpack=False)     # adsolve.py:193-195
    t3 = time.perf_counter()
    x = np.random.RandomState(1).standard_normal(n)
    ts = time.perf_counter()
    for _ in range(3):
        J @ x
    spmv = (time.perf_counter() - ts) / 3
    relres = np.linalg.norm(J @ du - np.ravel(res._value)) / np.linalg.norm(res._value)
    return dict(grid=N, n=n, nnz=int(J.nnz), residual_s=t1 - t0, assembly_s=t2 - t1,
                spsolve_s=t3 - t2, total_s=t3 - t0, spmv_s=spmv, relres=float(relres))


def pick_sample_grid(budget_s, iters):
    """largest grid whose iteration fits budget / iters, from the measured trend of
    BASELINE.md section 2.1 (spsolve 35.1 s at 400^2, ~n^1.5; diff + residual ~ n)"""
    per = budget_s / max(iters, 1)
    g = 400
    t = lambda N: 35.1 * (N / 400.) ** 3 + 1.9 * (N / 400.) ** 2
    while g > 32 and t(g) * 1.25 > per:
        g -= 8
    return g


def cpu_baseline(full_grid, sample_grid, steps=1, warmup=0, trend_grid=0, budget=200.0,
                 workload='cavity'):
    numpad, kind = import_reference()
    if not sample_grid:
        sample_grid = pick_sample_grid(budget, steps + warmup)
        if workload != 'cavity':        # 4 unknowns per cell, ~24-36 entries per row: ~3x the work
            sample_grid = max(24, int(sample_grid * 0.55) // 8 * 8)
        sample_grid = min(sample_grid, full_grid)
    n_full = full_grid * (3 * full_grid - 2) if workload == 'cavity' else 4 * full_grid * full_grid
    runs = [cpu_newton_iteration(numpad, kind, sample_grid, workload) for _ in range(warmup + steps)][warmup:]
    t = float(np.mean([r['total_s'] for r in runs]))
    r = runs[-1]
    frac = r['n'] / n_full
    trend = [dict(grid=r['grid'], n=r['n'], total_s=t, spsolve_s=float(np.mean([q['spsolve_s'] for q in runs])))]
    n15 = None
    if trend_grid and trend_grid > sample_grid:
        q = cpu_newton_iteration(numpad, kind, trend_grid, workload)
        trend.append(dict(grid=q['grid'], n=q['n'], total_s=q['total_s'], spsolve_s=q['spsolve_s']))
        # measured exponent of the factorisation between the two sizes, and the iteration time
        # it predicts at the full grid (reported, never used for `value`)
        p = np.log(q['spsolve_s'] / trend[0]['spsolve_s']) / np.log(q['n'] / trend[0]['n'])
        t_full = (q['spsolve_s'] * (n_full / q['n']) ** p
                  + (q['total_s'] - q['spsolve_s']) * (n_full / q['n']))
        n15 = dict(spsolve_exponent=float(p), predicted_full_grid_seconds=float(t_full),
                   predicted_full_grid_iters_per_sec=float(1.0 / t_full))
    return dict(
        value=frac / t, unit=UNIT, cores=1, kind=kind,
        sample=('one Newton iteration (residual %.2fs + scipy chain-rule assembly %.2fs + SuperLU '
                'spsolve %.2fs) of %s on the same %s residual at a %d^2 sub-grid (%d of %d '
                'unknowns = %.3g of the workload); value = that fraction / time, i.e. time scaled '
                'LINEARLY in unknowns -- an upper bound on the reference throughput at the full grid '
                '(the measured factorisation trend is in `trend`; the full grid itself needs hours and '
                "> 70 GB, BASELINE.md 2.1); scipy's sparse kernels and SuperLU are single-threaded "
                '(1 core of %d)'
                % (r['residual_s'], r['assembly_s'], r['spsolve_s'],
                   'the UNMODIFIED reference (baseline/_ref/numpad)' if kind == 'reference'
                   else 'the oracle port (baseline/_ref absent)', workload,
                   sample_grid, r['n'], n_full, frac, os.cpu_count())),
        sample_iters_per_sec=1.0 / t, sample_grid=sample_grid, sample_seconds=t,
        sample_relres=r['relres'], trend=trend, trend_extrapolation=n15,
        spmv_gbs=(12.0 * r['nnz'] + 4.0 * (r['n'] + 1) + 16.0 * r['n']) / r['spmv_s'] / 1e9)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_baseline(args.grid, args.sample_grid, steps=args.steps, warmup=args.warmup,
                      trend_grid=args.trend_grid if args.workload == 'cavity' else 0,
                      budget=args.ref_budget, workload=args.workload)
    line = {
        'impl': 'reference', 'metric': METRIC if args.workload == 'cavity' else METRIC.replace('cavity', args.workload + '_kec'),
        'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * cb['sample_seconds'], 'higher_is_better': True,
        'scaling': 'strong' if (args.gpus > 1 and args.parallel == 'sharded') else 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.grid, args.gpus, args.parallel, args.workload),
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t0,
        'loaded_engine_modules': sorted(m for m in sys.modules if m.startswith('numpad_b200')),
    }
    print(json.dumps(line))


def cpu_baseline_subprocess(args):
    """the GPU arm's `cpu_baseline` leg: the reference arm in a fresh process (so that the
    reference never shares an interpreter with the engine), short budget"""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--grid', str(args.grid),
           '--workload', args.workload, '--steps', '1', '--warmup', '0', '--ref-budget', '25',
           '--trend-grid', '0']
    env = {k: v for k, v in os.environ.items()
           if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT')}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    for l in out.stdout.splitlines():
        if l.startswith('{'):
            return json.loads(l)['cpu_baseline']
    raise RuntimeError('cpu baseline failed: ' + out.stderr[-1000:])


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

    import ctypes
    import numpad_b200 as npd
    from numpad_b200 import linsolve, adsolve, _lib, precond, shard
    from numpad_b200.workloads.cavity import CavityProblem
    assert linsolve.DEFAULTS['rtol'] == LINEAR_RTOL, 'bench must run at the product default tolerance'
    dev = torch.device('cuda', local)
    cavity = args.workload == 'cavity'
    sharded = world > 1 and args.parallel == 'sharded' and cavity
    if cavity:
        P = CavityProblem(npd, args.grid)
        step_args = lambda uold: (npd.adarray(uold), 1.0, 0)
    else:
        from numpad_b200.workloads.kec import KecProblem
        P = KecProblem(npd, args.grid, args.grid, viscous=args.workload == 'ns')
        kec_dt = 1.0 / args.grid if args.workload == 'ns' else 0.1
        step_args = lambda uold: (npd.adarray(uold), kec_dt)
    n = P.n
    if sharded:
        # the product's own switch: solve() routes its Newton linear solves through shard.py
        linsolve.DEFAULTS['shard_owner'] = P.partition(world)
    u0 = P.fixed_state()
    u_host = torch.from_numpy(u0.copy()).pin_memory()
    uold_host = torch.from_numpy(u0.copy()).pin_memory()
    unew_host = torch.empty(n, dtype=torch.float64).pin_memory()
    u_dev = u_host.to(dev)
    uold_dev = uold_host.to(dev)
    phases = {}
    L = _lib.lib().cdll

    def newton_step(ud, uoldd):
        """the timed call: ONE Newton iteration through the drop-in API.  max_iter=2 with
        rel_tol=1: iteration 0 assembles and solves, iteration 1 re-evaluates the residual,
        sees it smaller and returns the adsolution."""
        precond.forget()       # every step pays the full preconditioner setup ...
        linsolve.forget_factors()      # ... or the full direct factorisation (Euler / NS)
        # (Euler / NS: max_iter=1 -- the same work, one iteration + the closing residual
        # evaluation + adsolution -- because a first Newton step from the perturbed uniform flow
        # need not lower the max-norm residual, which the rel_tol=1 exit of the cavity step relies on)
        return npd.solve_newton_with_dt(P.residual, npd.adarray(ud), step_args(uoldd),
                                        {}, np.inf, 2 if cavity else 1, 0.0, 1.0, False)

    def instrumented():
        """the same arithmetic phase by phase (reported in `phases`, not timed for `value`)"""
        precond.forget()
        linsolve.forget_factors()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        u = npd.adarray(u_dev)
        res = P.residual(u, *step_args(uold_dev))
        ev[1].record()
        J = npd.diff_dev(res, u)
        ev[2].record()
        L.npb_bytes_count_reset()
        du = linsolve.solve(J, res._dev.reshape(-1))
        ev[3].record()
        torch.cuda.synchronize()
        li = dict(linsolve.last_info)
        kb = float(L.npb_bytes_count())
        phases.update(residual_ms=ev[0].elapsed_time(ev[1]), assembly_ms=ev[1].elapsed_time(ev[2]),
                      linear_solve_ms=ev[2].elapsed_time(ev[3]), precond_setup_ms=li.get('setup_ms'),
                      krylov_ms=li.get('krylov_ms'), krylov_iters=li.get('iters'),
                      method=li.get('method'), nnz=J.nnz, relres=li['rnorm'] / li['bnorm'],
                      precond=li.get('precond'), krylov_algorithmic_bytes=kb)
        for k in ('permute_ms', 'localize_ms', 'halo'):
            if k in li:
                phases[k] = li[k]
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
        return newton_step(u_dev, uold_dev)

    def step_e2e():
        ud = u_host.to(dev, non_blocking=True)
        uo = uold_host.to(dev, non_blocking=True)
        sol = newton_step(ud, uo)
        unew_host.copy_(sol._dev.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        sol = step_resident()
    assert sol._n_Newton == (2 if cavity else np.inf), 'the benchmark step must be exactly one Newton iteration'
    J, du = instrumented()
    with Clocks(local) as clk:
        L.npb_launch_count_reset()
        L.npb_profile_matvec(1)            # events around the Jacobian SpMV inside the timed steps
        ms = timed(step_resident, args.steps)
        launches = int(L.npb_launch_count())
        L.npb_profile_matvec(0)
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
    if ms_spmv_insitu:
        achieved, ms_roof, how = spmv_bytes / (ms_spmv_insitu * 1e-3) / 1e9, ms_spmv_insitu, \
            'CUDA events around each Jacobian SpMV launch of the Krylov loop inside the timed steps (%d launches)' % cnt.value
    else:
        achieved, ms_roof, how = standalone, ms_spmv, 'back-to-back loop of 50 launches after the timed steps'
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(str(args.grid))
    except Exception:
        pass
    per_step = 1 if (sharded or world == 1) else world
    value = per_step * args.steps / (ms * 1e-3)
    e2e = per_step * args.steps / (ms_e2e * 1e-3)
    if rank == 0:
        cfg = workload_config(args.grid, world, args.parallel, args.workload)
        kry = None
        if phases.get('krylov_ms') and phases.get('krylov_algorithmic_bytes'):
            kgbs = phases['krylov_algorithmic_bytes'] / (phases['krylov_ms'] * 1e-3) / 1e9
            kry = {'achieved': kgbs, 'unit': 'GB/s', 'frac': kgbs / peak,
                   'bytes': phases['krylov_algorithmic_bytes'], 'ms': phases['krylov_ms'],
                   'what': 'sum of the algorithmic bytes (each operand once) of every SpMV-shaped and '
                           'vector kernel launched by the Krylov phase (operator, preconditioner, '
                           'Gram-Schmidt) / the phase time, host gaps included'}
        line = {
            'metric': METRIC if cavity else METRIC.replace('cavity', args.workload + '_kec'),
            'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if sharded else 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': 16 * n,
                    'd2h_bytes_per_step': 8 * n, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'kernel': 'k_spmv_tma (persistent TMA-pipelined CSR SpMV, here of the Jacobian)',
                         'achieved': achieved, 'peak': peak,
                         'peak_source': 'measured' if peaks else 'fallback',
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                         'bytes_per_launch': spmv_bytes, 'us_per_launch': 1e3 * ms_roof,
                         'measured': how, 'standalone_gbs': standalone,
                         'standalone_us_per_launch': 1e3 * ms_spmv,
                         'frac_of_nominal_8TBs': achieved / 8000.0,
                         'krylov_phase': kry},
            'phases': phases, 'clocks': clocks,
        }
        if not args.no_cpu_baseline and world == 1:       # rank 0 at N = 1 only
            line['cpu_baseline'] = cpu_baseline_subprocess(args)
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
