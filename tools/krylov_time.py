"""one cavity Newton linear solve through the product path (linsolve.solve), repeated: ms per
solve, Krylov ms, iterations, relative residual; optional tuning switches

  python tools/krylov_time.py --grid 2048 [--f32three 1] [--reps 3]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import linsolve, precond, _lib
from numpad_b200.workloads.cavity import CavityProblem
ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=2048)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--f32three', type=int, default=-1)
ap.add_argument('--coarse-max', type=int, default=0)
ap.add_argument('--f32-min-rows', type=int, default=-1)
a = ap.parse_args()
if a.f32three >= 0:
    _lib.lib().cdll.npb_spmv_set_tma_f32_three(a.f32three)
if a.coarse_max:
    from numpad_b200 import amg
    amg.COARSE_MAX = a.coarse_max
if a.f32_min_rows >= 0:
    from numpad_b200 import amg as _amg
    _amg.F32_MIN_ROWS = a.f32_min_rows
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized(); b = res._dev.reshape(-1).contiguous()
for rep in range(a.reps + 1):
    precond.forget()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x = linsolve.solve(J, b)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    i = linsolve.last_info
    relres = float((b - J.matvec(x)).norm() / b.norm())
    print('%s solve %.1f ms, setup %.1f, krylov %.1f ms, %d iterations, true relres %.2e (%s)' % (
        'warm-up' if rep == 0 else 'rep %d ' % rep, 1e3 * (t1 - t0), i.get('setup_ms', float('nan')),
        i.get('krylov_ms', float('nan')), i.get('iters', -1), relres, i.get('method')), flush=True)
