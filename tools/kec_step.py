"""One implicit Newton iteration of the euler_kec / navierstokes residual (configs 4 / 5)
on one GPU at the synthetic duct mesh: residual, Jacobian assembly, linear solve.
  python tools/kec_step.py --ni 512 --nj 512 [--viscous] [--precond auto]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import linsolve
from numpad_b200.workloads.kec import KecProblem
ap = argparse.ArgumentParser()
ap.add_argument('--ni', type=int, default=256); ap.add_argument('--nj', type=int, default=256)
ap.add_argument('--viscous', action='store_true'); ap.add_argument('--dt', type=float, default=0.0)
ap.add_argument('--precond', default='auto'); ap.add_argument('--rtol', type=float, default=1e-10)
ap.add_argument('--reps', type=int, default=2)
a = ap.parse_args()
P = KecProblem(npd, a.ni, a.nj, viscous=a.viscous)
dt = a.dt or (1.0 / a.nj if a.viscous else 0.1)
w0 = P.fixed_state()
ev = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(a.reps):
    torch.cuda.synchronize(); e = [ev() for _ in range(4)]
    e[0].record()
    w = npd.array(w0.copy()); res = P.residual(w, npd.array(w0), dt)
    e[1].record()
    J = npd.diff_dev(res, w)
    e[2].record()
    x = linsolve.solve(J, res._dev.reshape(-1), rtol=a.rtol, precond=a.precond, restart=100, maxiter=1000)
    e[3].record(); torch.cuda.synchronize()
    li = linsolve.last_info
    print('%s %dx%d n=%d nnz=%d dt=%g | residual %.1f ms assembly %.1f ms solve %.1f ms (setup %.1f, krylov %.1f) | %s its %d relres %.2e | mem %.1f GB'
          % ('ns' if a.viscous else 'euler', a.ni, a.nj, P.n, J.nnz, dt, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]),
             e[2].elapsed_time(e[3]), li.get('setup_ms', 0), li.get('krylov_ms', 0), li.get('method'), li.get('iters', 0),
             li['rnorm'] / li['bnorm'], torch.cuda.max_memory_allocated() / 1e9), flush=True)
    if li.get('precond'):
        print('   ', li['precond'])
