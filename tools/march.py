"""The reference's cavity driver loop (test/cavity.py:97-114) on the device at
an arbitrary grid, with per-solve linear-solver reports:
  python tools/march.py --grid 256 --steps 3
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import linsolve
from numpad_b200.workloads.cavity import CavityProblem
ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=256); ap.add_argument('--steps', type=int, default=3)
a = ap.parse_args()
linsolve.DEFAULTS['verbose'] = True
linsolve.DEFAULTS['rtol'] = 1e-10
P = CavityProblem(npd, a.grid)
q = npd.zeros(P.n); force = npd.zeros(P.n); t, dt = 0, 1.
for step in range(a.steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    print('t =', t, 'dt =', dt, flush=True)
    q = npd.solve(P.residual, q, args=(q, dt, force), rel_tol=0.05, abs_tol=1e-8)
    torch.cuda.synchronize()
    print('   -> %d Newton iterations in %.2f s' % (q._n_Newton, time.perf_counter() - t0), flush=True)
    if q._n_Newton == 1: break
    elif q._n_Newton < 4: dt *= 2
    t += dt
    q.obliviate()
