"""Sweep LSC/AMG parameters on one cavity Jacobian (GPU): iterations and times."""
import argparse, os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import linsolve, amg
from numpad_b200.workloads.cavity import CavityProblem
ap = argparse.ArgumentParser(); ap.add_argument('--grid', type=int, default=512)
ap.add_argument('--configs', default='2,3,1,1;2,3,1,0;2,2,1,1;3,3,1,1;1,3,1,1')
ap.add_argument('--rtol', type=float, default=1e-12); ap.add_argument('--restart', type=int, default=100)
a = ap.parse_args()
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized(); b = res._dev.reshape(-1).contiguous()
for cfg in a.configs.split(';'):
    vals = cfg.split(',')
    ptr_ = float(vals[6]) if len(vals) > 6 else 0.0
    agl = vals[7] if len(vals) > 7 else 'mis1'
    aga = vals[8] if len(vals) > 8 else 'mis1'
    nua = int(vals[9]) if len(vals) > 9 else 2
    gl = int(vals[10]) if len(vals) > 10 else 1
    ga = int(vals[11]) if len(vals) > 11 else 1
    gf = int(vals[12]) if len(vals) > 12 else 2
    vals[1] = str(int(vals[1].split('+')[0]) | (int(vals[1].split('+')[1]) << 16)) if '+' in vals[1] else vals[1]
    ca, cl, nul, sm, ap_, apa = ([int(v) for v in vals[:6]] + [1, 1])[:6]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    M = amg.LscPreconditioner(J, cycles_a=ca, cycles_l=cl, nu_l=nul, smooth_a=bool(sm), agg_passes=ap_, agg_passes_a=apa, p_trunc=ptr_, agg_l=agl, agg_a=aga, nu_a=nua, gamma_l=gl, gamma_a=ga, gamma_from=gf)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    x, info = linsolve.fgmres(J, b, precond=M, rtol=a.rtol, restart=a.restart, maxiter=600)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print('agg %s/%s nu_a=%d gamma(L,A,from)=%d,%d,%d ' % (agl, aga, nua, gl, ga, gf), end='')
    print('grid %d cycles_a=%d pcg_l=%x nu_l=%d smooth_a=%d passes(L,A)=%d,%d trunc=%.2f: its %d status %d setup %.0f ms krylov %.0f ms | %s'
          % (a.grid, ca, cl, nul, sm, ap_, apa, ptr_, info['iters'], info['status'], 1e3*(t1-t0), 1e3*(t2-t1), M.describe()), flush=True)
