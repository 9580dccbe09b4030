"""Time every SpMV kernel variant on the operators of a cavity Newton step: the Jacobian,
its momentum / pressure blocks and the multigrid level operators (A, P, R).

  python tools/spmv_bench.py --grid 2048
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import _dev, amg
from numpad_b200.workloads.cavity import CavityProblem

ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=1024)
ap.add_argument('--variants', default='0,1,4,5')
ap.add_argument('--levels', type=int, default=3)
ap.add_argument('--geoms', default='')
a = ap.parse_args()
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized()
M = amg.LscPreconditioner(J)
ops = [('J', J), ('A', M.A), ('G', M.G), ('D', M.D), ('L', M.Lm)]
for name, h in (('A', M.amg_a), ('L', M.amg_l)):
    for l, lev in enumerate(h.levels[:a.levels]):
        if l > 0:
            ops.append(('%s.A%d' % (name, l), lev['A']))
        if 'P' in lev:
            ops.append(('%s.P%d' % (name, l), lev['P']))
            ops.append(('%s.R%d' % (name, l), lev['R']))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
variants = [int(v) for v in a.variants.split(',')] + [50 + int(g) for g in a.geoms.split(',') if g]
from numpad_b200 import _lib
peak = 6545.9
print('%-8s %10s %11s %6s | %s' % ('op', 'rows', 'nnz', 'nnz/r', '  '.join('v%d us  GB/s frac' % v for v in variants)))
for name, A in ops:
    A = A.materialized()
    m, n = A.shape
    x = torch.randn(n, dtype=torch.float64, device='cuda'); y = torch.empty(m, dtype=torch.float64, device='cuda')
    by = 12.0 * A.nnz + 4.0 * (m + 1) + 8.0 * n + 8.0 * m
    out = []
    ref = None
    for v in variants:
        if v >= 50:
            _lib.lib().cdll.npb_spmv_set_tma_geometry(v - 50); v = 5
        run = lambda v=v: _dev._call('npb_spmv_variant', m, A.nnz, A.indptr, A.indices, A.data, x, y, v)
        for _ in range(3):
            run()
        torch.cuda.synchronize(); e0.record()
        reps = 30
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if ref is None:
            ref = y.clone()
        err = float((y - ref).abs().max() / (ref.abs().max() + 1e-300))
        out.append('%7.1f %5.0f %.2f%s' % (1e3 * ms, by / ms / 1e6, by / ms / 1e6 / peak, '' if err < 1e-12 else ' ERR %.1e' % err))
    print('%-8s %10d %11d %6.1f | %s' % (name, m, A.nnz, A.nnz / m, '  '.join(out)), flush=True)
