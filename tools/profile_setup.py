"""Where the preconditioner setup time goes: wraps the device building blocks of
numpad_b200/amg.py with synchronising timers and prints one line per call.

  python tools/profile_setup.py --grid 2048
"""
import argparse, os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import amg, _dev
from numpad_b200.workloads.cavity import CavityProblem

ap = argparse.ArgumentParser(); ap.add_argument('--grid', type=int, default=1024)
ap.add_argument('--quiet', action='store_true')
ap.add_argument('--warp-rows', type=int, default=-1)
a = ap.parse_args()
if a.warp_rows >= 0:
    from numpad_b200 import _lib
    _lib.lib().cdll.npb_spgemm_set_warp_rows(a.warp_rows)
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized()
amg.LscPreconditioner(J)          # warm-up (allocator, kernels)
tot = collections.defaultdict(float); cnt = collections.Counter()
depth = [0]


def wrap(mod, name):
    f = getattr(mod, name)

    def g(*args, **kw):
        torch.cuda.synchronize(); t0 = time.perf_counter(); depth[0] += 1
        try:
            r = f(*args, **kw)
        finally:
            depth[0] -= 1
        torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t0)
        if depth[0] == 0:
            tot[name] += dt; cnt[name] += 1
            if not a.quiet:
                desc = ' '.join('%sx%s/%d' % (x.shape[0], x.shape[1], x.nnz) for x in args if hasattr(x, 'nnz'))
                out = ' -> %sx%s/%d' % (r.shape[0], r.shape[1], r.nnz) if hasattr(r, 'nnz') else ''
                print('  %-18s %8.2f ms  %s%s' % (name, dt, desc, out))
        return r
    setattr(mod, name, g)


for n in ('_csr_times_csr', '_csr_times_sel', '_csr_plus_csr', '_diag_times_csr', '_sel_times_csr',
          'coo_to_csr', 'aggregate', 'filtered', 'truncated', 'diagonal'):
    wrap(amg, n)
orig_T = _dev.DevCsr.transpose


def timed_T(self):
    torch.cuda.synchronize(); t0 = time.perf_counter(); depth[0] += 1
    try:
        r = orig_T(self)
    finally:
        depth[0] -= 1
    torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t0)
    if depth[0] == 0:
        tot['transpose'] += dt; cnt['transpose'] += 1
        if not a.quiet:
            print('  %-18s %8.2f ms  %sx%s/%d' % ('transpose', dt, self.shape[0], self.shape[1], self.nnz))
    return r


_dev.DevCsr.transpose = timed_T
torch.cuda.synchronize(); t0 = time.perf_counter()
M = amg.LscPreconditioner(J)
torch.cuda.synchronize(); T = 1e3 * (time.perf_counter() - t0)
print('setup total (with timer syncs) %.1f ms | %s' % (T, M.describe()))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('%-18s %8.1f ms %5.1f%%  n=%d' % (k, v, 100 * v / T, cnt[k]))
print('%-18s %8.1f ms' % ('(other)', T - sum(tot.values())))
