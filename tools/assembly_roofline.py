"""Per-family time, algorithmic bytes and achieved GB/s of the chain-rule kernels
during one real Jacobian assembly (cavity, adjoint order as inside Newton).

Algorithmic bytes: CSR operand = 12 B/nnz + 4 B/(row+1); diagonal
8 B/row; selection map 4 B/entry (+8 with weights); every operand counted once.
Each family is timed with CUDA events around the host-level op (it includes the
symbolic phase, the scan and the host read of the nnz)."""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import _dev
from numpad_b200.workloads.cavity import CavityProblem

ap = argparse.ArgumentParser(); ap.add_argument('--grid', type=int, default=1024)
a = ap.parse_args()
stats = collections.defaultdict(lambda: [0, 0.0, 0.0])


def csr_bytes(M):
    return 12.0 * M.nnz + 4.0 * (M.shape[0] + 1)


def timed(name, fn, bytes_fn):
    def wrap(*args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args)
        e1.record()
        torch.cuda.synchronize()
        s = stats[name]
        s[0] += 1
        s[1] += e0.elapsed_time(e1)
        s[2] += bytes_fn(out, *args)
        return out
    return wrap


def sel_bytes(S):
    return (4.0 + (8.0 if S.w is not None else 0.0)) * S.n_rows


def relabel_bytes(C, A, S):
    if C.indptr.data_ptr() == A.indptr.data_ptr() and C.data.data_ptr() == A.data.data_ptr():
        return 8.0 * A.nnz + sel_bytes(S)          # indices in + indices out only
    return csr_bytes(A) + sel_bytes(S) + csr_bytes(C)


_dev._csr_plus_csr = timed('spadd  (csr + csr, prune)', _dev._csr_plus_csr,
                           lambda C, A, B: csr_bytes(A) + csr_bytes(B) + csr_bytes(C))
_dev._csr_times_diag = timed('csr x diag (column scale)', _dev._csr_times_diag,
                             lambda C, A, d: 20.0 * A.nnz + 8.0 * A.shape[1])
_dev._csr_times_sel = timed('csr x selection (relabel)', _dev._csr_times_sel, relabel_bytes)
_dev._sel_times_csr = timed('selection x csr (row gather)', _dev._sel_times_csr,
                            lambda C, S, B: csr_bytes(B) + sel_bytes(S) + csr_bytes(C))
_dev._csr_times_csr = timed('general csr x csr (expand+sort)', _dev._csr_times_csr,
                            lambda C, A, B: csr_bytes(A) + csr_bytes(B) + csr_bytes(C))

P = CavityProblem(npd, a.grid)
u0 = P.fixed_state()
for rep in range(2):
    stats.clear()
    u = npd.array(u0.copy())
    res = P.residual(u, npd.array(u0), 1.0, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    J = npd.diff_dev(res, u)
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
peak = 6545.9
print('cavity %d^2, n = %d, nnz(J) = %d: assembly %.2f ms (with per-op synchronisation for this table)'
      % (a.grid, P.n, J.nnz, total))
print('%-34s %6s %10s %12s %10s %8s' % ('family', 'calls', 'ms', 'alg. GB', 'GB/s', 'of peak'))
tb = tt = 0
for name, (n, ms, b) in sorted(stats.items(), key=lambda kv: -kv[1][1]):
    print('%-34s %6d %10.3f %12.3f %10.0f %7.1f%%' % (name, n, ms, b / 1e9, b / ms / 1e6, 100 * b / ms / 1e6 / peak))
    tb += b; tt += ms
print('%-34s %6s %10.3f %12.3f %10.0f %7.1f%%' % ('all families', '', tt, tb / 1e9, tb / tt / 1e6, 100 * tb / tt / 1e6 / peak))
print('bytes per unknown: %.0f (reference formulation: ~8600 B/unknown, SURVEY fact 3); J itself: %.0f B/unknown'
      % (tb / P.n, (12.0 * J.nnz + 4 * P.n) / P.n))
