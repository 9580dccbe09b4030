"""time the momentum block of a cavity step (Jacobi epilogue) from fp32 CSR (TMA kernel) and from
its hybrid diagonal + ELL image, one and two rows per thread"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200 import amg, _lib
from numpad_b200._dev import _call
from numpad_b200.workloads.cavity import CavityProblem
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = CavityProblem(npd, N); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized()
M = amg.LscPreconditioner(J)
A = M.A.materialized()
n = A.shape[0]
x = torch.randn(n, dtype=torch.float64, device='cuda'); b = torch.randn_like(x); dinv = torch.rand_like(x); y = torch.empty_like(x)
amg.HYB_IMAGES = True
img = amg.dia_image(A)
offs, vals, ld, (w, eidx, eval_) = img
off_h = (ctypes.c_int32 * 8)(*(offs + [0] * (8 - len(offs))))
offp = ctypes.cast(off_h, ctypes.c_void_p).value
v32 = A.data.float()
def timeit(fn, reps=100):
    for _ in range(10): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
print('A: n = %d, nnz = %d, %d diagonals %s, ELL width %d' % (n, A.nnz, len(offs), offs, w))
for mode in (0, 1, 3):
    vec = 8.0 * n * (2 + (mode >= 1) + 2 * (mode == 3))
    t32 = timeit(lambda: _call('npb_spmv_f32', n, A.nnz, A.indptr, A.indices, v32, x, mode, b, dinv, 0.67, y))
    out = 'mode %d: csr32 %.1f us (%.0f GB/s)' % (mode, t32, (8.0 * A.nnz + 4 * n + vec) / t32 / 1e3)
    for rows in (2,):
        th = timeit(lambda: _call('npb_spmv_hyb', n, n, len(offs), offp, vals, ld, w, eidx, eval_, x, mode, b, dinv, 0.67, y))
        out += '  hyb R=%d %.1f us (%.0f GB/s)' % (rows, th, ((4.0 * len(offs) + 8.0 * w) * n + vec) / th / 1e3)
    print(out)
