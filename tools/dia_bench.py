"""time the finest pressure-level SpMV of a cavity step in its three storages (fp64 CSR, fp32
CSR, fp32 DIA), all through the TMA / DIA kernels, mode 1 (residual) and 3 (Jacobi sweep)"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import numpad_b200 as npd
from numpad_b200 import amg, _dev
from numpad_b200._dev import _call
from numpad_b200.workloads.cavity import CavityProblem
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = CavityProblem(npd, N); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized()
M = amg.LscPreconditioner(J)
L = M.Lm
n = L.shape[0]
x = torch.randn(n, dtype=torch.float64, device='cuda'); b = torch.randn_like(x); dinv = torch.rand_like(x); y = torch.empty_like(x)
offs, vals, ld, _ = amg.dia_image(L)
off_h = (ctypes.c_int32 * 8)(*(offs + [0] * (8 - len(offs))))
v32 = L.data.float()
def timeit(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
print('L: n = %d, nnz = %d, offsets %s' % (n, L.nnz, offs))
from numpad_b200 import _lib
for rows in (1, 2):
  _lib.lib().cdll.npb_spmv_dia_set_rows(rows)
  print('rows per thread', rows)
  for mode in (0, 1, 3):
      t64 = timeit(lambda: _call('npb_spmv_epilogue', n, L.nnz, L.indptr, L.indices, L.data, x, mode, b, dinv, 0.67, y, 0))
      t32 = timeit(lambda: _call('npb_spmv_f32', n, L.nnz, L.indptr, L.indices, v32, x, mode, b, dinv, 0.67, y))
      td = timeit(lambda: _call('npb_spmv_dia', n, n, len(offs), ctypes.cast(off_h, ctypes.c_void_p).value, vals, ld, x, mode, b, dinv, 0.67, y))
      vec = 8.0 * n * (2 + (mode >= 1) + 2 * (mode == 3))
      print('mode %d: csr64 %.1f us (%.0f GB/s)  csr32 %.1f us (%.0f GB/s)  dia32 %.1f us (%.0f GB/s)' % (
          mode, t64, (12.0 * L.nnz + 4 * n + vec) / t64 / 1e3, t32, (8.0 * L.nnz + 4 * n + vec) / t32 / 1e3,
          td, (4.0 * len(offs) * n + vec) / td / 1e3))
  # fused kernels of the pressure CG
  red = torch.zeros(_lib.lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8, device='cuda')
  dot = torch.zeros(2, dtype=torch.float64, device='cuda'); rz = torch.ones(2, dtype=torch.float64, device='cuda')
  p = torch.randn_like(x); q = torch.randn_like(x)
  offp = ctypes.cast(off_h, ctypes.c_void_p).value
  for ctas in (0, 2, 3, 4):
    _lib.lib().cdll.npb_spmv_dia_set_dot_ctas(ctas)
    tj = timeit(lambda: _call('npb_spmv_dia_jacobi_dot', n, n, len(offs), offp, vals, ld, x, b, dinv, 0.67, y, dot, red))
    tc = timeit(lambda: _call('npb_dia_cg_dir', n, len(offs), offp, vals, ld, x, rz[0:], rz[1:], 0, p, q, dot, red))
    print('%2d CTAs/SM: jacobi + dot %.1f us (%.0f GB/s)   cg direction (p, q, p.q) %.1f us (%.0f GB/s)' % (
        ctas, tj, (4.0 * len(offs) * n + 32.0 * n) / tj / 1e3, tc, (4.0 * len(offs) * n + 40.0 * n) / tc / 1e3))
