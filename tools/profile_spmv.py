"""SpMV of the cavity Jacobian in isolation (for ncu --set full):
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:k_spmv_stream -c 3 -o gpurun_out/prof python tools/profile_spmv.py --grid 1024
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd
from numpad_b200.workloads.cavity import CavityProblem
ap = argparse.ArgumentParser(); ap.add_argument('--grid', type=int, default=1024)
a = ap.parse_args()
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized()
x = torch.randn(P.n, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
for _ in range(5):
    J.matvec(x, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    J.matvec(x, out=y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
b = 12.0 * J.nnz + 4.0 * (P.n + 1) + 16.0 * P.n
print('grid %d: SpMV %.1f us, %.0f GB/s algorithmic (%d rows, %d nnz)' % (a.grid, 1e3 * ms, b / ms / 1e6, P.n, J.nnz))
from numpad_b200 import _dev
for variant in (1, 3, 4):
    for _ in range(3):
        _dev._call('npb_spmv_variant', P.n, J.nnz, J.indptr, J.indices, J.data, x, y, variant)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50):
        _dev._call('npb_spmv_variant', P.n, J.nnz, J.indptr, J.indices, J.data, x, y, variant)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print('  variant %d: %.1f us, %.0f GB/s' % (variant, 1e3 * ms, b / ms / 1e6))
torch.cuda.profiler.start()
for _ in range(3):
    J.matvec(x, out=y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
