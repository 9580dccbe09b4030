"""time the Gram-Schmidt kernels of the Krylov loop (multi-dot, multi-axpy) at the sizes of the
2048^2 cavity step: achieved GB/s on the algorithmic bytes 8 n (k + 1) / 8 n (k + 2)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpad_b200 as npd  # noqa: F401
from numpad_b200 import _lib
from numpad_b200._dev import _call

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12578816
ld = (n + 15) // 16 * 16
KMAX = 66
V = torch.randn(KMAX * ld, dtype=torch.float64, device='cuda')
w = torch.randn(ld, dtype=torch.float64, device='cuda')
h = torch.zeros(KMAX, dtype=torch.float64, device='cuda')
out = torch.zeros(KMAX + 1, dtype=torch.float64, device='cuda')
red = torch.zeros(_lib.lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8, device='cuda')


def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


for nn in (n, 4194303):
    print('n = %d' % nn)
    for k in (1, 2, 4, 6, 8, 16, 17, 32, 33, 48, 66):
        td = timeit(lambda: _call('npb_dot_multi', nn, k, V, ld, w, out, red))
        ta = timeit(lambda: _call('npb_axpy_multi', nn, k, V, ld, h, -1.0, w, out, red))
        print('  k = %2d: dot %7.1f us (%5.0f GB/s)   axpy %7.1f us (%5.0f GB/s)' % (
            k, td, 8.0 * nn * (k + 1) / td / 1e3, ta, 8.0 * nn * (k + 2) / ta / 1e3))
    ref = ld * 8 * 34
    x = V[:ld * 17]; y = V[ld * 17: ld * 34]
    tc = timeit(lambda: y.copy_(x))
    print('  torch copy of 17 vectors: %.1f us (%.0f GB/s)' % (tc, 2 * 17 * ld * 8 / tc / 1e3))
