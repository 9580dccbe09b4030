"""One Newton iteration of the cavity benchmark state, bracketed by
cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures:

  python tools/profile_step.py --grid 1024            (plain run: prints phase times)
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --grid 1024
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import numpad_b200 as npd  # noqa: E402
from numpad_b200 import linsolve, precond  # noqa: E402
from numpad_b200.workloads.cavity import CavityProblem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=1024)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--rtol', type=float, default=1e-12)
ap.add_argument('--only', default='all', help='all | assembly | solve')
a = ap.parse_args()

P = CavityProblem(npd, a.grid)
u0 = P.fixed_state()
ud = torch.from_numpy(u0).cuda()


def step(profile=False):
    t = [time.perf_counter()]
    def mark():
        torch.cuda.synchronize()
        t.append(time.perf_counter())
    if profile and a.only in ('all', 'assembly'):
        torch.cuda.profiler.start()
    u = npd.adarray(ud)
    res = P.residual(u, npd.adarray(ud.clone()), 1.0, 0)
    mark()
    J = npd.diff_dev(res, u)
    mark()
    if profile and a.only == 'assembly':
        torch.cuda.profiler.stop()
    if profile and a.only == 'solve':
        torch.cuda.profiler.start()
    precond._last['lsc'] = None
    du = linsolve.solve(J, res._dev.reshape(-1), rtol=a.rtol)
    mark()
    if profile and a.only in ('all', 'solve'):
        torch.cuda.profiler.stop()
    return [1e3 * (t[i + 1] - t[i]) for i in range(3)], dict(linsolve.last_info)


for _ in range(a.warmup):
    step()
ms, info = step(profile=True)
print('grid %d: residual %.2f ms, assembly %.2f ms, linear solve %.2f ms' % (a.grid, *ms))
print('linear solve info:', info)
print('kernel launches so far:', npd.launch_count())
