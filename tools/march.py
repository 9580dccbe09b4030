"""The reference's cavity driver loop (test/cavity.py:97-114) on the device at
an arbitrary grid, with per-solve linear-solver reports:
  python tools/march.py --grid 256 --steps 3
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import numpad_b200 as npd
from numpad_b200 import linsolve
from numpad_b200.workloads.cavity import CavityProblem
ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=256); ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--dt', type=float, default=1.0)
ap.add_argument('--adjoint', action='store_true', help='config 3: d(kinetic energy)/d(force) through the last solve (test/cavity_adjoint.py:6-8)')
ap.add_argument('--rtol', type=float, default=1e-12)
a = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
if world > 1:      # torchrun: the unmodified driver loop with row-sharded linear solves
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
linsolve.DEFAULTS['verbose'] = rank == 0
linsolve.DEFAULTS['rtol'] = a.rtol
P = CavityProblem(npd, a.grid)
if world > 1:
    linsolve.DEFAULTS['shard_owner'] = P.partition(world)
q = npd.zeros(P.n); force = npd.zeros(P.n); t, dt = 0, a.dt
for step in range(a.steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if rank == 0:
        print('t =', t, 'dt =', dt, flush=True)
    q = npd.solve(P.residual, q, args=(q, dt, force), rel_tol=0.05, abs_tol=1e-8)
    torch.cuda.synchronize()
    if rank == 0:
        print('   -> %d Newton iterations in %.2f s, |q| = %.12e' % (q._n_Newton, time.perf_counter() - t0, float(np.linalg.norm(q._value))), flush=True)
    if q._n_Newton == 1: break
    elif q._n_Newton < 4: dt *= 2
    t += dt
    if step < a.steps - 1:
        q.obliviate()
if a.adjoint:
    N = a.grid
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ke = 0.5 * npd.sum(q[N * N:] ** 2)
    adj = np.asarray(ke.diff(force).todense()).ravel()
    torch.cuda.synchronize()
    if rank == 0:
        print('adjoint of the kinetic energy w.r.t. the force: |adj| = %.12e in %.2f s (%s, %d its)'
              % (float(np.linalg.norm(adj)), time.perf_counter() - t0, linsolve.last_info.get('method'),
                 linsolve.last_info.get('iters', -1)), flush=True)
if world > 1:
    from numpad_b200 import shard
    shard.shutdown()
    dist.destroy_process_group()
