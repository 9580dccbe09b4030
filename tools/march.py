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
a = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
if world > 1:      # torchrun: the unmodified driver loop with row-sharded linear solves
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
linsolve.DEFAULTS['verbose'] = rank == 0
linsolve.DEFAULTS['rtol'] = 1e-10
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
    q.obliviate()
if world > 1:
    from numpad_b200 import shard
    shard.shutdown()
    dist.destroy_process_group()
