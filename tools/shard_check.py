"""Row-sharded linear solve vs the single-GPU solve on the same cavity Jacobian.
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_check.py --grid 256
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=256); ap.add_argument('--reps', type=int, default=2)
ap.add_argument('--no-single', action='store_true')
a = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import numpad_b200 as npd
from numpad_b200 import linsolve, shard, precond
from numpad_b200.workloads.cavity import CavityProblem
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized(); b = res._dev.reshape(-1).contiguous()
owner = P.partition(world)
x1 = None
if not a.no_single:
    for _ in range(a.reps):
        precond._last['lsc'] = None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        x1 = linsolve.solve(J, b, rtol=1e-10, restart=100)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    li = dict(linsolve.last_info)
    if rank == 0:
        print('single : %.1f ms (setup %.1f krylov %.1f) its %d relres %.2e' % (
            1e3 * (t1 - t0), li['setup_ms'], li['krylov_ms'], li['iters'], li['rnorm'] / li['bnorm']), flush=True)
for _ in range(a.reps):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    x2 = shard.solve(J, b, owner=owner, rtol=1e-10, restart=100)
    torch.cuda.synchronize(); t1 = time.perf_counter()
si = dict(shard.last_info)
r = b - J.matvec(x2)
rel = float(r.norm() / b.norm())
if rank == 0:
    print('sharded x%d: %.1f ms (permute %.1f setup %.1f localize %.1f krylov %.1f gather %.1f) its %d true relres %.2e first_rep %s halo %s'
          % (world, 1e3 * (t1 - t0), si['permute_ms'], si['setup_ms'], si['localize_ms'], si['krylov_ms'], si['gather_ms'],
             si['iters'], rel, si['first_replicated_level'], si['halo']), flush=True)
    if x1 is not None:
        print('        |x_sharded - x_single| / |x| = %.2e' % float((x2 - x1).norm() / x1.norm()), flush=True)
assert rel < 1e-9, rel
shard.shutdown()
dist.destroy_process_group()
