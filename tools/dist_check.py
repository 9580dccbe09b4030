"""Row-sharded Newton linear solve vs the single-GPU solve (run under torchrun):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/dist_check.py --grid 256
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
ap = argparse.ArgumentParser(); ap.add_argument('--grid', type=int, default=256)
ap.add_argument('--reps', type=int, default=2)
a = ap.parse_args()
local = int(os.environ.get('LOCAL_RANK', 0)); rank = int(os.environ.get('RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import numpad_b200 as npd
from numpad_b200 import linsolve, parallel
from numpad_b200.workloads.cavity import CavityProblem
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized(); b = res._dev.reshape(-1).contiguous()
owner = torch.from_numpy(P.partition(dist.get_world_size()))
for rep in range(a.reps):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    x, info = parallel.solve_sharded(J, b, owner, rtol=1e-10)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    r = b - J.matvec(x)
    if rank == 0:
        print('grid %d world %d: sharded solve %.1f ms, its %d status %d, true relres %.2e, halo exchanges %d (%d B each)'
              % (a.grid, dist.get_world_size(), 1e3 * (t1 - t0), info['iters'], info['status'],
                 float(r.norm() / b.norm()), info['halo_exchanges'], info['halo_bytes']), flush=True)
if rank == 0:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x1 = linsolve.solve(J, b, rtol=1e-10, precond='lsc', restart=100)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print('single GPU: %.1f ms, its %d; max|x_sharded - x_single|/max|x| = %.2e'
          % (1e3 * (t1 - t0), linsolve.last_info['iters'], float((x - x1).abs().max() / x1.abs().max())), flush=True)
dist.barrier()
dist.destroy_process_group()
