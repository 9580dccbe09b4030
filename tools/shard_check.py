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
ap.add_argument('--strict', action='store_true', help='assert agreement with the single-GPU solve and identical replicated operators (pytest)')
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
shard.DIGESTS = a.strict
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
if a.strict:
    # (1) every rank must hold bit-identical replicated operators: `localize` derives the halo
    # lists independently per rank from them
    def digest(t):
        v = t.reshape(-1)
        v = v.view(torch.int64) if v.dtype == torch.float64 else v.to(torch.int64)
        w = torch.arange(1, v.numel() + 1, device=v.device, dtype=torch.int64)
        return int(((v * (w % 1000003 + 1)).sum() % 9223372036854775783).item())
    sig = [digest(J.indptr), digest(J.indices), digest(J.data), digest(b)]
    for name in ('lsc_digests',):
        sig += list(si.get(name, []))
    t = torch.tensor(sig, dtype=torch.int64, device='cuda')
    allsig = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allsig, t)
    assert all(torch.equal(allsig[0], s_) for s_ in allsig), 'replicated operators differ across ranks'
    # (2) the sharded solution is the single-GPU solution
    assert x1 is not None
    diff = float((x2 - x1).norm() / x1.norm())
    assert diff < 1e-7 and abs(si['iters'] - li['iters']) <= 3, (diff, si['iters'], li['iters'])
    xs = [torch.empty_like(x2) for _ in range(world)]
    dist.all_gather(xs, x2)
    assert all(torch.equal(xs[0], x_) for x_ in xs), 'ranks returned different solutions'
    if rank == 0:
        print('replicated operators identical on all ranks (%d digests)' % len(sig), flush=True)
        print('STRICT OK: |x_sharded - x_single|/|x| = %.2e, its %d vs %d' % (diff, si['iters'], li['iters']), flush=True)
shard.shutdown()
dist.destroy_process_group()
