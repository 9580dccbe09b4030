"""Per-operator cost of the row-sharded SpMV (halo exchange + local kernel) against the
single-GPU kernel on the full operator.  torchrun --nproc-per-node N tools/shard_micro.py"""
import argparse, os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
ap = argparse.ArgumentParser(); ap.add_argument('--grid', type=int, default=2048)
a = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import numpad_b200 as npd
from numpad_b200 import shard, amg, _lib, _dev
from numpad_b200._dev import coo_to_csr, stream_ptr
from numpad_b200.workloads.cavity import CavityProblem
P = CavityProblem(npd, a.grid); u0 = P.fixed_state()
u = npd.array(u0.copy()); res = P.residual(u, npd.array(u0), 1.0, 0)
J = npd.diff_dev(res, u).materialized()
handle = shard.init()
torch.manual_seed(0)
perm, inv, cuts = shard.strip_permutation(torch.from_numpy(P.partition(world)).cuda(), world)
Jp = shard.permuted(J, perm, inv)
M = amg.LscPreconditioner(Jp)
shard._heap_reset()
S = shard.ShardedLsc(M, cuts)
L = _lib.lib().cdll
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if rank == 0:
    print('exchange:', 'p2p' if shard._handle['p2p'] else 'nccl', 'world', world, flush=True)


def timeit(fn, reps=100):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def bench(name, glob, loc, rc, cc):
    g = glob.materialized()
    xg = torch.randn(g.shape[1], dtype=torch.float64, device='cuda'); yg = torch.empty(g.shape[0], dtype=torch.float64, device='cuda')
    t_full = timeit(lambda: g.matvec(xg, out=yg))
    m = amg.CsrMat(); shard._fill_mat(m, loc, handle)
    xl = xg[cc[rank]:cc[rank + 1]].contiguous() if loc.sharded_cols else xg
    yl = torch.empty(loc.shape[0], dtype=torch.float64, device='cuda')
    ctx = ctypes.c_void_p(ctypes.addressof(m))
    t_sh = timeit(lambda: L.npb_mat_matvec_cb(ctx, xl.data_ptr(), yl.data_ptr(), stream_ptr()))
    # local kernel alone: same matrix without the exchange
    m2 = amg.CsrMat(); shard._fill_mat(m2, loc, handle); m2.halo.dist = None
    ctx2 = ctypes.c_void_p(ctypes.addressof(m2))
    xe = torch.randn(loc.shape[1], dtype=torch.float64, device='cuda')
    t_loc = timeit(lambda: L.npb_mat_matvec_cb(ctx2, xe.data_ptr(), yl.data_ptr(), stream_ptr()))
    t_ex = timeit(lambda: L.npb_mat_halo_exchange(ctx, xl.data_ptr(), stream_ptr()))
    L.npb_mat_matvec_cb(ctx, xl.data_ptr(), yl.data_ptr(), stream_ptr()); torch.cuda.synchronize()
    err = float((yl - yg[rc[rank]:rc[rank + 1]]).abs().max() / (yg.abs().max() + 1e-300))
    if rank == 0:
        print('%-8s rows %9d full %7.1f us | ideal %7.1f | sharded %7.1f (local kernel %7.1f, exchange ~%5.1f, alone %5.1f) maxb %d n_send %d err %.1e'
              % (name, g.shape[0], t_full, t_full / world, t_sh, t_loc, t_sh - t_loc, t_ex, loc.maxb, int(loc.send_idx.numel()), err), flush=True)


Jl = shard._loc(Jp, cuts, cuts)
bench('J', Jp, Jl, cuts, cuts)
bench('A', M.A, S.A, S.cv, S.cv); bench('G', M.G, S.G, S.cv, S.cp); bench('D', M.D, S.D, S.cp, S.cv); bench('L', M.Lm, S.Lm, S.cp, S.cp)
for name, sh in (('A', S.amg_a), ('L', S.amg_l)):
    k = 0
    for l in range(sh.first_rep):
        lev = sh.h.levels[l]; c, cn = sh.lev_cuts[l], sh.lev_cuts[l + 1]
        a_, p_, r_ = sh.keep[k], sh.keep[k + 1], sh.keep[k + 2]
        k += 5 if l + 1 == sh.first_rep else 4
        if l > 0:
            bench('%s.A%d' % (name, l), lev['A'], a_, c, c)
        bench('%s.P%d' % (name, l), lev['P'], p_, c, cn)
        bench('%s.R%d' % (name, l), lev['R'], r_, cn, c)
# all-reduce of one double
buf = torch.ones(4, dtype=torch.float64, device='cuda')
t = timeit(lambda: L.npb_dist_allreduce_cb(ctypes.c_void_p(handle), buf.data_ptr(), 1, stream_ptr()))
if rank == 0:
    print('allreduce(1 double): %.1f us' % t, flush=True)
shard.shutdown(); dist.destroy_process_group()
