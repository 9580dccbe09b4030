"""The reference's own distributed tests (admpi.py:248-327) and its 2-D Poisson
demo (admpisolve.py:462-533, here on a strip decomposition) on torch.distributed:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29512 tools/admpi_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import numpad_b200 as npd
from numpad_b200.admpi import COMM_WORLD, diff_mpi
from numpad_b200.admpisolve import MpiJacobian, _lgmres, _norm2, _dot

rank, size = COMM_WORLD.Get_rank(), COMM_WORLD.Get_size()


def maxabs(m):
    m = sp.csr_matrix(m)
    return np.abs(m.data).max() if m.nnz else 0.0


# ---- admpi.py:249-274 testSendRecv ---- #
if size > 1:
    N = 1000
    u = npd.ones(N)
    if rank == 0:
        COMM_WORLD.Recv(u, 1)
    elif rank == 1:
        COMM_WORLD.Send(u, 0)
    f = u * u
    d = diff_mpi(f, u, mode='tangent')
    nz = 1 if rank == 0 else rank
    assert maxabs(d[nz] - 2 * sp.eye(N, N)) < 1e-14, 'send/recv tangent'
    for r in d:
        if r != nz and not isinstance(d[r], int):
            assert maxabs(d[r]) == 0
    print('rank %d: testSendRecv ok' % rank, flush=True)

# ---- admpi.py:276-327 testPoisson1DResidual ---- #
N = 2000
u = npd.array(np.random.RandomState(rank).random(N))
dx = 1. / (N * size + 1)
u_right, u_left = npd.zeros(1), npd.zeros(1)
if rank > 0:
    COMM_WORLD.Send(u[:1], rank - 1)
if rank < size - 1:
    COMM_WORLD.Recv(u_right, rank + 1)
if rank < size - 1:
    COMM_WORLD.Send(u[-1:], rank + 1)
if rank > 0:
    COMM_WORLD.Recv(u_left, rank - 1)
u_ext = npd.hstack([u_left, u, u_right])
f = (u_ext[2:] + u_ext[:-2] - 2 * u_ext[1:-1]) / dx ** 2
d = diff_mpi(f, u, 'tangent')
lapl = -2 * sp.eye(N, N) + sp.dia_matrix((np.ones(N), 1), (N, N)) + sp.dia_matrix((np.ones(N), -1), (N, N))
assert maxabs(d[rank] - lapl / dx ** 2) < 1e-6, 'diagonal block'
lapl_l = sp.csr_matrix(([1.], ([0], [N - 1])), shape=(N, N))
if rank > 0:
    assert maxabs(d[rank - 1] - lapl_l / dx ** 2) < 1e-6, 'lower block'
if rank < size - 1:
    assert maxabs(d[rank + 1] - lapl_l.T / dx ** 2) < 1e-6, 'upper block'
for r in range(size):
    if abs(r - rank) > 1 and r in d:
        assert isinstance(d[r], int) and d[r] == 0
print('rank %d: testPoisson1DResidual ok (blocks: %s)' % (rank, sorted(d)), flush=True)

# ---- admpisolve.py:462-533: 2-D Poisson, strips of M rows, LGMRES + block-Jacobi ---- #
M_, N_ = 40, 60
uu = npd.zeros((M_, N_))
ff = npd.ones((M_, N_))
dxx, dyy = 1. / (N_ + 1), 1. / (M_ * size + 1)


def residual(u):
    u_up, u_down = npd.zeros(N_), npd.zeros(N_)
    if rank > 0:
        COMM_WORLD.Send(u[0, :], rank - 1)
    if rank < size - 1:
        COMM_WORLD.Recv(u_down, rank + 1)
    if rank < size - 1:
        COMM_WORLD.Send(u[-1, :], rank + 1)
    if rank > 0:
        COMM_WORLD.Recv(u_up, rank - 1)
    zc = npd.zeros((M_ + 2, 1))
    u_ext = npd.vstack([u_up[np.newaxis, :], u, u_down[np.newaxis, :]])
    u_ext = npd.hstack([zc, u_ext, zc])
    return (u_ext[1:-1, 2:] + u_ext[1:-1, :-2] - 2 * u_ext[1:-1, 1:-1]) / dxx ** 2 \
        + (u_ext[2:, 1:-1] + u_ext[:-2, 1:-1] - 2 * u_ext[1:-1, 1:-1]) / dyy ** 2 - ff


r = residual(uu)
from numpad_b200.admpi import diff_mpi_dev
J = MpiJacobian(diff_mpi_dev(r, uu, 'tangent'), 'CSR')
x, info = _lgmres(J.matvec, r._dev.reshape(-1), x0=uu._dev.reshape(-1), tol=1e-10, maxiter=1000,
                  M=J.approx_solve, inner_m=30, outer_k=3)
res = r._dev.reshape(-1) - J.matvec(x)
relres = _norm2(res) / _norm2(r._dev.reshape(-1))
# global reference on the host: the same 5-point operator assembled directly
Mg = M_ * size
T1 = sp.diags([1, -2, 1], [-1, 0, 1], shape=(N_, N_)) / dxx ** 2
T2 = sp.diags([1, -2, 1], [-1, 0, 1], shape=(Mg, Mg)) / dyy ** 2
A = sp.kron(sp.identity(Mg), T1) + sp.kron(T2, sp.identity(N_))
import scipy.sparse.linalg as spla
xg = spla.spsolve(A.tocsc(), -np.ones(Mg * N_))
mine = xg.reshape(Mg, N_)[rank * M_:(rank + 1) * M_].ravel()
err = np.abs(x.cpu().numpy() - mine).max() / np.abs(xg).max()
assert info == 0 and relres < 1e-9 and err < 1e-8, (info, relres, err)
print('rank %d: MpiJacobian + _lgmres ok: relres %.1e, |x - x_global|/|x| = %.1e, <x,x> = %.6e'
      % (rank, relres, err, _dot(x, x)), flush=True)
if dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
