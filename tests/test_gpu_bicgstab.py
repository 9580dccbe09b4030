"""BiCGStab and the block-Jacobi preconditioner (csrc/bicgstab.cu) -- the Krylov / preconditioner
pair BASELINE.json's north_star names for the full-diagonal Jacobians -- against scipy's SuperLU:
a non-symmetric convection-diffusion operator, block systems in both unknown layouts, and one
implicit step of the reference's test/euler_kec.py residual at a moderate time step."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def npd():
    import numpad_b200
    from numpad_b200 import _dev
    _dev.device()
    return numpad_b200


def _dev_csr(a):
    from numpad_b200._dev import DevCsr
    a = sp.csr_matrix(a)
    a.sort_indices()
    return DevCsr.from_scipy(a)


def _convdiff(N, c=0.4):
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    C = sp.diags([-1.0, 1.0], [-1, 1], shape=(N, N))
    return (sp.kron(sp.identity(N), T) + sp.kron(T, sp.identity(N)) + c * sp.kron(sp.identity(N), C)
            + 0.3 * sp.identity(N * N)).tocsr()


@pytest.mark.parametrize('pc', ['none', 'jacobi'])
def test_bicgstab_convection_diffusion(npd, pc):
    import torch
    from numpad_b200 import linsolve, precond
    a = _convdiff(60)
    rng = np.random.RandomState(0)
    b = rng.standard_normal(a.shape[0])
    A = _dev_csr(a)
    M = precond.JacobiPreconditioner(A) if pc == 'jacobi' else None
    x, info = linsolve.bicgstab(A, torch.from_numpy(b).cuda(), precond=M, rtol=1e-11, maxiter=2000)
    assert info['status'] == 0, info
    ref = spla.spsolve(a.tocsc(), b)
    assert np.abs(x.cpu().numpy() - ref).max() <= 1e-8 * np.abs(ref).max()
    assert np.linalg.norm(b - a @ x.cpu().numpy()) <= 1.01e-11 * np.linalg.norm(b)
    # the same system through solve() with the option
    x2 = linsolve.solve(A, torch.from_numpy(b).cuda(), precond='jacobi', krylov='bicgstab', rtol=1e-11, maxiter=2000)
    assert linsolve.last_info['method'] == 'bicgstab+jacobi'
    assert np.abs(x2.cpu().numpy() - ref).max() <= 1e-8 * np.abs(ref).max()


@pytest.mark.parametrize('layout', ['strided', 'interleaved'])
@pytest.mark.parametrize('bs', [2, 4, 7])
def test_block_jacobi_blocks_and_solve(npd, layout, bs):
    """block-diagonal inverse against numpy, and a block system with strong intra-block and weak
    inter-block coupling, where block Jacobi makes BiCGStab / FGMRES converge in a few steps"""
    import torch
    from numpad_b200 import linsolve, precond
    rng = np.random.RandomState(bs)
    nblk = 900
    n = nblk * bs
    blocks = rng.standard_normal((nblk, bs, bs)) + 3.0 * np.eye(bs)
    row = (lambda c, k: c + k * nblk) if layout == 'strided' else (lambda c, k: c * bs + k)
    ii, jj, vv = [], [], []
    cidx = np.arange(nblk)
    for k in range(bs):
        for j in range(bs):
            ii.append(row(cidx, k)); jj.append(row(cidx, j)); vv.append(blocks[:, k, j])
    a = sp.csr_matrix((np.concatenate(vv), (np.concatenate(ii), np.concatenate(jj))), shape=(n, n))
    a = a + 0.05 * sp.random(n, n, density=4.0 / n, random_state=rng, format='csr')
    a = sp.csr_matrix(a); a.sort_indices()
    A = _dev_csr(a)
    M = precond.BlockJacobiPreconditioner(A, bs, layout)
    assert M.singular_blocks == 0
    # apply == blockdiag(A)^-1 (the diagonal blocks of a, including what the random part added)
    x = rng.standard_normal(n)
    y = torch.empty(n, dtype=torch.float64, device='cuda')
    from numpad_b200 import _lib
    _lib.lib().cdll.npb_block_jacobi_apply_cb(M.c_ctx, torch.from_numpy(x).cuda().data_ptr(), y.data_ptr(), None)
    dense = a.toarray() if n <= 8000 else None
    want = np.empty(n)
    for c in range(nblk):
        rows = np.array([row(c, k) for k in range(bs)])
        blk = dense[np.ix_(rows, rows)] if dense is not None else a[rows][:, rows].toarray()
        want[rows] = np.linalg.solve(blk, x[rows])
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-10, atol=1e-12)
    b = rng.standard_normal(n)
    ref = spla.spsolve(a.tocsc(), b)
    for krylov in ('bicgstab', 'fgmres'):
        xs = linsolve.solve(A, torch.from_numpy(b).cuda(), precond='block-jacobi', block=(bs, layout),
                            krylov=krylov, rtol=1e-12)
        assert linsolve.last_info['method'] == krylov + '+block-jacobi'
        assert linsolve.last_info['iters'] <= 40
        assert np.abs(xs.cpu().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()


def test_singular_block_is_reported(npd):
    from numpad_b200 import precond
    a = sp.identity(8, format='csr').tolil()
    a[1, 1] = 0.0; a[1, 5] = 0.0; a[5, 1] = 0.0; a[5, 5] = 0.0       # block c = 1 of the strided 2 x 2 layout
    a[1, 0] = 1.0; a[5, 0] = 1.0
    M = precond.BlockJacobiPreconditioner(_dev_csr(sp.csr_matrix(a)), 2, 'strided')
    assert M.singular_blocks == 1


@pytest.mark.parametrize('viscous', [False, True])
def test_kec_step_bicgstab_block_jacobi(npd, viscous):
    """configs 4 / 5 with the solver pair north_star names: one implicit step of the euler_kec /
    navierstokes residual on the synthetic duct at a time step where the cell blocks dominate
    (dt = 5e-5, ~12 iterations; the scripts' own dt = 0.1 needs the direct solver, tests/test_gpu_btlu.py) --
    du against SuperLU on the oracle's Jacobian"""
    import torch
    from numpad_b200 import linsolve
    from numpad_b200.workloads.kec import KecProblem
    from oracle import numpad_oracle as orc
    Ni, Nj = 32, 24
    Pg, Po = KecProblem(npd, Ni, Nj, viscous=viscous), KecProblem(orc, Ni, Nj, viscous=viscous)
    w0 = Pg.fixed_state()
    dt = 5e-5
    wg = npd.array(w0.copy())
    rg = Pg.residual(wg, npd.array(w0.copy()), dt)
    Jg = npd.diff_dev(rg, wg)
    wo = orc.array(w0.copy())
    ro = Po.residual(wo, orc.array(w0.copy()), dt)
    Jo = orc.canonical(ro.diff(wo)).tocsc()
    ref = spla.spsolve(Jo, np.ravel(ro._value))
    du = linsolve.solve(Jg, rg._dev.reshape(-1).contiguous(), precond='block-jacobi', block=(4, 'strided'),
                        krylov='bicgstab', rtol=1e-12, maxiter=2000)
    assert linsolve.last_info['method'] == 'bicgstab+block-jacobi', linsolve.last_info
    assert linsolve.last_info['status'] == 0, linsolve.last_info
    assert np.abs(du.cpu().numpy() - ref).max() <= 1e-8 * np.abs(ref).max()
