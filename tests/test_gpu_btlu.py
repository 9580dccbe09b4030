"""The block-tridiagonal direct solver (numpad_b200/blocktri.py, csrc/blocktri.cu): kernels
against numpy, solves against scipy's SuperLU on the systems it exists for -- the
convection-dominated Newton systems of the reference's cavity march from rest
(test/cavity.py:97-114), which defeat the multigrid-preconditioned Krylov solver, and the
Euler / Navier-Stokes Jacobians at the scripts' time step."""
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


def test_window_kernels_vs_numpy(npd):
    import torch
    from numpad_b200._dev import _call
    rng = np.random.RandomState(0)
    n = 700
    a = sp.random(n, n, density=0.02, random_state=rng, format='csr') + sp.identity(n)
    a = sp.csr_matrix(a)
    a.sort_indices()
    A = _dev_csr(a)
    dense = a.toarray()
    r0, r1, c0, c1 = 100, 433, 250, 690
    out = torch.zeros((r1 - r0, c1 - c0), dtype=torch.float64, device='cuda')
    _call('npb_bt_window_to_dense', A.indptr, A.indices, A.data, r0, r1, c0, c1, 2.0, out, c1 - c0)
    np.testing.assert_array_equal(out.cpu().numpy(), 2.0 * dense[r0:r1, c0:c1])
    B = torch.from_numpy(rng.standard_normal((c1 - c0, 777))).cuda()
    C = torch.from_numpy(rng.standard_normal((r1 - r0, 777))).cuda()
    want = 0.5 * C.cpu().numpy() - 3.0 * dense[r0:r1, c0:c1] @ B.cpu().numpy()
    _call('npb_bt_window_spmm', A.indptr, A.indices, A.data, r0, r1, c0, c1, B, 777, 777, -3.0, 0.5, C, 777)
    np.testing.assert_allclose(C.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
    x = torch.from_numpy(rng.standard_normal(c1 - c0)).cuda()
    y = torch.from_numpy(rng.standard_normal(r1 - r0)).cuda()
    want = 2.0 * y.cpu().numpy() - dense[r0:r1, c0:c1] @ x.cpu().numpy()
    y0 = y.clone()
    _call('npb_bt_window_spmv', A.indptr, A.indices, A.data, r0, r1, c0, c1, x, -1.0, 2.0, None, y)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
    y2 = torch.empty_like(y)
    _call('npb_bt_window_spmv', A.indptr, A.indices, A.data, r0, r1, c0, c1, x, -1.0, 2.0, y0, y2)
    np.testing.assert_allclose(y2.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
    from numpad_b200 import _lib
    for m in (1, 63, 64, 257, 1000):
        M = torch.from_numpy(rng.standard_normal((m, m))).cuda()
        v = torch.from_numpy(rng.standard_normal(m)).cuda()
        y0 = torch.from_numpy(rng.standard_normal(m)).cuda()
        scratch = torch.empty(int(_lib.lib().cdll.npb_bt_gemv_scratch_doubles(m)), dtype=torch.float64, device='cuda')
        for trans in (0, 1):
            y = torch.empty(m, dtype=torch.float64, device='cuda')
            _call('npb_bt_gemv', m, M, m, v, -1.5, 0.25, y0, y, trans, scratch)
            Mh = M.cpu().numpy().T if trans else M.cpu().numpy()
            np.testing.assert_allclose(y.cpu().numpy(), 0.25 * y0.cpu().numpy() - 1.5 * Mh @ v.cpu().numpy(),
                                       rtol=1e-11, atol=1e-11)


def _march_systems(N, steps):
    """every Newton linear system (J, rhs, du) of the first `steps` time steps of the
    reference's driver loop, from the oracle"""
    from oracle import numpad_oracle as orc
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(orc, N)
    out = []
    orc.newton_trace = lambda it, u, J, rhs, du: out.append((sp.csr_matrix(J), rhs, du))
    try:
        q, force, dt = orc.zeros(P.n), orc.zeros(P.n), 1.0
        for _ in range(steps):
            q = orc.solve(P.residual, q, args=(q, dt, force), rel_tol=0.05, abs_tol=1e-8, verbose=False)
            if q._n_Newton < 4:
                dt *= 2
            q.obliviate()
    finally:
        orc.newton_trace = None
    return out


def test_btlu_on_convection_dominated_march_systems(npd):
    import torch
    from numpad_b200.blocktri import BlockTriLU
    systems = _march_systems(48, 2)
    assert len(systems) >= 6
    for k, (J, rhs, du) in enumerate(systems[1::2]):
        A = _dev_csr(J)
        M = BlockTriLU(A, twisted=(k != 0))                       # one-chain and twisted forms
        assert (M.M == M.K - 1) == (k == 0)
        b = torch.from_numpy(rhs).cuda()
        x = M.solve(b)
        x = x + M.solve(b - A.matvec(x))          # one refinement step
        scale = np.abs(du).max()
        assert np.abs(x.cpu().numpy() - du).max() <= 1e-9 * scale
        xt = M.solve(b, transpose=True)
        xt = xt + M.solve(b - A.transpose().materialized().matvec(xt), transpose=True)
        want = spla.spsolve(J.T.tocsc(), rhs)
        assert np.abs(xt.cpu().numpy() - want).max() <= 1e-9 * np.abs(want).max()


def test_linsolve_falls_back_to_btlu(npd):
    """the path the march takes: multigrid-preconditioned FGMRES fails on the convection-
    dominated Jacobian, the block-tridiagonal solver takes over and stays for what follows"""
    import torch
    from numpad_b200 import linsolve, precond
    systems = _march_systems(64, 1)
    assert len(systems) >= 3
    J, rhs, du = systems[-2]
    A = _dev_csr(J)
    old = precond.DIRECT_WORK
    precond.DIRECT_WORK = 0.0            # (N = 64 would otherwise take the banded LU)
    linsolve.reset_hints()
    try:
        x = linsolve.solve(A, torch.from_numpy(rhs).cuda())
        assert linsolve.last_info['method'].startswith('btlu'), linsolve.last_info
        assert linsolve.last_info['rnorm'] <= 1e-12 * linsolve.last_info['bnorm']
        assert np.abs(x.cpu().numpy() - du).max() <= 1e-9 * np.abs(du).max()
        # the next system of the sequence: straight to the robust solver, old factors reused
        J2, rhs2, du2 = systems[-1]
        x2 = linsolve.solve(_dev_csr(J2), torch.from_numpy(rhs2).cuda())
        assert linsolve.last_info['method'].startswith('btlu')
        assert np.abs(x2.cpu().numpy() - du2).max() <= 1e-9 * np.abs(du2).max()
    finally:
        precond.DIRECT_WORK = old
        linsolve.reset_hints()


@pytest.mark.parametrize('viscous', [False, True])
def test_kec_newton_step_vs_spsolve(npd, viscous):
    """configs 4 / 5 at a parity-test size: one implicit step of euler_kec / navierstokes at the
    scripts' dt; the linear solve must agree with SuperLU on the oracle's Jacobian"""
    import torch
    from oracle import numpad_oracle as orc
    from numpad_b200 import linsolve, precond
    from numpad_b200.workloads.kec import KecProblem
    Ni = Nj = 64
    Pg, Po = KecProblem(npd, Ni, Nj, viscous=viscous), KecProblem(orc, Ni, Nj, viscous=viscous)
    dt = 1.0 / Nj if viscous else 0.1
    w0 = Pg.fixed_state()
    wg, wo = npd.array(w0.copy()), orc.array(w0.copy())
    rg = Pg.residual(wg, npd.array(w0), dt)
    ro = Po.residual(wo, orc.array(w0), dt)
    Jo = ro.diff(wo).tocsr()
    want = spla.spsolve(Jo.tocsc(), np.ravel(ro._value))
    J = npd.diff_dev(rg, wg)
    old = precond.DIRECT_WORK
    precond.DIRECT_WORK = 0.0
    linsolve.reset_hints()
    try:
        x = linsolve.solve(J, rg._dev.reshape(-1))
    finally:
        precond.DIRECT_WORK = old
        linsolve.reset_hints()
    assert linsolve.last_info['rnorm'] <= 1e-12 * linsolve.last_info['bnorm']
    Jh = J.materialized().toscipy()
    r = np.ravel(rg._value)
    xh = x.cpu().numpy()
    assert np.linalg.norm(Jh @ xh - r) <= 1e-11 * np.linalg.norm(r)
    assert np.abs(xh - want).max() <= 1e-6 * np.abs(want).max(), linsolve.last_info


def test_cavity_march_from_rest_256(npd, golden):
    """config 2(i) at a size the oracle can still march on the CPU: the reference's driver loop
    (test/cavity.py:97-114) from rest at 256^2, six time steps.  Fixture: the oracle's march
    (tools/proto_states.py --grid 256 --steps 6; Newton counts, norms and every 97th entry of
    the accepted states).  The convection-dominated Jacobians of every step defeat the
    multigrid-preconditioned solver and go through the block-tridiagonal one."""
    from numpad_b200 import linsolve
    from numpad_b200.workloads.cavity import CavityProblem
    z = golden.cavity_march_256
    P = CavityProblem(npd, 256)
    q, force, dt = npd.zeros(P.n), npd.zeros(P.n), 1.0
    linsolve.reset_hints()
    methods = set()
    try:
        for k in range(len(z['n_newton'])):
            assert dt == float(z['dt'][k])
            q = npd.solve(P.residual, q, args=(q, dt, force), rel_tol=0.05, abs_tol=1e-8, verbose=False)
            methods.add(linsolve.last_info['method'].split(' ')[0])
            assert q._n_Newton == int(z['n_newton'][k]), (k, q._n_Newton)
            v = np.asarray(q._value).ravel()
            assert abs(np.linalg.norm(v) - z['norm'][k]) <= 1e-10 * z['norm'][k]
            assert np.abs(v[::int(z['stride'])] - z['sample'][k]).max() <= 1e-10 * np.abs(z['sample'][k]).max()
            if q._n_Newton < 4:
                dt *= 2
            q.obliviate()
    finally:
        linsolve.reset_hints()
    assert 'btlu' in methods
