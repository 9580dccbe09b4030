"""Parity of the CUDA engine (through the numpad API -> C-ABI) with

  * the golden fixtures produced by the unmodified reference, and
  * the CPU oracle on the same seeded inputs at larger sizes.

Bar (BASELINE.json north_star): canonical sparsity pattern bit-exact (raw nnz
included, i.e. scipy's zero-pruning rule reproduced), Jacobian values within
1e-12 relative, Newton states within 1e-10.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_csr, assert_csr_parity
from opcases import CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def npd():
    import numpad_b200
    from numpad_b200 import _dev
    _dev.device()
    return numpad_b200


@pytest.fixture(scope='module')
def orc():
    from oracle import numpad_oracle
    return numpad_oracle


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('mode', ['tangent', 'adjoint'])
def test_opcase_vs_reference_golden(npd, golden, name, mode):
    z = golden.opcases
    f, u = CASES[name](npd, np.random.RandomState(1234))
    np.testing.assert_allclose(f._value, z[name + '.value'], rtol=1e-13, atol=1e-300)
    J = npd.diff(f, u, mode)
    key = '%s.%s' % (name, mode)
    assert_csr_parity(J, load_csr(z, key), rtol=1e-12,
                      raw_nnz=int(z[key + '.raw_nnz']), what=key)
    assert bool(z[key + '.is_dia']) == type(J).__name__.startswith('dia'), key


@pytest.mark.parametrize('N', [6, 16])
def test_cavity_jacobian_vs_reference_golden(npd, golden, N):
    from numpad_b200.workloads.cavity import CavityProblem
    z = golden.cavity_jacobian
    P = CavityProblem(npd, N)
    u0 = P.fixed_state()
    for mode in ('adjoint', 'tangent'):
        u = npd.array(u0.copy())
        res = P.residual(u, npd.array(u0), 1.0, 0)
        key = 'N%d.fixed.%s' % (N, mode)
        assert_csr_parity(res.diff(u, mode), load_csr(z, key), rtol=1e-12,
                          raw_nnz=int(z[key + '.raw_nnz']), what=key)
    np.testing.assert_allclose(res._value, z['N%d.fixed.res' % N], rtol=1e-13, atol=1e-13)
    u = npd.array(np.zeros(P.n))
    res = P.residual(u, npd.array(np.zeros(P.n)), 1.0, 0)
    key = 'N%d.zero.adjoint' % N
    assert_csr_parity(res.diff(u), load_csr(z, key), rtol=1e-12,
                      raw_nnz=int(z[key + '.raw_nnz']), what=key)


@pytest.mark.parametrize('N,state', [(64, 'fixed'), (64, 'zero'), (160, 'fixed')])
def test_cavity_jacobian_vs_oracle(npd, orc, N, state):
    from numpad_b200.workloads.cavity import CavityProblem
    Pg, Po = CavityProblem(npd, N), CavityProblem(orc, N)
    u0 = Pg.fixed_state() if state == 'fixed' else np.zeros(Pg.n)
    modes = ('adjoint', 'tangent') if N <= 64 else ('adjoint',)
    for mode in modes:
        ug, uo = npd.array(u0.copy()), orc.array(u0.copy())
        rg = Pg.residual(ug, npd.array(u0), 1.0, 0)
        ro = Po.residual(uo, orc.array(u0), 1.0, 0)
        np.testing.assert_allclose(rg._value, ro._value, rtol=1e-13, atol=1e-12)
        Jo = ro.diff(uo, mode)
        assert_csr_parity(rg.diff(ug, mode), Jo, rtol=1e-12, raw_nnz=Jo.nnz,
                          what='cavity N=%d %s %s' % (N, state, mode))
    if state == 'fixed':
        assert Jo.nnz == 26 * N * N - 42 * N + 11


def test_tape_size_matches_reference(npd):
    """336 states per cavity residual, independent of N"""
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(npd, 8)
    u0 = P.fixed_state()
    u = npd.array(u0.copy())
    res = P.residual(u, npd.array(u0), 1.0, 0)
    assert res._current_state._state_id - u._initial_state._state_id == 331


def test_poisson_solve_tangent_adjoint(npd, golden):
    """reference adsolve.py:275-356 known answers + fixture"""
    z = golden.poisson_solve
    N = 40
    dx = np.float64(1.) / (N + 1)

    def residual(u, f, dx):
        w = npd.hstack([0, u, 0])
        return (w[2:] + w[:-2] - 2 * w[1:-1]) / dx ** 2 + f
    f = npd.ones(N)
    dxa = npd.array(dx)
    u = npd.solve(residual, npd.ones(N), args=(f, dxa), verbose=False)
    assert u._n_Newton == 2
    x = np.linspace(0, 1, N + 2)[1:-1]
    np.testing.assert_allclose(u._value, 0.5 * x * (1 - x), atol=1e-11)
    np.testing.assert_allclose(u._value, z['u'], rtol=0, atol=1e-11)
    assert_csr_parity(npd.sum(u).diff(f), load_csr(z, 'du_df.adjoint'), rtol=1e-9)
    assert_csr_parity(u.diff(dxa), load_csr(z, 'du_ddx.tangent'), rtol=1e-9)
    assert_csr_parity(u.diff(f, 'tangent'), load_csr(z, 'du_df.tangent'), rtol=1e-9)


def test_hard_solve_continuation(npd):
    """reference adsolve.py _HardSolveTest: Newton from x=100 on sin(2x)+x
    fails, pseudo-time continuation rescues it"""
    def f(x):
        return npd.sin(2 * x) + x
    u = npd.solve(f, npd.array(np.array([100.0])), verbose=False)
    assert np.isfinite(u._n_Newton)
    assert abs(float(u._value[0])) < 1e-6


def cavity_march(xp, N):
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(xp, N)
    q = xp.zeros(P.n)
    force = xp.zeros(P.n)
    t, dt = 0, 1.
    states, newton = [], []
    while True:
        q = xp.solve(P.residual, q, args=(q, dt, force), rel_tol=0.05, abs_tol=1E-8,
                     verbose=False)
        states.append(np.array(q._value))
        newton.append(q._n_Newton)
        if q._n_Newton == 1:
            break
        elif q._n_Newton < 4:
            dt *= 2
        t += dt
        q.obliviate()
    return P, q, force, np.array(states), newton, t, dt


def test_cavity_march_small_vs_reference_golden(npd, golden):
    """the reference's own driver loop (test/cavity.py:97-114) at N=8: every
    linear solve, every accepted state, the dt schedule and the adjoint of
    test/cavity_adjoint.py:6-8"""
    import numpad_b200.adsolve as ads
    z = golden.cavity_march_small
    N = int(z['N'])
    lin = []
    ads.newton_trace = lambda it, u, J, b, du: lin.append(du.cpu().numpy().copy())
    try:
        P, q, force, states, newton, t, dt = cavity_march(npd, N)
    finally:
        ads.newton_trace = None
    assert len(lin) == int(z['n_linear_solves'])
    np.testing.assert_array_equal(newton, z['n_newton'])
    assert t == float(z['final_t']) and dt == float(z['final_dt'])
    scale = np.abs(z['states']).max()
    np.testing.assert_allclose(states, z['states'], rtol=0, atol=1e-10 * scale)
    np.testing.assert_allclose(np.array(lin), z['lin_du'], rtol=0,
                               atol=1e-9 * np.abs(z['lin_du']).max())
    ke = 0.5 * npd.sum(q[N * N:] ** 2)
    adj = ke.diff(force).toarray().ravel()
    np.testing.assert_allclose(adj, z['adjoint'], rtol=0,
                               atol=1e-8 * np.abs(z['adjoint']).max())


@pytest.mark.parametrize('N', [64, 128])
def test_cavity_newton_step_lsc_vs_oracle(npd, orc, N):
    """one Newton linear system of the benchmark state solved by FGMRES + LSC/AMG
    against SuperLU on the oracle's Jacobian"""
    import scipy.sparse.linalg as spla
    from numpad_b200 import linsolve
    from numpad_b200.workloads.cavity import CavityProblem
    Pg, Po = CavityProblem(npd, N), CavityProblem(orc, N)
    u0 = Pg.fixed_state()
    ug = npd.array(u0.copy())
    rg = Pg.residual(ug, npd.array(u0), 1.0, 0)
    J = npd.diff_dev(rg, ug)
    du = linsolve.solve(J, rg._dev.reshape(-1).contiguous(), precond='lsc')
    info = dict(linsolve.last_info)
    uo = orc.array(u0.copy())
    ro = Po.residual(uo, orc.array(u0), 1.0, 0)
    ref = spla.spsolve(orc.canonical(ro.diff(uo)).tocsc(), ro._value)
    print('N=%d lsc: %s' % (N, info))
    assert info['iters'] <= 80
    np.testing.assert_allclose(du.cpu().numpy(), ref, rtol=0, atol=1e-9 * np.abs(ref).max())


def test_cavity_script_config1_vs_reference_golden(npd, golden):
    """BASELINE.json configs[0]: test/cavity.py verbatim (N=50, 29 time steps,
    71 Newton linear solves, dt doubling to 8192) and test/cavity_adjoint.py:6-8,
    against the unmodified reference's results."""
    import numpad_b200.adsolve as ads
    z = golden.cavity_march_script
    N = int(z['N'])
    count = []
    ads.newton_trace = lambda it, u, J, b, du: count.append(it)
    try:
        P, q, force, states, newton, t, dt = cavity_march(npd, N)
    finally:
        ads.newton_trace = None
    assert len(count) == int(z['n_linear_solves']) == 71
    np.testing.assert_array_equal(newton, z['n_newton'])
    assert len(states) == 29 and t == float(z['final_t']) == 16476.0 and dt == 8192.0
    scale = np.abs(z['states']).max()
    np.testing.assert_allclose(states[[0, 1, -1]], z['states'], rtol=0, atol=1e-10 * scale)
    np.testing.assert_allclose(np.linalg.norm(states, axis=1), z['state_norms'], rtol=1e-10)
    assert abs(np.linalg.norm(states[-1]) - 6.015146088185117) < 1e-9
    assert abs(q._res_norm - float(z['final_res_norm'])) < 1e-6 * float(z['final_res_norm']) + 1e-12
    ke = 0.5 * npd.sum(q[N * N:] ** 2)
    assert abs(float(ke._value) - float(z['kinetic_energy'])) < 1e-10 * abs(float(z['kinetic_energy']))
    adj = ke.diff(force).toarray().ravel()
    np.testing.assert_allclose(adj, z['adjoint'], rtol=0, atol=1e-8 * np.abs(z['adjoint']).max())
    assert abs(np.linalg.norm(adj) - 7.079873951678084e+02) < 1e-5


@pytest.mark.parametrize('N,dt', [(96, 0.02), (96, 1.0)])
def test_cavity_adjoint_through_solve_vs_oracle(npd, orc, N, dt):
    """BASELINE.json configs[2] at sizes the oracle finishes in seconds: one
    implicit time step solved by Newton with the Krylov solver (no direct solve
    at these sizes), then the adjoint of the kinetic energy w.r.t. the forcing:
    one TRANSPOSED solve J^T lambda = g^T (reference adsolve.py:49-86).
    dt = 0.02 stays in the diffusion-dominated regime (LSC + multigrid);
    dt = 1 from rest drives the lid flow into the convection-dominated regime
    where the multigrid momentum solve fails and the block-tridiagonal direct solver
    ('btlu') carries the Newton sequence and, through its transposed
    sweeps, the adjoint."""
    from numpad_b200 import linsolve
    from numpad_b200.workloads.cavity import CavityProblem
    out = {}
    for name, xp in (('gpu', npd), ('cpu', orc)):
        P = CavityProblem(xp, N)
        q0 = xp.array(P.fixed_state(amp=1e-3))
        force = xp.zeros(P.n)
        q = xp.solve(P.residual, q0, args=(q0, dt, force), rel_tol=1e-9, abs_tol=1e-9,
                     verbose=False)
        ke = 0.5 * xp.sum(q[N * N:] ** 2)
        adj = ke.diff(force).toarray().ravel()
        out[name] = (np.array(q._value), q._n_Newton, adj)
        if name == 'gpu':
            m = linsolve.last_info['method']
            assert m.startswith('fgmres+lsc') if dt < 1 else m.startswith('btlu'), m
    assert out['gpu'][1] == out['cpu'][1]
    scale = np.abs(out['cpu'][0]).max()
    np.testing.assert_allclose(out['gpu'][0], out['cpu'][0], rtol=0, atol=1e-10 * scale)
    np.testing.assert_allclose(out['gpu'][2], out['cpu'][2], rtol=0,
                               atol=1e-8 * np.abs(out['cpu'][2]).max())


@pytest.mark.parametrize('tag', ['euler', 'ns'])
def test_kec_jacobian_vs_reference_golden(npd, golden, tag):
    """BASELINE.json configs[3]/[4] residuals (test/euler_kec.py,
    test/navierstokes.py: broadcasting products, axis sums, list indexing,
    augmented slice assignment) against the unmodified reference"""
    from numpad_b200.workloads.kec import KecProblem
    z = golden.kec_jacobian
    Ni, Nj = int(z['Ni']), int(z['Nj'])
    P = KecProblem(npd, Ni, Nj, viscous=(tag == 'ns'))
    w0 = P.fixed_state()
    dt = 0.1 if tag == 'euler' else 1. / Nj
    for mode in ('adjoint', 'tangent'):
        u = npd.array(w0.copy())
        res = P.residual(u, npd.array(w0.copy()), dt)
        key = '%s.%s' % (tag, mode)
        assert_csr_parity(res.diff(u, mode), load_csr(z, key), rtol=1e-12,
                          raw_nnz=int(z[key + '.raw_nnz']), what=key)
    scale = np.abs(z[tag + '.res']).max()
    np.testing.assert_allclose(res._value, z[tag + '.res'], rtol=0, atol=1e-12 * scale)
    assert res._current_state._state_id - u._initial_state._state_id == int(z[tag + '.states'])


@pytest.mark.parametrize('tag,Ni,Nj', [('euler', 48, 32), ('ns', 40, 24)])
def test_kec_jacobian_vs_oracle(npd, orc, tag, Ni, Nj):
    from numpad_b200.workloads.kec import KecProblem
    Pg = KecProblem(npd, Ni, Nj, viscous=(tag == 'ns'))
    Po = KecProblem(orc, Ni, Nj, viscous=(tag == 'ns'))
    w0 = Pg.fixed_state()
    dt = 0.1 if tag == 'euler' else 1. / Nj
    ug, uo = npd.array(w0.copy()), orc.array(w0.copy())
    rg = Pg.residual(ug, npd.array(w0.copy()), dt)
    ro = Po.residual(uo, orc.array(w0.copy()), dt)
    Jo = ro.diff(uo)
    # v = 0 in the base flow makes a handful of entries exact cancellations
    assert_csr_parity(rg.diff(ug), Jo, rtol=1e-12, what=tag, noise=1e-15)
    assert (Jo.diagonal() != 0).all()           # full diagonal


def test_admpi_single_rank(npd):
    """admpi / admpisolve surface on one rank (the multi-rank run is
    tools/admpi_check.py under torchrun): diff_mpi == diff, MpiJacobian.matvec,
    block-Jacobi approx_solve and _lgmres with the reference's signature"""
    import torch
    from numpad_b200.admpi import COMM_WORLD, diff_mpi, diff_mpi_dev
    from numpad_b200.admpisolve import MpiJacobian, _lgmres, _dot, _norm2
    assert COMM_WORLD.Get_size() == 1 and COMM_WORLD.Get_rank() == 0
    N = 300
    u = npd.array(np.random.RandomState(0).random(N))
    w = npd.hstack([0, u, 0])
    f = (w[2:] + w[:-2] - 2 * w[1:-1]) * 4.0 - npd.exp(u)
    d = diff_mpi(f, u, 'tangent')
    assert sorted(d) == [0]
    assert_csr_parity(d[0], f.diff(u, 'tangent'), rtol=0.0)
    with pytest.raises(NotImplementedError):
        diff_mpi(f, u, 'adjoint')
    J = MpiJacobian(diff_mpi_dev(f, u, 'tangent'), 'CSR')
    x = torch.randn(N, dtype=torch.float64, device='cuda')
    Jh = d[0]
    np.testing.assert_allclose(J.matvec(x).cpu().numpy(), Jh @ x.cpu().numpy(), rtol=1e-12)
    b = J.matvec(x)
    np.testing.assert_allclose(J.approx_solve(b).cpu().numpy(), x.cpu().numpy(), atol=1e-9)
    sol, info = _lgmres(J.matvec, b, tol=1e-12, M=J.approx_solve)
    assert info == 0
    np.testing.assert_allclose(sol.cpu().numpy(), x.cpu().numpy(), atol=1e-9)
    assert abs(_dot(x, x) - float(x @ x)) < 1e-9 and abs(_norm2(x) - float(x.norm())) < 1e-10


def test_multi_rhs_tangent_through_solve_vs_oracle(npd, orc):
    """d u / d p for 16 parameters through one cavity solve (reference
    adsolve.py:113-147: one LU solve per column of the densified dR/dp).  Here the solver is set
    up once for all 16 right-hand sides (linsolve.solve_multi), operands stay on the device;
    the result must match the oracle's to 1e-8."""
    import time
    from numpad_b200 import linsolve
    from numpad_b200.workloads.cavity import CavityProblem
    N, k = 96, 16
    rng = np.random.RandomState(7)
    n = N * (3 * N - 2)
    B = np.zeros((n, k))
    for j in range(k):                       # 16 localised forcing patterns
        idx = rng.randint(N * N, n, 200)
        B[idx, j] = rng.standard_normal(200)
    out = {}
    for name, xp in (('gpu', npd), ('cpu', orc)):
        P = CavityProblem(xp, N)
        q0 = xp.array(P.fixed_state(amp=1e-3))
        p = xp.array(0.1 * np.ones(k))
        Bx = xp.array(B)

        def residual(q, q_old, dt, p):
            return P.residual(q, q_old, dt, xp.dot(Bx, p))
        t0 = time.perf_counter()
        q = xp.solve(residual, q0, args=(q0, 0.02, p), rel_tol=1e-9, abs_tol=1e-9, verbose=False)
        t1 = time.perf_counter()
        dq = q.diff(p, 'tangent')
        t2 = time.perf_counter()
        out[name] = (np.asarray(dq.todense()), t1 - t0, t2 - t1)
        if name == 'gpu':
            assert 'reused' in linsolve.last_info['method'], linsolve.last_info['method']
    g, c = out['gpu'][0], out['cpu'][0]
    assert g.shape == c.shape == (n, k)
    np.testing.assert_allclose(g, c, rtol=0, atol=1e-8 * np.abs(c).max())
    print('multi-RHS tangent: GPU solve %.2fs, 16-column tangent %.2fs; oracle %.2fs / %.2fs'
          % (out['gpu'][1], out['gpu'][2], out['cpu'][1], out['cpu'][2]))
