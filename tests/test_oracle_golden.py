"""Pin the CPU oracle (oracle/numpad_oracle.py) to the reference.

The fixtures were produced by the unmodified reference (tests/golden/
make_golden.py).  The oracle must reproduce them: patterns bit-exact (raw nnz
included -- the zero-pruning rule of scipy's CSR sum), values to 1e-14.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_csr, assert_csr_parity
from opcases import CASES
from oracle import numpad_oracle as orc
from numpad_b200.workloads.cavity import CavityProblem


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('mode', ['tangent', 'adjoint'])
def test_opcase(golden, name, mode):
    z = golden.opcases
    f, u = CASES[name](orc, np.random.RandomState(1234))
    np.testing.assert_array_equal(f._value, z[name + '.value'])
    J = orc.diff(f, u, mode)
    key = '%s.%s' % (name, mode)
    assert_csr_parity(J, load_csr(z, key), rtol=0.0,
                      raw_nnz=int(z[key + '.raw_nnz']), what=key)
    assert bool(z[key + '.is_dia']) == type(J).__name__.startswith('dia')


@pytest.mark.parametrize('N', [6, 16])
def test_cavity_jacobian(golden, N):
    z = golden.cavity_jacobian
    P = CavityProblem(orc, N)
    u0 = P.fixed_state()
    for mode in ('adjoint', 'tangent'):
        u = orc.array(u0.copy())
        res = P.residual(u, orc.array(u0), 1.0, 0)
        key = 'N%d.fixed.%s' % (N, mode)
        assert_csr_parity(res.diff(u, mode), load_csr(z, key), rtol=0.0,
                          raw_nnz=int(z[key + '.raw_nnz']), what=key)
    np.testing.assert_array_equal(res._value, z['N%d.fixed.res' % N])
    u = orc.array(np.zeros(P.n))
    res = P.residual(u, orc.array(np.zeros(P.n)), 1.0, 0)
    key = 'N%d.zero.adjoint' % N
    assert_csr_parity(res.diff(u), load_csr(z, key), rtol=0.0,
                      raw_nnz=int(z[key + '.raw_nnz']), what=key)


def test_cavity_nnz_formula():
    """nnz(J) = 26 N^2 - 42 N + 11 at a generic state."""
    for N in (6, 16, 24):
        P = CavityProblem(orc, N)
        u0 = P.fixed_state()
        u = orc.array(u0.copy())
        J = P.residual(u, orc.array(u0), 1.0, 0).diff(u)
        assert J.nnz == 26 * N * N - 42 * N + 11
        d = J.diagonal()
        assert (d[:N * N] == 0).sum() == N * N - 1      # saddle point


def cavity_march(xp, N, max_steps=None):
    P = CavityProblem(xp, N)
    q = xp.zeros(P.n)
    force = xp.zeros(P.n)
    t, dt = 0, 1.
    states, newton = [], []
    while True:
        q = xp.solve(P.residual, q, args=(q, dt, force), rel_tol=0.05,
                     abs_tol=1E-8, verbose=False)
        states.append(q._value.copy())
        newton.append(q._n_Newton)
        if q._n_Newton == 1:
            break
        elif q._n_Newton < 4:
            dt *= 2
        t += dt
        if max_steps and len(states) >= max_steps:
            break
        q.obliviate()
    return P, q, force, np.array(states), newton, t, dt


def test_cavity_march_small(golden):
    z = golden.cavity_march_small
    N = int(z['N'])
    lin = []
    orc.newton_trace = lambda it, u, J, b, du: lin.append(du.copy())
    try:
        P, q, force, states, newton, t, dt = cavity_march(orc, N)
    finally:
        orc.newton_trace = None
    assert len(lin) == int(z['n_linear_solves'])
    np.testing.assert_array_equal(newton, z['n_newton'])
    np.testing.assert_allclose(states, z['states'], rtol=0, atol=1e-13)
    np.testing.assert_allclose(np.array(lin), z['lin_du'], rtol=0, atol=1e-12)
    assert t == float(z['final_t']) and dt == float(z['final_dt'])
    ke = 0.5 * orc.sum(q[N * N:] ** 2)
    adj = ke.diff(force).toarray().ravel()
    np.testing.assert_allclose(adj, z['adjoint'], rtol=1e-11,
                               atol=1e-12 * np.abs(z['adjoint']).max())


def test_poisson_solve(golden):
    z = golden.poisson_solve
    N = 40
    dx = np.float64(1.) / (N + 1)

    def residual(u, f, dx):
        w = orc.hstack([0, u, 0])
        return (w[2:] + w[:-2] - 2 * w[1:-1]) / dx ** 2 + f
    f = orc.ones(N)
    dxa = orc.array(dx)
    u = orc.solve(residual, orc.ones(N), args=(f, dxa), verbose=False)
    np.testing.assert_allclose(u._value, z['u'], rtol=0, atol=1e-14)
    x = np.linspace(0, 1, N + 2)[1:-1]
    np.testing.assert_allclose(u._value, 0.5 * x * (1 - x), atol=1e-12)
    assert_csr_parity(orc.sum(u).diff(f), load_csr(z, 'du_df.adjoint'), rtol=1e-13)
    assert_csr_parity(u.diff(dxa), load_csr(z, 'du_ddx.tangent'), rtol=1e-13)
    assert_csr_parity(u.diff(f, 'tangent'), load_csr(z, 'du_df.tangent'), rtol=1e-13)


@pytest.mark.parametrize('tag', ['euler', 'ns'])
def test_kec_jacobian(golden, tag):
    """configs 4/5 residuals (test/euler_kec.py, test/navierstokes.py) on the
    synthetic duct mesh: oracle + restated workload vs the reference, bit-exact"""
    from numpad_b200.workloads.kec import KecProblem
    z = golden.kec_jacobian
    Ni, Nj = int(z['Ni']), int(z['Nj'])
    P = KecProblem(orc, Ni, Nj, viscous=(tag == 'ns'))
    w0 = P.fixed_state()
    np.testing.assert_array_equal(w0, z[tag + '.w0'])
    dt = 0.1 if tag == 'euler' else 1. / Nj
    for mode in ('adjoint', 'tangent'):
        u = orc.array(w0.copy())
        res = P.residual(u, orc.array(w0.copy()), dt)
        key = '%s.%s' % (tag, mode)
        assert_csr_parity(res.diff(u, mode), load_csr(z, key), rtol=0.0,
                          raw_nnz=int(z[key + '.raw_nnz']), what=key)
    np.testing.assert_array_equal(res._value, z[tag + '.res'])
    assert res._current_state.id - u._initial_state.id == int(z[tag + '.states'])
