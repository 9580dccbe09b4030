"""Parity at the benchmarked sizes (VERDICT r1, weak point 1): the Jacobian against the CPU
oracle at 512^2 and 1024^2 (bit-exact pattern, values to 1e-12), and the Newton step du of
the benchmark configuration at 1024^2 / 2048^2 checked by an INDEPENDENT scipy residual
||J du - r|| / ||r|| on the host (the solver's own report is not trusted), at the product's
default tolerance -- the one bench.py times.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import assert_csr_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def npd():
    import numpad_b200
    from numpad_b200 import _dev
    _dev.device()
    return numpad_b200


@pytest.mark.parametrize('N', [512, 1024])
def test_cavity_jacobian_vs_oracle_at_scale(npd, N):
    from oracle import numpad_oracle as orc
    from numpad_b200.workloads.cavity import CavityProblem
    Pg, Po = CavityProblem(npd, N), CavityProblem(orc, N)
    u0 = Pg.fixed_state()
    ug, uo = npd.array(u0.copy()), orc.array(u0.copy())
    rg = Pg.residual(ug, npd.array(u0), 1.0, 0)
    ro = Po.residual(uo, orc.array(u0), 1.0, 0)
    np.testing.assert_allclose(rg._value, ro._value, rtol=1e-13, atol=1e-12)
    Jo = ro.diff(uo)                       # auto -> adjoint, the mode Newton uses
    assert Jo.nnz == 26 * N * N - 42 * N + 11
    Jg = rg.diff(ug)
    assert_csr_parity(Jg, Jo, rtol=1e-12, raw_nnz=Jo.nnz, what='cavity N=%d' % N)


@pytest.mark.parametrize('N', [1024, 2048])
def test_newton_step_residual_checked_on_host(npd, N):
    """one Newton iteration through the API exactly as bench.py makes it; du = u - u_new is
    then checked against J and r rebuilt on the host from the device Jacobian"""
    from numpad_b200 import linsolve, precond
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(npd, N)
    u0 = P.fixed_state()
    precond.forget()
    sol = npd.solve_newton_with_dt(P.residual, npd.adarray(u0.copy()), (npd.adarray(u0.copy()), 1.0, 0),
                                   {}, np.inf, 2, 0.0, 1.0, False)
    assert sol._n_Newton == 2
    assert linsolve.last_info['method'] == 'fgmres+lsc'
    du = u0 - np.asarray(sol._value)
    u = npd.adarray(u0.copy())
    res = P.residual(u, npd.adarray(u0.copy()), 1.0, 0)
    J = res.diff(u)                                  # scipy csr on the host
    assert J.nnz == 26 * N * N - 42 * N + 11
    r = np.asarray(res._value).ravel()
    relres = np.linalg.norm(J @ du - r) / np.linalg.norm(r)
    # u - (u - du) loses ~1e-16 |u| / |du| of du; the solve itself is held to 1e-12
    assert relres <= 5e-12, relres
