"""The rest of the package surface (reference __init__.py:1-7): `random`,
`linalg.solve`, `sparse.csr_matrix` / `spsolve`, `interp`, `visual.dot`, `collect` -- the
reference's own unit tests restated (adtools.py:94-127, adlinalg.py:28-44, adsparse.py:98-149)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def npd():
    import numpad_b200
    from numpad_b200 import _dev
    _dev.device()
    return numpad_b200


def test_interp_matches_knots_and_derivatives(npd):
    np.random.seed(3)
    N = 11
    x0 = npd.random.random(N)
    x0.sort()
    y0 = npd.random.random(N)
    for kind in ('linear', 'cubic'):
        y = npd.interp(x0, y0, kind)
        np.testing.assert_allclose(npd.value(y(x0.copy())), npd.value(y0), atol=1e-12)
    y = npd.interp(x0, y0, 'cubic')
    x = x0.copy()
    slopes = np.asarray(npd.value(y.y0[:, 1]))
    np.testing.assert_allclose(npd.value(y.derivative(x)), slopes, atol=1e-9)
    np.testing.assert_allclose(np.diag(y(x).diff(x).todense()), slopes, atol=1e-9)
    # linear: extrapolates the last interval
    x0 = npd.array(np.arange(N, dtype=float))
    y0 = npd.array(np.arange(N, dtype=float))
    y0[-1] = 0
    y = npd.interp(x0, y0, 'linear')
    x = npd.linspace(-1, N - 2, 1000)
    np.testing.assert_allclose(npd.value(y(x)), npd.value(x), atol=1e-12)


def test_linalg_solve_derivative_of_inverse(npd):
    """d(A + s I)^-1 / ds = -A^-1 A^-1   (reference adlinalg.py:29-44)"""
    np.random.seed(5)
    N = 3
    s = npd.array(1.0)
    A = npd.random.random([N, N]) + s * npd.eye(N)
    Ainv = npd.linalg.solve(A, npd.eye(N))
    np.testing.assert_allclose(npd.value(Ainv), np.linalg.inv(npd.value(A)), atol=1e-12)
    d = np.array(Ainv.diff(s).todense()).reshape([N, N])
    np.testing.assert_allclose(d, -np.dot(npd.value(Ainv), npd.value(Ainv)), atol=1e-10)


def test_sparse_matrix_and_spsolve_adjoint(npd):
    """-(a u')' = b through the residual route and through csr_matrix + spsolve: same
    sensitivities d(sum u)/da, both equal to finite differences (reference adsparse.py:98-149)"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    N = 40
    dx = 1. / N
    a, b = npd.ones(N), npd.ones(N - 1)

    def resid(u):
        u = npd.hstack([0, u, 0])
        adu = a * (u[1:] - u[:-1]) / dx
        return (adu[1:] - adu[:-1]) / dx - b
    u = npd.solve(resid, npd.ones(N - 1), verbose=False)
    adj = np.array(npd.sum(u).diff(a).todense()).ravel()

    def tridiag(a):
        lower = upper = a[1:-1] / dx ** 2
        diag = -(a[:-1] + a[1:]) / dx ** 2
        i = np.hstack([np.arange(1, N - 1), np.arange(N - 2), np.arange(N - 1)])
        j = np.hstack([np.arange(N - 2), np.arange(1, N - 1), np.arange(N - 1)])
        return npd.sparse.csr_matrix([npd.hstack([lower, upper, diag]), (i, j)])
    u1 = npd.sparse.spsolve(tridiag(a), b)
    np.testing.assert_allclose(npd.value(u1), npd.value(u), atol=1e-9)
    adj1 = np.array(npd.sum(u1).diff(a).todense()).ravel()
    np.testing.assert_allclose(adj1, adj, atol=1e-8)
    fd = np.zeros(N)
    base = spla.spsolve(tridiag(npd.ones(N))._value.tocsc(), np.ones(N - 1))
    for k in range(N):
        ap = np.ones(N)
        ap[k] += 1e-6
        fd[k] = (spla.spsolve(tridiag(npd.array(ap))._value.tocsc(), np.ones(N - 1)) - base).sum() / 1e-6
    np.testing.assert_allclose(adj1, fd, atol=1e-4)
    # matrix times a stack of vectors
    B = npd.transpose([[b, b], [b, b]])
    out = tridiag(a) * B
    assert out.shape == (N - 1, 2, 2)
    np.testing.assert_allclose(npd.value(out)[:, 1, 0], tridiag(a)._value @ np.ones(N - 1), atol=1e-12)


def test_visual_dot_and_collect(npd):
    a = npd.ones(4)
    b = npd.exp(a) * 2 + a
    g = npd.visual.dot(b)
    assert g.startswith('digraph G{\n') and g.endswith('}\n')
    assert '[label="exp"' in g and '->' in g
    from numpad_b200.adgarbagecollect import collect
    c = npd.exp(npd.ones(3) * 3.0)          # the intermediate (ones * 3) host is gone
    collect(c._current_state)
    assert c._current_state.other is None or c._current_state.other.host() is not None
