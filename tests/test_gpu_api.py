"""Drop-in behaviours of the Python surface that scripts rely on:
writable `_value`, `adsolution` with a lazily assembled Jacobian, and `solve()`
surviving a failing linear solve through pseudo-time continuation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def npd():
    import numpad_b200
    from numpad_b200 import _dev
    _dev.device()
    return numpad_b200


def test_value_is_a_writable_mirror(npd):
    a = npd.array(np.arange(6.0))
    v = a._value
    assert v.flags.writeable
    v -= 1.0                                    # reference adsolve.py:198 mutates _value in place
    np.testing.assert_array_equal((a * 2)._value, 2 * (np.arange(6.0) - 1))
    a._value[2] = 10.0                          # item assignment through a fresh handle
    a._value[4:] = [7.0, 8.0]                   # ... and through a view
    np.testing.assert_array_equal(a._dev.cpu().numpy(), [-1, 0, 10, 2, 7, 8])
    np.multiply(a._value, 2.0, out=a._value)    # ufunc with out=
    np.testing.assert_array_equal(a._dev.cpu().numpy(), [-2, 0, 20, 4, 14, 16])
    # arithmetic on the mirror yields plain arrays; reading does not dirty anything
    assert type(a._value * 2) is np.ndarray and type(np.linalg.norm(a._value)) is not type(v)
    # device-side changes invalidate the mirror: the next read sees them
    a[0] = 5.0
    assert a._value[0] == 5.0
    # a stale handle from before the device change is detached: writing it changes nothing
    v[1] = 99.0
    assert a._value[1] == 0.0
    a._value = np.ones(6)
    np.testing.assert_array_equal(a._dev.cpu().numpy(), np.ones(6))


def test_lazy_jacobian_matches_eager(npd):
    from numpad_b200 import adsolve
    N = 30
    dx = 1.0 / (N + 1)

    def residual(u, f):
        w = npd.hstack([0, u, 0])
        return (w[2:] + w[:-2] - 2 * w[1:-1]) / dx ** 2 + f * npd.exp(-u)
    out = {}
    for lazy in (True, False):
        adsolve.lazy_jacobian = lazy
        try:
            f = npd.ones(N)
            u = npd.solve(residual, npd.zeros(N), args=(f,), verbose=False)
            J = npd.sum(u ** 2).diff(f)
            out[lazy] = (np.asarray(u._value).copy(), np.asarray(J.todense()),
                         u._current_state.jacobian.toscipy().toarray())
        finally:
            adsolve.lazy_jacobian = True
    np.testing.assert_array_equal(out[True][0], out[False][0])
    np.testing.assert_array_equal(out[True][2], out[False][2])
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=1e-13, atol=1e-15)


def test_failed_linear_solve_falls_back_to_continuation(npd, monkeypatch):
    """a LinearSolveError inside Newton must not leave solve(); the step counts as
    not converged (`_n_Newton = inf`) and pseudo-time continuation, whose 1/dt shift changes the
    matrix, takes over -- the reference's contract (adsolve.py:225-264)."""
    from numpad_b200 import linsolve, adsolve
    real = linsolve.solve
    calls = {'failed': 0}

    def flaky(A, b, **kw):
        if calls['failed'] < 1:           # the first system (plain Newton, dt = inf) stagnates
            calls['failed'] += 1
            raise linsolve.LinearSolveError('synthetic stagnation')
        return real(A, b, **kw)
    monkeypatch.setattr(adsolve.linsolve, 'solve', flaky)

    def residual(u):
        return u ** 3 + 2 * u - 1.0
    u = npd.solve(residual, npd.zeros(5), verbose=False)
    assert calls['failed'] >= 1
    assert np.isfinite(u._n_Newton)
    r = np.asarray(u._value)
    np.testing.assert_allclose(r ** 3 + 2 * r - 1.0, 0.0, atol=1e-5)


def test_lsc_graph_replay_and_fp32_operators_do_not_change_the_solve(npd):
    """the preconditioner's launch sequence replayed as a CUDA graph gives bit-identical
    iterates to eager launches; storing the preconditioner's large operators in fp32 changes
    the iteration count by at most a few and never the answer (FGMRES is fp64 outside)"""
    import torch
    from numpad_b200 import linsolve, precond, amg, _lib
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(npd, 320)             # pressure block above amg.F32_MIN_ROWS? no: force it
    u0 = P.fixed_state()
    u = npd.array(u0.copy())
    res = P.residual(u, npd.array(u0), 1.0, 0)
    J = npd.diff_dev(res, u).materialized()
    b = res._dev.reshape(-1).contiguous()
    L = _lib.lib().cdll
    out = {}
    old_min = amg.F32_MIN_ROWS
    try:
        for tag, graphs, f32 in (('eager64', 0, 0), ('graph64', 1, 0), ('graph32', 1, 50000)):
            amg.F32_MIN_ROWS = f32
            precond.forget()
            linsolve.reset_hints()
            L.npb_lsc_graphs_enable(graphs)
            x = linsolve.solve(J, b, precond='lsc')
            out[tag] = (x.clone(), linsolve.last_info['iters'],
                        linsolve.last_info['rnorm'] / linsolve.last_info['bnorm'])
    finally:
        amg.F32_MIN_ROWS = old_min
        L.npb_lsc_graphs_enable(1)
        precond.forget()
    assert torch.equal(out['eager64'][0], out['graph64'][0]) and out['eager64'][1] == out['graph64'][1]
    assert out['graph32'][2] <= 1e-12 and abs(out['graph32'][1] - out['eager64'][1]) <= 4
    ref = out['eager64'][0]
    assert float((out['graph32'][0] - ref).norm() / ref.norm()) < 1e-9


def test_assembly_plan_replay_is_bit_identical_and_self_checking(npd):
    """trace once / replay: the second and later Jacobians of a tape reuse the
    recorded sparsity patterns (no symbolic kernels, one host sync per assembly) and must equal
    the plan-free assembly bit for bit; a state whose Jacobian has exact zeros (the script's
    first step from rest) must fall back, and a different residual of the same size must not
    be served from the stale plan"""
    import torch
    from numpad_b200 import _dev
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(npd, 48)

    def jac(state, replay):
        _dev.REPLAY = replay
        u = npd.array(state.copy())
        res = P.residual(u, npd.array(state.copy()), 1.0, 0)
        return npd.diff_dev(res, u).materialized()
    same = lambda a, b: (a.nnz == b.nnz and torch.equal(a.indptr, b.indptr) and
                         torch.equal(a.indices, b.indices) and torch.equal(a.data, b.data))
    try:
        _dev.plans.clear()
        for k in _dev.plan_stats:
            _dev.plan_stats[k] = 0
        s1, s2 = P.fixed_state(seed=1), P.fixed_state(seed=2)
        ref1, ref2, ref0 = jac(s1, False), jac(s2, False), jac(np.zeros(P.n), False)
        a = jac(s1, True)                       # records
        assert _dev.plan_stats['recorded'] == 1 and same(a, ref1)
        b = jac(s2, True)                       # replays
        assert _dev.plan_stats['replayed'] == 1 and same(b, ref2)
        z = jac(np.zeros(P.n), True)            # exact zeros -> mismatch -> plan-free assembly
        assert _dev.plan_stats['mismatch'] == 1 and same(z, ref0)
        assert z.nnz < ref1.nnz
        c = jac(s1, True)                       # records again (the zero state's run is not kept)
        d = jac(s2, True)
        assert same(c, ref1) and same(d, ref2) and _dev.plan_stats['replayed'] >= 2
        # another function of the same sizes: different index maps -> guarded
        u = npd.array(s1.copy())
        other = npd.hstack([u[1:] * u[:-1], u[:1]]) + npd.roll(u, 3)
        J1 = npd.diff_dev(other, u).materialized()
        _dev.REPLAY = False
        u = npd.array(s1.copy())
        other = npd.hstack([u[1:] * u[:-1], u[:1]]) + npd.roll(u, 3)
        assert same(J1, npd.diff_dev(other, u).materialized())
    finally:
        _dev.REPLAY = True
        _dev.plans.clear()
