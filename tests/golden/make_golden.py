"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

The reference is imported read-only through the numpy-2 shim of SURVEY.md
Appendix C (a no-op ``np.set_numeric_ops``, ``adarray.__array_ufunc__ = None``)
from a temporary directory holding a symlink ``numpad -> /root/reference``;
nothing of the reference is copied into this repository -- only the numbers it
produces.  The GPU box has no /root/reference; tests there read the fixtures.
"""
import io
import os
import sys
import tempfile
import time
import contextlib

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'


def import_reference():
    tmp = tempfile.mkdtemp(prefix='numpad_ref_')
    os.symlink(REF, os.path.join(tmp, 'numpad'))
    ops = {k: getattr(np, k) for k in
           ('add', 'subtract', 'multiply', 'true_divide', 'divide')}
    np.set_numeric_ops = lambda **kw: dict(ops)
    time.clock = time.perf_counter
    sys.path.insert(0, tmp)
    import warnings
    warnings.filterwarnings('ignore')
    import numpad
    numpad.adarray.__array_ufunc__ = None
    return numpad


def pack(prefix, m, out):
    """raw CSR as the reference returned it + its canonical form"""
    if isinstance(m, (int, float)):
        raise TypeError(m)
    raw = sp.csr_matrix(m)
    out[prefix + '.raw_nnz'] = np.int64(raw.nnz)
    out[prefix + '.is_dia'] = np.bool_(sp.isspmatrix_dia(m) or
                                      type(m).__name__.startswith('dia'))
    can = sp.csr_matrix(m, copy=True)
    can.sum_duplicates()
    can.sort_indices()
    out[prefix + '.shape'] = np.array(can.shape, np.int64)
    out[prefix + '.indptr'] = can.indptr.astype(np.int64)
    out[prefix + '.indices'] = can.indices.astype(np.int32)
    out[prefix + '.data'] = can.data.astype(np.float64)


def gen_opcases(numpad):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from opcases import CASES
    out = {}
    for name, fn in CASES.items():
        for mode in ('tangent', 'adjoint'):
            f, u = fn(numpad, np.random.RandomState(1234))
            J = numpad.diff(f, u, mode)
            pack('%s.%s' % (name, mode), J, out)
            out['%s.value' % name] = np.asarray(f._value, np.float64)
    np.savez_compressed(os.path.join(HERE, 'opcases.npz'), **out)
    print('opcases:', len(CASES), 'cases')


def cavity_namespace(numpad, N):
    src = open(os.path.join(REF, 'test', 'cavity.py')).read()
    src = src.split('# ---------------------- time integration')[0]
    g = {}
    exec(src, g)
    g.update(N=N, dx=1. / N, dy=1. / N, Re=10000)
    return g


def gen_cavity(numpad):
    out = {}
    for N in (6, 16):
        g = cavity_namespace(numpad, N)
        n = N * (3 * N - 2)
        # (i) fixed random state, SURVEY section 8(d) config 2(ii)
        u0 = 1e-2 * np.random.RandomState(0).standard_normal(n)
        for mode in ('adjoint', 'tangent'):
            u = numpad.array(u0.copy())
            res = g['cavity'](u, numpad.array(u0), 1.0, 0)
            pack('N%d.fixed.%s' % (N, mode), res.diff(u, mode), out)
        out['N%d.fixed.res' % N] = res._value.copy()
        # (ii) first Newton step of the script: u = 0 (value-pruned pattern)
        u = numpad.array(np.zeros(n))
        res = g['cavity'](u, numpad.array(np.zeros(n)), 1.0, 0)
        pack('N%d.zero.adjoint' % N, res.diff(u), out)
        out['N%d.zero.res' % N] = res._value.copy()
    np.savez_compressed(os.path.join(HERE, 'cavity_jacobian.npz'), **out)
    print('cavity jacobians done')


def gen_cavity_march(numpad, N, tag, max_steps=None):
    """Run the reference's own driver loop (test/cavity.py:97-114) and record
    every Newton linear system's solution, the accepted state of every time
    step and the adjoint of test/cavity_adjoint.py:6-8."""
    import scipy.sparse.linalg as spla
    import numpad.adsolve as adsolve
    g = cavity_namespace(numpad, N)
    g.update(t=0, dt=1.)
    rec = {'steps': [], 'n_newton': [], 'dt': [], 'lin_u': [], 'lin_du': []}
    orig = spla.spsolve

    def spy(J, b, use_umfpack=False):
        x = orig(J, b, use_umfpack=use_umfpack)
        rec['lin_du'].append(x.copy())
        return x
    adsolve.splinalg.spsolve = spy
    n = N * (3 * N - 2)
    u_and_p = numpad.zeros(n)
    force = numpad.zeros(n)
    t, dt = 0, 1.
    with contextlib.redirect_stdout(io.StringIO()):
        while True:
            u_and_p = numpad.solve(g['cavity'], u_and_p,
                                   args=(u_and_p, dt, force),
                                   rel_tol=0.05, abs_tol=1E-8)
            rec['steps'].append(u_and_p._value.copy())
            rec['n_newton'].append(u_and_p._n_Newton)
            rec['dt'].append(dt)
            if u_and_p._n_Newton == 1:
                break
            elif u_and_p._n_Newton < 4:
                dt *= 2
            t += dt
            if max_steps and len(rec['steps']) >= max_steps:
                break
            u_and_p.obliviate()
    adsolve.splinalg.spsolve = orig
    # adjoint of the kinetic energy w.r.t. the forcing through the last solve
    ke = 0.5 * numpad.sum(u_and_p[N * N:] ** 2)
    adj = ke.diff(force).toarray().ravel()
    out = {
        'N': np.int64(N),
        'states': np.array(rec['steps']),
        'n_newton': np.array(rec['n_newton'], np.float64),
        'dt': np.array(rec['dt']),
        'n_linear_solves': np.int64(len(rec['lin_du'])),
        'final_t': np.float64(t), 'final_dt': np.float64(dt),
        'final_res_norm': np.float64(u_and_p._res_norm),
        'kinetic_energy': np.float64(ke._value),
        'adjoint': adj,
    }
    if tag == 'small':
        out['lin_du'] = np.array(rec['lin_du'])
    else:      # keep the N=50 fixture small: first/last states + all step norms
        out['state_norms'] = np.linalg.norm(out['states'], axis=1)
        out['states'] = out['states'][[0, 1, -1]]
    np.savez_compressed(os.path.join(HERE, 'cavity_march_%s.npz' % tag), **out)
    print('cavity march', tag, 'N=%d: %d steps, %d linear solves, final t=%g'
          % (N, len(rec['steps']), len(rec['lin_du']), t),
          '|u|=%.15e |adj|=%.15e' % (np.linalg.norm(u_and_p._value),
                                     np.linalg.norm(adj)))


def gen_poisson_solve(numpad):
    """Known-answer Newton + implicit tangent/adjoint, adsolve.py:275-356."""
    out = {}
    N = 40
    dx = np.float64(1.) / (N + 1)

    def residual(u, f, dx):
        w = numpad.hstack([0, u, 0])
        return (w[2:] + w[:-2] - 2 * w[1:-1]) / dx ** 2 + f
    f = numpad.ones(N)
    dxa = numpad.array(dx)
    with contextlib.redirect_stdout(io.StringIO()):
        u = numpad.solve(residual, numpad.ones(N), args=(f, dxa), verbose=False)
    out['u'] = u._value.copy()
    pack('du_df.adjoint', numpad.sum(u).diff(f), out)
    pack('du_ddx.tangent', u.diff(dxa), out)
    pack('du_df.tangent', u.diff(f, 'tangent'), out)
    np.savez_compressed(os.path.join(HERE, 'poisson_solve.npz'), **out)
    print('poisson solve done')


def kec_namespace(numpad, script, Ni, Nj):
    """exec the part of test/euler_kec.py / test/navierstokes.py above the time
    integration with a stub `pylab` (numpy names, no plotting) and bind the
    module globals the residual reads"""
    import types
    stub = types.ModuleType('pylab')
    stub.__dict__.update({k: getattr(np, k) for k in ('pi', 'sin', 'cos', 'sqrt', 'linspace',
                                                       'arange', 'meshgrid', 'vstack')})
    sys.modules['pylab'] = stub
    src = open(os.path.join(REF, 'test', script)).read()
    src = src.split('# ---------------------- time integration')[0]
    src = src.replace("sys.path.append('../..')", '')
    g = {}
    exec(src, g)
    sys.path.insert(0, ROOT)
    from numpad_b200.workloads.kec import synthetic_mesh
    x, y = synthetic_mesh(Ni, Nj)
    g.update(Ni=Ni, Nj=Nj, p_out=1E5, mu=1.0,
             pt_in=1.2E5 if script.startswith('navier') else 1.05E5)
    g['geo'] = g['geo2d']([x, y])
    return g


def gen_kec(numpad):
    out = {}
    Ni, Nj = 7, 5
    for script, fn, tag in (('euler_kec.py', 'euler_kec', 'euler'),
                            ('navierstokes.py', 'ns_kec', 'ns')):
        g = kec_namespace(numpad, script, Ni, Nj)
        w = np.empty([4, Ni, Nj])
        w[0] = 1.0
        w[1] = 0.3 * np.sqrt(1.4E5)
        w[2] = 0.0
        w[3] = 1E5 / 0.4 + 0.5 * w[1] ** 2
        w *= 1 + 1e-3 * np.random.RandomState(0).standard_normal(w.shape)
        w0 = w.ravel()
        dt = 0.1 if tag == 'euler' else 1. / Nj
        for mode in ('adjoint', 'tangent'):
            u = numpad.array(w0.copy())
            res = g[fn](u, numpad.array(w0.copy()), g['geo'], dt)
            pack('%s.%s' % (tag, mode), res.diff(u, mode), out)
        out['%s.res' % tag] = res._value.copy()
        out['%s.w0' % tag] = w0
        out['%s.states' % tag] = np.int64(res._current_state._state_id - u._initial_state._state_id)
    out['Ni'], out['Nj'] = np.int64(Ni), np.int64(Nj)
    np.savez_compressed(os.path.join(HERE, 'kec_jacobian.npz'), **out)
    print('kec jacobians done: euler nnz', int(out['euler.adjoint.raw_nnz']),
          'ns nnz', int(out['ns.adjoint.raw_nnz']))


if __name__ == '__main__':
    numpad = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == 'kec':
        gen_kec(numpad)
        sys.exit(0)
    gen_opcases(numpad)
    gen_cavity(numpad)
    gen_poisson_solve(numpad)
    gen_cavity_march(numpad, 8, 'small')
    gen_cavity_march(numpad, 50, 'script')
    gen_kec(numpad)
