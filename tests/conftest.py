import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """GPU tests fail (not skip) on a box without a device unless deselected
    with -m "not gpu": the product has no CPU fallback."""


def load_csr(z, prefix):
    shape = tuple(int(s) for s in z[prefix + '.shape'])
    return sp.csr_matrix((z[prefix + '.data'], z[prefix + '.indices'],
                          z[prefix + '.indptr']), shape=shape)


def canonical(m):
    m = sp.csr_matrix(m, copy=True)
    m.sum_duplicates()
    m.sort_indices()
    return m


def assert_csr_parity(got, want, rtol=1e-12, raw_nnz=None, what='', noise=None):
    """canonical (indptr, indices) bit-exact, values
    within rtol * max|J| entrywise.

    noise: entries that are exact cancellations up to the ORDER of a floating
    point sum (|v| <= noise * row max, e.g. 1e-14 against 4e4) are stored or
    pruned depending on scipy's internal, unsorted summation order, which the
    engine does not reproduce.  With `noise` set, the two patterns may differ
    in such entries only, and in fewer than 1e-4 of all entries."""
    if noise is not None:
        g, w = canonical(got), canonical(want)
        assert g.shape == w.shape, (what, g.shape, w.shape)
        pg, pw = g.copy(), w.copy()
        pg.data[:] = 1.0
        pw.data[:] = 1.0
        diff = (pg - pw).tocoo()              # +1: only in got, -1: only in want
        assert diff.nnz <= 1e-4 * max(w.nnz, 1) + 2, (what, 'pattern differs in', diff.nnz)
        rowmax = np.maximum(abs(g).max(axis=1).toarray().ravel(),
                            abs(w).max(axis=1).toarray().ravel())
        for i, j, s_ in zip(diff.row, diff.col, diff.data):
            v = g[i, j] if s_ > 0 else w[i, j]
            assert abs(v) <= noise * rowmax[i], (what, i, j, v, rowmax[i])
        scale = np.abs(w.data).max() if w.nnz else 1.0
        err = abs(g - w)
        assert (err.max() if err.nnz else 0.0) <= rtol * scale, (what, 'values')
        return
    if raw_nnz is not None:
        assert sp.csr_matrix(got).nnz == raw_nnz, \
            '%s raw nnz %d != %d' % (what, sp.csr_matrix(got).nnz, raw_nnz)
    g, w = canonical(got), canonical(want)
    assert g.shape == w.shape, (what, g.shape, w.shape)
    assert np.array_equal(g.indptr, w.indptr), what + ' indptr differs'
    assert np.array_equal(g.indices, w.indices), what + ' indices differ'
    scale = np.abs(w.data).max() if w.nnz else 1.0
    err = np.abs(g.data - w.data).max() if w.nnz else 0.0
    assert err <= rtol * (scale if scale > 0 else 1.0), \
        '%s values differ: %g (scale %g)' % (what, err, scale)


@pytest.fixture(scope='session')
def golden():
    class G:
        def __getattr__(self, name):
            z = np.load(os.path.join(GOLDEN, name + '.npz'))
            setattr(self, name, z)
            return z
    return G()
