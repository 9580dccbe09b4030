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


def drop_noise(m, noise):
    """remove stored entries that are cancellation residue: |v| <= noise * the
    largest magnitude in their row"""
    m = canonical(m)
    rowmax = np.maximum.reduceat(np.abs(np.r_[m.data, 0.0]),
                                 np.minimum(m.indptr[:-1], m.nnz))
    rowmax[np.diff(m.indptr) == 0] = 0.0
    rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    keep = np.abs(m.data) > noise * rowmax[rows]
    out = sp.csr_matrix((m.data[keep], (rows[keep], m.indices[keep])), shape=m.shape)
    return out, int((~keep).sum())


def assert_csr_parity(got, want, rtol=1e-12, raw_nnz=None, what='', noise=None):
    """canonical (indptr, indices) bit-exact, values
    within rtol * max|J| entrywise.

    noise: entries that are exact cancellations up to the ORDER of a floating
    point sum (|v| <= noise * row max, e.g. 1e-14 against 4e4) are stored or
    pruned depending on scipy's internal, unsorted summation order, which the
    engine does not reproduce; with `noise` set they are removed from both
    sides first and their number must stay below 1e-4 of the pattern."""
    if noise is not None:
        got, ng = drop_noise(got, noise)
        want, nw = drop_noise(want, noise)
        assert max(ng, nw) <= 1e-4 * max(want.nnz, 1) + 2, (what, ng, nw)
        raw_nnz = None
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
