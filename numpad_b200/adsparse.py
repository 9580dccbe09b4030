"""Differentiable sparse matrix and sparse solve (reference adsparse.py:16-78) on the device.

`csr_matrix((data, (i, j)), shape)` / `csr_matrix((data, indices, indptr), shape)` keeps the
entries as an adarray `data`, so that `A * b` records TWO tape edges -- d(Ab)/db = A and
d(Ab)/d data = the (row i_k, entry k) matrix holding b[j_k] -- and `spsolve(A, b)` is an
`adsolution` whose residual is `A * x - b` (differentiable with respect to data and b).
"""
import numpy as np
import scipy.sparse as sp
import torch

from .adarray import adarray, array, value, transpose
from .adsolve import adsolution
from ._dev import DevCsr, csr_jac, device
from . import linsolve


class csr_matrix:
    def __init__(self, data, shape=None):
        if len(data) == 3:
            vals, col_ind, row_ptr = data
            vals = array(vals)
            host = sp.csr_matrix((np.asarray(value(vals)), np.asarray(value(col_ind), int),
                                  np.asarray(value(row_ptr), int)), shape=shape)
            i = np.repeat(np.arange(host.shape[0]), np.diff(host.indptr))
            j = host.indices.copy()
        elif len(data) == 2:
            vals, (i, j) = data
            vals = array(vals)
            i, j = np.asarray(value(i), int), np.asarray(value(j), int)
        else:
            raise NotImplementedError()
        self.data = vals.copy()
        self.i, self.j = i.copy(), j.copy()
        self.shape = tuple(shape) if shape is not None else (int(i.max()) + 1, int(j.max()) + 1)
        self._i_dev = torch.from_numpy(self.i.astype(np.int32)).to(device())
        self._j_dev = torch.from_numpy(self.j.astype(np.int32)).to(device())

    @property
    def _value(self):
        """scipy image of the current values (duplicates summed), as in the reference"""
        return sp.csr_matrix((np.asarray(value(self.data)).ravel(), (self.i, self.j)), shape=self.shape)

    def _dev_csr(self):
        return DevCsr.from_coo(self.shape, self._i_dev, self._j_dev, self.data._dev.reshape(-1))

    def __mul__(self, b):
        b = array(b)
        if b.ndim != 1:
            shape = b.shape[1:]
            cols = b.reshape([b.shape[0], -1])
            out = transpose([self * cols[:, k] for k in range(cols.shape[1])])
            return out.reshape((out.shape[0],) + shape)
        assert b.shape[0] == self.shape[1]
        A = self._dev_csr()
        out = adarray(A.matvec(b._dev))
        out.next_state(A, b, '*')
        k = torch.arange(self.data.size, dtype=torch.int32, device=device())
        out.next_state(csr_jac(b._dev[self._j_dev.long()], self._i_dev, k,
                               shape=(self.shape[0], self.data.size)), self.data, '*')
        return out


def spsolve(A, b):
    """AD counterpart of scipy.sparse.linalg.spsolve on the device solver."""
    b = array(b)
    x = adarray(linsolve.solve(A._dev_csr(), b._dev.reshape(-1)).reshape(b.shape))
    return adsolution(x, A * x - b, 1)
