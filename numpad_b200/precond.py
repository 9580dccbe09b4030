"""Preconditioner selection for the Newton linear solve (seam S3).

build(A, kind) -> (M, method):
  'none'   : M is None -- plain restarted GMRES (full GMRES when n is tiny)
  'direct' : M.solve(b) is a device banded LU (csrc/bandlu.cu), for systems
             small enough that n * bandwidth^2 is cheap
  'lsc'    : block preconditioner for saddle-point Jacobians (rows with a
             zero diagonal = constraints): least-squares-commutator Schur
             complement, multigrid sub-solves (csrc/amg.cu)
  'auto'   : picks from the structure of A
"""
from . import _dev


def build(A, kind='auto'):
    n = A.shape[0]
    if kind == 'auto':
        kind = 'none'
    if kind == 'none':
        return None, 'gmres'
    raise ValueError('unknown preconditioner %r' % kind)
