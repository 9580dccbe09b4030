"""CPU (scipy) prototype of momentum-block solvers for convection-dominated cavity
Jacobians (states from tools/proto_states.py).  Test infrastructure only.

    python tools/proto_robust.py --states /tmp/proto_states_128.npz --keys s0_i1,s5_i1 --variants exactA,upw
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proto_lsc as pl                                  # noqa: E402
from proto_lsc import Hier, pcg, fgmres, mv, WORK       # noqa: E402


def upwinded(A):
    """algebraic artificial diffusion (Kuzmin): d_ij = max(0, a_ij, a_ji) for i != j;
    a_ij -= d_ij, a_ii += d_ij: all off-diagonals <= 0, row sums unchanged."""
    A = A.tocsr()
    At = A.T.tocsr()
    # union pattern
    U = (abs(A) + abs(At)).tocsr()
    U.sort_indices()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(U.indptr))
    cols = U.indices
    a = np.asarray(A[rows, cols]).ravel()
    at = np.asarray(At[rows, cols]).ravel()
    off = rows != cols
    d = np.where(off, np.maximum(0.0, np.maximum(a, at)), 0.0)
    newv = a - d
    B = sp.csr_matrix((newv, (rows, cols)), shape=A.shape)
    B = B + sp.diags(np.bincount(rows, weights=d, minlength=A.shape[0]))
    return B.tocsr()


def inner_gmres(matvec, prec, b, k):
    """k steps of right-preconditioned GMRES from x0 = 0 (no restart)"""
    x, its, rr = fgmres(matvec, prec, b, rtol=1e-14, restart=k, maxiter=k)
    return x


def run(J, b, name, maxiter=300, rtol=1e-10):
    n = J.shape[0]
    d = J.diagonal()
    vs, ps = np.nonzero(d != 0)[0], np.nonzero(d == 0)[0]
    A, G, D = J[vs][:, vs].tocsr(), J[vs][:, ps].tocsr(), J[ps][:, vs].tocsr()
    L = (-(D @ G)).tocsr()
    t0 = time.time()
    hl = Hier(L, theta=0.0, nu=1)
    cl = 4
    opts = dict(x.split('=') for x in name.split(':')[1:])
    kind = name.split(':')[0]
    nu = int(opts.get('nu', 2))
    k_in = int(opts.get('k', 4))
    luA = None
    if kind == 'exactA':
        luA = spla.splu(A.tocsc())
        asolve = luA.solve
    elif kind == 'base':
        ha = Hier(A, theta=0.25, nu=nu)
        asolve = ha.vcycle
    elif kind == 'upw':            # V-cycle of the upwinded operator as A^-1
        Au = upwinded(A)
        ha = Hier(Au, theta=0.25, nu=nu, agg=opts.get('agg', 'mis1'))
        asolve = ha.vcycle
    elif kind == 'upwg':           # k GMRES steps on A, preconditioned by that V-cycle
        Au = upwinded(A)
        ha = Hier(Au, theta=0.25, nu=nu, agg=opts.get('agg', 'mis1'))
        asolve = lambda r: inner_gmres(lambda v: mv(A, v), ha.vcycle, r, k_in)
    elif kind == 'jacg':           # current robust mode: k Jacobi-GMRES steps
        dinv = 1.0 / A.diagonal()
        asolve = lambda r: inner_gmres(lambda v: mv(A, v), lambda v: dinv * v, r, k_in)
    elif kind == 'ilug':           # k GMRES steps preconditioned by ILU(0)
        ilu = spla.spilu(A.tocsc(), drop_tol=0.0, fill_factor=1.0, permc_spec='NATURAL',
                         diag_pivot_thresh=0.0)
        asolve = lambda r: inner_gmres(lambda v: mv(A, v), ilu.solve, r, k_in)
    else:
        raise ValueError(kind)
    exactL = opts.get('L') == 'exact'
    luL = spla.splu(L.tocsc()) if exactL else None

    def prec(r):
        rv, rp = r[vs], r[ps]
        p1 = luL.solve(rp) if exactL else pcg(hl, L, rp, cl)
        q = mv(D, mv(A, mv(G, p1)))
        p = -(luL.solve(q) if exactL else pcg(hl, L, q, cl))
        rv = rv - mv(G, p)
        u = asolve(rv)
        z = np.empty(n)
        z[vs], z[ps] = u, p
        return z

    WORK[0] = 0.0
    x, its, rr = fgmres(lambda v: mv(J, v), prec, b, rtol=rtol, maxiter=maxiter)
    w = WORK[0]
    jb = 12.0 * J.nnz + 20.0 * n
    print('  %-22s its %3d relres %.1e  work/it = %6.1f J-SpMV  total = %7.0f J-SpMV  %.0fs'
          % (name, its, rr, w / jb / max(its, 1), w / jb, time.time() - t0), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--states', default='/tmp/proto_states_128.npz')
    ap.add_argument('--keys', default='s0_i1,s5_i1')
    ap.add_argument('--variants', default='exactA,base,upw,upwg')
    ap.add_argument('--maxiter', type=int, default=300)
    a = ap.parse_args()
    Z = np.load(a.states)
    for k in a.keys.split(','):
        J = sp.csr_matrix((Z[k + '.data'], Z[k + '.indices'], Z[k + '.indptr']))
        b = Z[k + '.rhs']
        print('%s: n=%d nnz=%d' % (k, J.shape[0], J.nnz), flush=True)
        for v in a.variants.split(','):
            try:
                run(J, b, v, maxiter=a.maxiter)
            except Exception as e:
                print('  %-22s FAILED %r' % (v, e), flush=True)
