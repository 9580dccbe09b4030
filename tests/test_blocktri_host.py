"""Host-side logic of the block-tridiagonal solver (csrc/blocktri.cu): the level structures.
One-sided BFS levels and the two-sided structure of the twisted factorisation must make the
matrix block tridiagonal (|level(i) - level(j)| <= 1 for every entry), cover every unknown, and
-- two-sided -- balance the two elimination chains and seed them on unknowns with a diagonal.
No GPU needed: the functions take host arrays."""
import ctypes

import numpy as np
import scipy.sparse as sp

from numpad_b200 import _lib


def _saddle(N):
    """a MAC-grid Stokes-like saddle point [[A, G], [D, 0]] (pressure rows have no diagonal)"""
    def d1(n):
        return sp.diags([-np.ones(n - 1), np.ones(n - 1)], [0, 1], shape=(n - 1, n))
    I = sp.identity
    Gx = sp.kron(d1(N), I(N))                  # (N-1)N x N^2
    Gy = sp.kron(I(N), d1(N))                  # N(N-1) x N^2
    lap = lambda m: sp.diags([-1, 4, -1], [-1, 0, 1], shape=(m, m))
    Ax = sp.kron(lap(N - 1), I(N)) + sp.kron(I(N - 1), lap(N))
    Ay = sp.kron(lap(N), I(N - 1)) + sp.kron(I(N), lap(N - 1))
    A = sp.block_diag([Ax, Ay])
    G = sp.vstack([Gx, Gy])
    J = sp.bmat([[None, -G.T], [G, A]]).tocsr()      # pressure first, like the cavity layout
    J.sort_indices()
    return J


def _levels(J, two_sided):
    L = _lib.lib().cdll
    Jt = J.T.tocsr()
    n = J.shape[0]
    c = lambda a: np.ascontiguousarray(a.astype(np.int32))
    p, i, pt, it = c(J.indptr), c(J.indices), c(Jt.indptr), c(Jt.indices)
    lev = np.empty(n, np.int32)
    nl, mid = ctypes.c_int32(0), ctypes.c_int32(-1)
    if two_sided:
        hd = np.ascontiguousarray((J.diagonal() != 0).astype(np.uint8))
        rc = L.npb_bfs_levels_two_sided_host(n, p.ctypes.data, i.ctypes.data, pt.ctypes.data,
                                             it.ctypes.data, hd.ctypes.data, lev.ctypes.data,
                                             ctypes.byref(nl), ctypes.byref(mid))
    else:
        rc = L.npb_bfs_levels_host(n, p.ctypes.data, i.ctypes.data, pt.ctypes.data, it.ctypes.data,
                                   lev.ctypes.data, ctypes.byref(nl))
    return rc, lev, nl.value, mid.value


def test_one_sided_levels_are_block_tridiagonal():
    J = _saddle(20)
    rc, lev, K, _ = _levels(J, False)
    assert rc == 0 and lev.min() == 0 and lev.max() == K - 1
    coo = J.tocoo()
    assert np.abs(lev[coo.row] - lev[coo.col]).max() <= 1
    assert (np.bincount(lev, minlength=K) > 0).all()


def test_two_sided_levels_balance_the_chains():
    J = _saddle(24)
    rc, lev, K, M = _levels(J, True)
    assert rc == 0 and 0 < M < K - 1
    coo = J.tocoo()
    assert np.abs(lev[coo.row] - lev[coo.col]).max() <= 1        # still block tridiagonal
    m = np.bincount(lev, minlength=K)
    assert (m > 0).all()
    top, bottom = m[:M].sum(), m[M + 1:].sum()
    assert abs(top - bottom) <= 0.1 * J.shape[0]
    # the seeds of both chains (levels 0 and K-1) carry a diagonal entry
    d = J.diagonal()
    assert (d[lev == 0] != 0).all() and (d[lev == K - 1] != 0).all()
    # numerically: the twisted elimination over these levels solves the system
    perm = np.argsort(lev, kind='stable')
    off = np.concatenate([[0], np.cumsum(m)])
    Jp = J[perm][:, perm].tocsr()
    blk = lambda a, b: Jp[off[a]:off[a + 1], off[b]:off[b + 1]]
    # pin the pressure level (the Stokes saddle point is singular up to a constant)
    Jp = Jp.tolil()
    prow = int(np.nonzero(Jp.diagonal() == 0)[0][0])
    Jp[prow, prow] = 1.0
    Jp = Jp.tocsr()
    Sinv = [None] * K
    for k in range(M):
        S = blk(k, k).toarray() - (blk(k, k - 1) @ (Sinv[k - 1] @ blk(k - 1, k).toarray()) if k else 0)
        Sinv[k] = np.linalg.inv(S)
    for k in range(K - 1, M, -1):
        S = blk(k, k).toarray() - (blk(k, k + 1) @ (Sinv[k + 1] @ blk(k + 1, k).toarray()) if k < K - 1 else 0)
        Sinv[k] = np.linalg.inv(S)
    S = blk(M, M).toarray() - blk(M, M - 1) @ (Sinv[M - 1] @ blk(M - 1, M).toarray()) \
        - blk(M, M + 1) @ (Sinv[M + 1] @ blk(M + 1, M).toarray())
    assert np.isfinite(np.linalg.cond(S)) and np.linalg.cond(S) < 1e10


def test_two_sided_refuses_disconnected_graphs():
    J = sp.block_diag([_saddle(6), _saddle(6)]).tocsr()
    rc, _, _, _ = _levels(J, True)
    assert rc != 0
    rc, lev, K, _ = _levels(J, False)             # the one-sided levels number the components in turn
    coo = J.tocoo()
    assert rc == 0 and np.abs(lev[coo.row] - lev[coo.col]).max() <= 1


def test_host_mirror_flags_writes_only():
    """adarray._HostMirror (pure numpy): writes through the mirror or its views flag the owner,
    reads and arithmetic do not and give plain ndarrays; a detached mirror flags nobody"""
    import weakref
    from numpad_b200.adarray import _HostMirror

    class Owner:
        _mirror_gen, _mirror_dirty = 3, False
    o = Owner()
    m = np.arange(6.0).view(_HostMirror)
    m._owner = (weakref.ref(o), 3)
    assert type(m * 2) is np.ndarray and type(np.sqrt(m)) is np.ndarray and float(m.sum()) == 15.0
    assert not o._mirror_dirty
    m[1:3][0] = 9.0                      # through a view
    assert o._mirror_dirty and m[1] == 9.0
    o._mirror_dirty = False
    m -= 1.0
    assert o._mirror_dirty and type(m) is _HostMirror and m[0] == -1.0
    o._mirror_dirty = False
    np.add(m, 1.0, out=m)
    assert o._mirror_dirty
    o._mirror_dirty, o._mirror_gen = False, 4          # the device values moved on: detached
    m[0] = 5.0
    assert not o._mirror_dirty
