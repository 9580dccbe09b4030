"""Block-tridiagonal direct solver over the level sets of the matrix graph (host side of
csrc/blocktri.cu).

Replaces `splinalg.spsolve(J, rhs)` (SuperLU; reference adsolve.py:193-195) and
`factorized(J^T)` (adsolve.py:78-82) on the systems where preconditioned Krylov does not
converge: the convection-dominated saddle points of the reference's own cavity march from
rest (test/cavity.py:97-114) and the Euler / Navier-Stokes Jacobians at the scripts' time steps
(test/euler_kec.py:223-257, test/navierstokes.py:262-296).

  levels   : BFS from a pseudo-peripheral node over the graph of A + A^T (host, cached per
             pattern: it is the same for every Newton iteration of a run);
  ordering : unknowns sorted by level; in that order A is block tridiagonal (L_k | D_k | U_k);
  factor   : S_0 = D_0, S_k = D_k - L_k S_{k-1}^-1 U_{k-1}; the dense inverses S_k^-1 are kept
             (sum m_k^2 doubles).  L_k S^-1 and (L_k S^-1) U_{k-1} are window products of the
             CSR matrix with dense rows (a few rows each, csrc/blocktri.cu); the dense
             inversion is torch.linalg.inv_ex (cuSOLVER getrf + getrs: a plain library call);
  solve    : forward and backward sweep, one dense matrix-vector product per level and sweep
             (HBM-bound: the stored inverses are read twice), A^T x = b from the same factors.

Memory is the limit: 51 GB for the 1024^2 cavity (n = 3.1 M), O(N^3) on an N x N grid;
`feasible()` checks it against the free HBM before anything is allocated.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._dev import device, stream_ptr, _call

_level_cache = {}
# fraction of the free HBM the stored inverses may take
MEMORY_FRACTION = 0.8


def _pattern_key(A):
    n, nnz = A.shape[0], A.nnz
    w = torch.arange(1, nnz + 1, dtype=torch.int64, device=A.indices.device)
    return (n, nnz, int((A.indices.to(torch.int64) * (w % 8191)).sum().item()),
            int(A.indptr.to(torch.int64).sum().item()))


def level_structure(A):
    """-> (perm, inv, offsets): perm[new] = old index (device int64), offsets (python list,
    len = levels + 1).  Cached on the sparsity pattern."""
    key = _pattern_key(A)
    hit = _level_cache.get(key)
    if hit is not None and torch.equal(hit[0], A.indptr) and torch.equal(hit[1], A.indices):
        return hit[2]
    n = A.shape[0]
    At = A.transpose().materialized()
    ptr = np.ascontiguousarray(A.indptr.cpu().numpy())
    idx = np.ascontiguousarray(A.indices.cpu().numpy())
    ptr_t = np.ascontiguousarray(At.indptr.cpu().numpy())
    idx_t = np.ascontiguousarray(At.indices.cpu().numpy())
    level = np.empty(n, np.int32)
    nlev = ctypes.c_int32(0)
    rc = _lib.lib().cdll.npb_bfs_levels_host(n, ptr.ctypes.data, idx.ctypes.data, ptr_t.ctypes.data,
                                             idx_t.ctypes.data, level.ctypes.data, ctypes.byref(nlev))
    if rc != 0:
        raise RuntimeError('npb_bfs_levels_host failed')
    counts = np.bincount(level, minlength=nlev.value)
    offsets = [0] + [int(v) for v in np.cumsum(counts)]
    lev_d = torch.from_numpy(level).to(device())
    perm = torch.sort(lev_d.to(torch.int64), stable=True)[1]
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, dtype=torch.int64, device=device())
    out = (perm, inv, offsets)
    if len(_level_cache) > 4:
        _level_cache.clear()
    _level_cache[key] = (A.indptr.clone(), A.indices.clone(), out)
    return out


def memory_bytes(offsets):
    m = np.diff(np.asarray(offsets, dtype=np.float64))
    return float(8.0 * (m * m).sum() + 4 * 8.0 * (m.max() ** 2 if m.size else 0.0))


def feasible(A, budget_bytes=None):
    """can the dense inverses of the diagonal blocks be held in HBM?"""
    if A.shape[0] != A.shape[1] or A.nnz == 0:
        return False
    _, _, offsets = level_structure(A)
    if max(np.diff(offsets)) > 60000:
        return False
    if budget_bytes is None:
        free, _ = torch.cuda.mem_get_info()
        # torch's caching allocator may hold freed blocks that it can hand back
        free += torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
        budget_bytes = MEMORY_FRACTION * free
    return memory_bytes(offsets) <= budget_bytes


class BlockTriLU:
    """P A P^T = block tridiagonal; dense inverses of the Schur complements on the device."""

    def __init__(self, A):
        from .shard import permuted
        A = A.materialized()
        self.n = n = A.shape[0]
        self.perm, self.inv, self.off = level_structure(A)
        off = self.off
        K = self.K = len(off) - 1
        self.Ap = permuted(A, self.perm.to(torch.int32), self.inv.to(torch.int32)).materialized()
        self.At = self.Ap.transpose().materialized()
        Ap, At = self.Ap, self.At
        dev = device()
        mmax = max(off[k + 1] - off[k] for k in range(K))
        self.mmax = mmax
        self.Sinv = []
        infos = []
        for k in range(K):
            r0, r1 = off[k], off[k + 1]
            m = r1 - r0
            # S^T is assembled (not S): the library inverse comes back column-major, and the
            # column-major image of (S^T)^-1 IS the row-major S^-1 the kernels read
            if k == 0:
                St = torch.zeros((m, m), dtype=torch.float64, device=dev)
            else:
                q0, mp = off[k - 1], r0 - off[k - 1]
                # Y = L_k S_{k-1}^-1   (m x mp): each row combines the few rows of S^-1 that
                # the row of L_k references
                Y = torch.empty((m, mp), dtype=torch.float64, device=dev)
                _call('npb_bt_window_spmm', Ap.indptr, Ap.indices, Ap.data, r0, r1, q0, r0,
                      self.Sinv[k - 1], mp, mp, 1.0, 0.0, Y, mp)
                # -W^T = -U_{k-1}^T Y^T  (m x m): rows of A^T in level k, columns in level k-1
                Yt = Y.t().contiguous()
                St = torch.empty((m, m), dtype=torch.float64, device=dev)
                _call('npb_bt_window_spmm', At.indptr, At.indices, At.data, r0, r1, q0, r0,
                      Yt, m, m, -1.0, 0.0, St, m)
                del Y, Yt
            _call('npb_bt_window_to_dense', At.indptr, At.indices, At.data, r0, r1, r0, r1, 1.0, St, m)
            Ri, info = torch.linalg.inv_ex(St)
            del St
            Si = Ri.t()
            self.Sinv.append(Si if Si.is_contiguous() else Si.contiguous())
            infos.append(info)
        bad = torch.stack(infos).ne(0).nonzero().reshape(-1)
        if bad.numel():
            raise np.linalg.LinAlgError('block-tridiagonal LU: singular diagonal block at level %d'
                                        % int(bad[0]))
        self.scratch = torch.empty(int(_lib.lib().cdll.npb_bt_gemv_scratch_doubles(mmax)),
                                   dtype=torch.float64, device=dev)
        self.bytes = sum(s.numel() for s in self.Sinv) * 8

    def describe(self):
        return ('block-tridiagonal LU: %d levels, largest block %d, %.1f GB of dense inverses'
                % (self.K, self.mmax, self.bytes / 1e9))

    def solve(self, b, transpose=False):
        """x = A^-1 b (or A^-T b)"""
        off, K = self.off, self.K
        M = self.At if transpose else self.Ap
        tr = 1 if transpose else 0
        bp = b.reshape(-1)[self.perm].contiguous()
        z = torch.empty_like(bp)
        c = torch.empty(self.mmax, dtype=torch.float64, device=bp.device)
        st = stream_ptr()
        L = _lib.lib().cdll
        ip, ix, dat = M.indptr.data_ptr(), M.indices.data_ptr(), M.data.data_ptr()
        zp, bpp, cp, sp_ = z.data_ptr(), bp.data_ptr(), c.data_ptr(), self.scratch.data_ptr()
        rc = 0
        # forward: c_k = b_k - (lower block) z_{k-1};  z_k = S_k^-1 c_k
        for k in range(K):
            r0, r1 = off[k], off[k + 1]
            m = r1 - r0
            if k == 0:
                src = bpp + 8 * r0
            else:
                q0 = off[k - 1]
                rc |= L.npb_bt_window_spmv(ip, ix, dat, r0, r1, q0, r0, zp + 8 * q0, -1.0, 0.0, cp, st)
                rc |= L.npb_axpby(m, 1.0, bpp + 8 * r0, 1.0, cp, st)
                src = cp
            rc |= L.npb_bt_gemv(m, self.Sinv[k].data_ptr(), m, src, 1.0, 0.0, None, zp + 8 * r0, tr, sp_, st)
        # backward: x_k = z_k - S_k^-1 (upper block) x_{k+1}   (in place in z)
        for k in range(K - 2, -1, -1):
            r0, r1 = off[k], off[k + 1]
            m = r1 - r0
            rc |= L.npb_bt_window_spmv(ip, ix, dat, r0, r1, r1, off[k + 2], zp + 8 * r1, 1.0, 0.0, cp, st)
            rc |= L.npb_bt_gemv(m, self.Sinv[k].data_ptr(), m, cp, -1.0, 1.0, zp + 8 * r0, zp + 8 * r0, tr, sp_, st)
        if rc:
            raise RuntimeError('block-tridiagonal solve failed: %s' % _lib.lib().last_error())
        x = torch.empty_like(z)
        x[self.perm] = z
        return x

    def apply(self, r, z):
        z.copy_(self.solve(r))

    def apply_t(self, r, z):
        z.copy_(self.solve(r, transpose=True))
