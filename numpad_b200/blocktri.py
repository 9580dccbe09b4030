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
  factor   : twisted -- two elimination chains, from the first level down and from the last
             level up (S_k = D_k - L_k S_{k-1}^-1 U_{k-1}, resp. D_k - U_k S_{k+1}^-1 L_{k+1}),
             issued alternately on two streams, meet at a middle level; the dense inverses are kept
             (sum m_k^2 doubles).  L_k S^-1 and (L_k S^-1) U_{k-1} are window products of the
             CSR matrix with dense rows (a few rows each, csrc/blocktri.cu); the dense
             inversion is torch.linalg.inv_ex (cuSOLVER getrf + getrs: a plain library call);
  solve    : forward and backward sweep, one dense matrix-vector product per level and sweep
             (HBM-bound: the stored inverses are read twice), A^T x = b from the same factors.

Memory is the limit: 51 GB for the 1024^2 cavity (n = 3.1 M), O(N^3) on an N x N grid;
`feasible()` checks it against the free HBM before anything is allocated.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._dev import device, stream_ptr, _call

_level_cache = {}
# fraction of the free HBM the stored inverses may take
MEMORY_FRACTION = 0.8
# twisted factorisation (two elimination chains on two streams); NPB_BT_TWISTED=0: one chain
TWISTED = os.environ.get('NPB_BT_TWISTED', '1') != '0'


def _pattern_key(A):
    n, nnz = A.shape[0], A.nnz
    w = torch.arange(1, nnz + 1, dtype=torch.int64, device=A.indices.device)
    return (n, nnz, int((A.indices.to(torch.int64) * (w % 8191)).sum().item()),
            int(A.indptr.to(torch.int64).sum().item()))


def level_structure(A, two_sided=True):
    """-> (perm, inv, offsets, middle): perm[new] = old index (device int64), offsets (python
    list, len = levels + 1), middle = the level at which the two elimination chains of the
    twisted factorisation meet (levels - 1: one chain).  two_sided: BFS balls around a
    pseudo-peripheral node and around the node farthest from it with one middle level between
    them (npb_bfs_levels_two_sided_host) -- both chains then eliminate growing BFS balls;
    otherwise (or for disconnected graphs) plain BFS levels from one node.  Cached on the
    sparsity pattern."""
    key = _pattern_key(A) + (bool(two_sided),)
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
    nlev, mid = ctypes.c_int32(0), ctypes.c_int32(-1)
    L = _lib.lib().cdll
    rc = -1
    if two_sided:
        from . import amg
        has_diag = np.ascontiguousarray((amg.diagonal(A) != 0).to(torch.uint8).cpu().numpy())
        rc = L.npb_bfs_levels_two_sided_host(n, ptr.ctypes.data, idx.ctypes.data, ptr_t.ctypes.data,
                                             idx_t.ctypes.data, has_diag.ctypes.data, level.ctypes.data,
                                             ctypes.byref(nlev), ctypes.byref(mid))
    if rc != 0:
        rc = L.npb_bfs_levels_host(n, ptr.ctypes.data, idx.ctypes.data, ptr_t.ctypes.data,
                                   idx_t.ctypes.data, level.ctypes.data, ctypes.byref(nlev))
        if rc != 0:
            raise RuntimeError('npb_bfs_levels_host failed')
        mid = ctypes.c_int32(nlev.value - 1)
    counts = np.bincount(level, minlength=nlev.value)
    offsets = [0] + [int(v) for v in np.cumsum(counts)]
    lev_d = torch.from_numpy(level).to(device())
    perm = torch.sort(lev_d.to(torch.int64), stable=True)[1]
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, dtype=torch.int64, device=device())
    out = (perm, inv, offsets, int(mid.value))
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
    _, _, offsets, _ = level_structure(A, TWISTED)
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

    def __init__(self, A, twisted=None):
        from .shard import permuted
        A = A.materialized()
        self.n = n = A.shape[0]
        twisted = TWISTED if twisted is None else bool(twisted)
        try:
            self._factor(A, twisted, permuted)
        except np.linalg.LinAlgError:
            if not twisted:
                raise
            self._factor(A, False, permuted)         # a singular block of the second chain

    def _factor(self, A, twisted, permuted):
        n = self.n
        self.perm, self.inv, self.off, M = level_structure(A, twisted)
        off = self.off
        K = self.K = len(off) - 1
        self.Ap = permuted(A, self.perm.to(torch.int32), self.inv.to(torch.int32)).materialized()
        self.At = self.Ap.transpose().materialized()
        Ap, At = self.Ap, self.At
        dev = device()
        mmax = max(off[k + 1] - off[k] for k in range(K))
        self.mmax = mmax
        # ---- twisted factorisation: two elimination chains, from level 0 downwards and from
        # level K-1 upwards, meet at level M (level_structure balances the unknowns they
        # eliminate: the dense inversions are latency-bound, ~5 us per column, not flop-bound).
        # The chains are independent until M, so they are issued alternately on two streams.
        self.M = M
        Sinv = [None] * K
        infos = []

        def contribution(k, nb):
            """-(coupling of level k with the eliminated neighbour level nb)^T as an m_k x m_k
            matrix: -(C S_nb^-1 B)^T with C = A[k, nb], B = A[nb, k]"""
            r0, r1 = off[k], off[k + 1]
            q0, q1 = off[nb], off[nb + 1]
            m, mp = r1 - r0, q1 - q0
            # Y = C S_nb^-1  (m x mp): each row combines the few rows of S^-1 that the row of C
            # references
            Y = torch.empty((m, mp), dtype=torch.float64, device=dev)
            _call('npb_bt_window_spmm', Ap.indptr, Ap.indices, Ap.data, r0, r1, q0, q1,
                  Sinv[nb], mp, mp, 1.0, 0.0, Y, mp)
            # -(Y B)^T = -B^T Y^T: rows of A^T in level k, columns in level nb
            Yt = Y.t().contiguous()
            Wt = torch.empty((m, m), dtype=torch.float64, device=dev)
            _call('npb_bt_window_spmm', At.indptr, At.indices, At.data, r0, r1, q0, q1,
                  Yt, m, m, -1.0, 0.0, Wt, m)
            return Wt

        def eliminate(k, nbs):
            """S_k^T = D_k^T - sum of the neighbours' contributions; keep S_k^-1 (row-major).
            S^T is assembled, not S: the library inverse comes back column-major, and the
            column-major image of (S^T)^-1 IS the row-major S^-1 the kernels read"""
            r0, r1 = off[k], off[k + 1]
            m = r1 - r0
            St = None
            for nb in nbs:
                Wt = contribution(k, nb)
                St = Wt if St is None else St.add_(Wt)
            if St is None:
                St = torch.zeros((m, m), dtype=torch.float64, device=dev)
            _call('npb_bt_window_to_dense', At.indptr, At.indices, At.data, r0, r1, r0, r1, 1.0, St, m)
            Ri, info = torch.linalg.inv_ex(St)
            Si = Ri.t()
            Sinv[k] = Si if Si.is_contiguous() else Si.contiguous()
            infos.append(info)

        if M == K - 1:
            for k in range(K):
                eliminate(k, [k - 1] if k > 0 else [])
        else:
            cur = torch.cuda.current_stream()
            sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
            sa.wait_stream(cur)
            sb.wait_stream(cur)
            ia, ib = 0, K - 1
            while ia < M or ib > M:
                if ia < M:
                    with torch.cuda.stream(sa):
                        eliminate(ia, [ia - 1] if ia > 0 else [])
                    ia += 1
                if ib > M:
                    with torch.cuda.stream(sb):
                        eliminate(ib, [ib + 1] if ib < K - 1 else [])
                    ib -= 1
            cur.wait_stream(sa)
            cur.wait_stream(sb)
            eliminate(M, [M - 1, M + 1])
            for t in Sinv:                       # allocated on the side streams, used on `cur`
                t.record_stream(cur)
        self.Sinv = Sinv
        bad = torch.stack(infos).ne(0).nonzero().reshape(-1)
        if bad.numel():
            raise np.linalg.LinAlgError('block-tridiagonal LU: singular diagonal block at level %d'
                                        % int(bad[0]))
        self.scratch = torch.empty(2 * int(_lib.lib().cdll.npb_bt_gemv_scratch_doubles(mmax)),
                                   dtype=torch.float64, device=dev)
        self.bytes = sum(s.numel() for s in self.Sinv) * 8
        # host descriptors + device work buffers of the solve phase
        self.off_h = np.asarray(off, dtype=np.int64)
        self.sinv_h = np.asarray([s.data_ptr() for s in self.Sinv], dtype=np.uint64)
        self.bp = torch.empty(n, dtype=torch.float64, device=dev)
        self.z = torch.empty(n, dtype=torch.float64, device=dev)
        self.c = torch.empty(2 * mmax, dtype=torch.float64, device=dev)     # one half per chain
        BlockTriLU._next_graph_id += 1
        self.graph_id = BlockTriLU._next_graph_id

    def describe(self):
        return ('block-tridiagonal LU: %d levels (chains meet at %d), largest block %d, %.1f GB of '
                'dense inverses' % (self.K, self.M, self.mmax, self.bytes / 1e9))

    _next_graph_id = 0

    def __del__(self):
        try:
            _lib.lib().cdll.npb_bt_graph_release(int(self.graph_id))
        except Exception:
            pass

    def solve(self, b, transpose=False):
        """x = A^-1 b (or A^-T b): both sweeps in ONE C call (csrc/blocktri.cu npb_bt_solve), which
        replays them as a CUDA graph from the second call on -- the work buffers below belong to
        the factorisation and never move"""
        M = self.At if transpose else self.Ap
        self.bp.copy_(b.reshape(-1)[self.perm])
        rc = _lib.lib().cdll.npb_bt_solve(self.K, self.M, self.off_h.ctypes.data, self.sinv_h.ctypes.data,
                                          M.indptr.data_ptr(), M.indices.data_ptr(), M.data.data_ptr(),
                                          self.bp.data_ptr(), self.z.data_ptr(), self.c.data_ptr(),
                                          self.scratch.data_ptr(), self.mmax, 1 if transpose else 0,
                                          int(self.graph_id), stream_ptr())
        if rc:
            raise RuntimeError('block-tridiagonal solve failed: %s' % _lib.lib().last_error())
        x = torch.empty_like(self.z)
        x[self.perm] = self.z
        return x

    def apply(self, r, z):
        z.copy_(self.solve(r))

    def apply_t(self, r, z):
        z.copy_(self.solve(r, transpose=True))
