"""Device-resident Jacobian algebra: the host side of seam S1/S2.

`DevCsr` is a canonical CSR matrix in HBM (int32 indptr/indices, fp64 data)
with a chain of pending scalar factors; `DevEye` is `val * I` (what scipy keeps
as a dia_matrix).  `mul` / `add` reproduce the semantics of the reference's
`_multiply_ops` / `_add_ops` (adstate.py:282-297) -- including scipy's rules
that sums AND products drop exactly-zero results (csr_plus_csr, csr_matmat in
scipy 1.18) while `number * csr` keeps them, and that an empty
product collapses to the int 0 -- by dispatching to the kernels of
csrc/assembly.cu and csrc/coo.cu through the C-ABI.  No arithmetic happens on
the host; torch is used for allocation, streams and index plumbing only.
"""
import numbers

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from ._lib import make_scales, NO_SCALES, MAX_SCALES

INT32_MAX = 2 ** 31 - 1
# rows longer than this leave the row-parallel kernels for the sort-based path
LONG_ROW = 4096

_state = {'device': None}


def device():
    """The CUDA device of this process (cuda:LOCAL_RANK under torchrun).
    Raises when no GPU is present: there is no CPU path."""
    if _state['device'] is None:
        if not torch.cuda.is_available():
            raise RuntimeError('numpad_b200 needs a CUDA device (B200, sm_100a); '
                               'there is no CPU fallback')
        import os
        idx = int(os.environ.get('LOCAL_RANK', '0')) % torch.cuda.device_count()
        torch.cuda.set_device(idx)
        _state['device'] = torch.device('cuda', idx)
    return _state['device']


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _call(name, *args):
    return _lib.lib().call(name, *args, stream_ptr())


def _empty(n, dtype):
    return torch.empty(int(n), dtype=dtype, device=device())


def _scan_ws(n):
    return _empty(_lib.lib().cdll.npb_scan_workspace_elems(int(n)), torch.int64)


def _read_stats(stats):
    total, max_row = stats.tolist()          # the one host sync of a symbolic phase
    if total > INT32_MAX:
        raise OverflowError('Jacobian block with %d nonzeros exceeds the int32 '
                            'index range of one GPU shard' % total)
    return int(total), int(max_row)


# --------------------------------------------------------------------------- #
# Assembly plans: trace once, replay (the reference redoes all index
# bookkeeping on every diff(), adarray.py:470-473, 761-774)
# --------------------------------------------------------------------------- #
# The symbolic half of a two-pass kernel (row counts, scan, output size: one host sync each)
# only depends on the index structure of its operands, and that is the same for every Jacobian
# of a Newton run.  While a plan RECORDS, every such phase stores its result; while it REPLAYS,
# the stored row pointers are handed out instead -- no count kernel, no scan, no sync -- and
# only the numeric kernels run.  Safety does not rest on trust:
#   * every entry carries the shapes / sizes of its operands: a different operation sequence
#     raises PlanMismatch;
#   * index maps that define a structure are compared ON THE DEVICE with the recorded ones;
#   * exact zeros are pruned like scipy does (a value-dependent structure: e.g. the 0/1 mask of
#     an assignment): the recording keeps the PRUNED structure, and a replay fills it with
#     kernels that stay inside every row's recorded slots and count the rows whose number of
#     entries differs; products recorded without zeros count the zeros they produce now.
# The counters are summed once at the end of the assembly (ONE sync); a non-zero total raises
# PlanMismatch and the caller assembles again without a plan.
class PlanMismatch(Exception):
    pass


class AssemblyPlan:
    def __init__(self):
        self.entries = []          # (kind, signature, result, guard tensors)
        self.pos = 0
        self.recording = True
        self.ok = True
        self.checks = []           # device counters of this replay that must all be 0


_active_plan = None
plans = {}                  # (f.size, u.size, mode) -> AssemblyPlan
REPLAY = True               # switch (tests / tuning)
plan_stats = {'recorded': 0, 'replayed': 0, 'mismatch': 0}


class _PlanScope:
    def __init__(self, plan, recording):
        self.plan, self.recording = plan, recording

    def __enter__(self):
        global _active_plan
        self.nested = _active_plan is not None
        if self.nested:
            return None                     # (a lazy Jacobian assembled inside a sweep)
        p = self.plan
        p.recording, p.pos, p.checks = self.recording, 0, []
        _active_plan = p
        return p

    def __exit__(self, et, ev, tb):
        global _active_plan
        if self.nested:
            return False
        p, _active_plan = _active_plan, None
        if et is None and not self.recording:
            if p.pos != len(p.entries):
                raise PlanMismatch('fewer operations than recorded')
            bad = torch.zeros(1, dtype=torch.int64, device=device())
            for c in p.checks:
                bad = bad + c.reshape(-1)[:1].to(torch.int64)
            if int(bad.item()) != 0:        # the one sync of a replayed assembly
                raise PlanMismatch('structure changed (zeros produced or index maps differ)')
        return False


def recording(plan):
    return _PlanScope(plan, True)


def replaying(plan):
    return _PlanScope(plan, False)


def _symbolic(kind, sig, compute, guards=()):
    """result of a symbolic phase: computed (and recorded), or replayed from the active plan.
    guards: index tensors the structure depends on; compared with the recorded ones on the
    device during a replay"""
    p = _active_plan
    if p is None:
        return compute()
    if p.recording:
        out = compute()
        p.entries.append((kind, sig, out, tuple(guards)))
        return out
    if p.pos >= len(p.entries):
        raise PlanMismatch('more operations than recorded')
    k, sg, out, gs = p.entries[p.pos]
    p.pos += 1
    if k != kind or sg != sig or len(gs) != len(guards):
        raise PlanMismatch('%s %r != recorded %s %r' % (kind, sig, k, sg))
    for a, b in zip(gs, guards):
        if a is not b:
            if a.shape != b.shape:
                raise PlanMismatch('guard shape')
            p.checks.append((a != b).sum())
    return out


class DevEye:
    """val * identity(n): the seed of a sweep and everything reached from it
    through scalar blocks only (scipy keeps those as dia_matrix)."""
    __slots__ = ('n', 'val')

    def __init__(self, n, val=1.0):
        self.n, self.val = int(n), float(val)

    shape = property(lambda s: (s.n, s.n))
    nnz = property(lambda s: s.n)

    def scaled(self, a):
        return DevEye(self.n, self.val * a)

    def todev(self):
        """dia.tocsr(): an all-zero diagonal converts to an EMPTY matrix"""
        if self.val == 0.0:
            return DevCsr.empty((self.n, self.n))
        ptr, idx, dat = _empty(self.n + 1, torch.int32), _empty(self.n, torch.int32), \
            _empty(self.n, torch.float64)
        _call('npb_diag_to_csr', self.n, None, self.val, ptr, idx, dat)
        return DevCsr((self.n, self.n), ptr, idx, dat, (), self.n, 1 if self.n else 0)

    def toscipy(self):
        return sp.dia_matrix((np.full((1, self.n), self.val), [0]),
                             shape=(self.n, self.n))


class DevCsr:
    """`clean` = known to hold no explicit zero: true for every sum / product
    result (scipy prunes both), unknown for uploaded or converted blocks."""
    __slots__ = ('shape', 'indptr', 'indices', 'data', 'scales', 'nnz', 'max_row', 'clean')

    def __init__(self, shape, indptr, indices, data, scales, nnz, max_row, clean=False):
        self.shape = (int(shape[0]), int(shape[1]))
        self.indptr, self.indices, self.data = indptr, indices, data
        self.scales = tuple(scales)
        self.nnz, self.max_row = int(nnz), int(max_row)
        self.clean = bool(clean) and all(v != 0.0 for v in self.scales)

    @staticmethod
    def empty(shape):
        return DevCsr(shape, torch.zeros(int(shape[0]) + 1, dtype=torch.int32, device=device()),
                      _empty(0, torch.int32), _empty(0, torch.float64), (), 0, 0, True)

    @staticmethod
    def from_scipy(m):
        m = sp.csr_matrix(m, copy=True)
        m.sum_duplicates()
        m.sort_indices()
        if m.nnz > INT32_MAX:
            raise OverflowError('matrix too large for int32 indices')
        d = device()
        lens = np.diff(m.indptr)
        return DevCsr(m.shape, torch.from_numpy(m.indptr.astype(np.int32)).to(d),
                      torch.from_numpy(m.indices.astype(np.int32)).to(d),
                      torch.from_numpy(m.data.astype(np.float64)).to(d), (), m.nnz,
                      int(lens.max()) if lens.size else 0,
                      clean=bool(np.count_nonzero(m.data) == m.nnz))

    @staticmethod
    def from_coo(shape, rows, cols, vals):
        """device COO triplets (duplicates summed, explicit zeros kept) -> canonical CSR"""
        return coo_to_csr(shape, _as_dev_i32(rows), _as_dev_i32(cols), _as_dev_f64(vals), prune=False)

    def scaled(self, a):
        if len(self.scales) >= MAX_SCALES:
            return self.materialized().scaled(a)
        return DevCsr(self.shape, self.indptr, self.indices, self.data,
                      self.scales + (float(a),), self.nnz, self.max_row, self.clean)

    def materialized(self):
        """apply the pending scalar chain (one pass over data)"""
        if not self.scales:
            return self
        out = _empty(self.nnz, torch.float64)
        _call('npb_csr_apply_scales', self.nnz, self.data, make_scales(self.scales), out)
        return DevCsr(self.shape, self.indptr, self.indices, out, (), self.nnz, self.max_row,
                      self.clean)

    def pruned(self, _plan_aware=True):
        """without its exactly-zero entries (what scipy's product / sum emit)"""
        if self.clean or self.nnz == 0:
            return self
        if any(v == 0.0 for v in self.scales):
            return DevCsr.empty(self.shape)
        p = _active_plan if _plan_aware else None
        if p is not None and not p.recording:
            rec = _symbolic('prune', (self.shape, self.nnz), None)
            if rec is None:                  # nothing was pruned then: verify there is no zero now
                p.checks.append((self.data == 0).sum())
                self.clean = True
                return self
            return self._pruned_into(*rec)
        m = self.shape[0]
        sc = make_scales(self.scales)
        cnt, ptr = _empty(m, torch.int32), _empty(m + 1, torch.int32)
        stats = _empty(2, torch.int64)
        _call('npb_csr_prune_symbolic', m, self.indptr, self.data, sc, cnt, ptr, _scan_ws(m), stats)
        nnz, max_row = _read_stats(stats)
        if nnz == self.nnz:
            self.clean = True
            if p is not None:
                _symbolic('prune', (self.shape, self.nnz), lambda: None)
            return self
        if p is not None:
            _symbolic('prune', (self.shape, self.nnz), lambda: (ptr, nnz, max_row))
        idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
        _call('npb_csr_prune_numeric', m, self.indptr, self.indices, self.data, sc, ptr, idx, dat)
        return DevCsr(self.shape, ptr, idx, dat, (), nnz, max_row, True)

    def _pruned_into(self, ptr, nnz, max_row):
        """prune into a RECORDED structure (assembly plan replay): rows whose non-zero count
        differs from the recording are counted on the device and fail the replay"""
        idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
        bad = _empty(1, torch.int64)
        _call('npb_csr_prune_numeric_checked', self.shape[0], self.indptr, self.indices, self.data,
              make_scales(self.scales), ptr, idx, dat, bad)
        _active_plan.checks.append(bad)
        return DevCsr(self.shape, ptr, idx, dat, (), nnz, max_row, True)

    def todev(self):
        return self

    def toscipy(self):
        m = self.materialized()
        return sp.csr_matrix((m.data.cpu().numpy(), m.indices.cpu().numpy(),
                              m.indptr.cpu().numpy()), shape=self.shape)

    def rows(self):
        """COO row index of every stored entry"""
        r = _empty(self.nnz, torch.int32)
        _call('npb_csr_expand_rows', self.shape[0], self.indptr, r)
        return r

    def transpose(self):
        m = self.materialized()
        t = coo_to_csr((self.shape[1], self.shape[0]), m.indices, m.rows(), m.data,
                       prune=False)
        t.clean = self.clean
        return t

    def matvec(self, x, out=None):
        m = self.materialized()
        if out is None:
            out = _empty(self.shape[0], torch.float64)
        _call('npb_spmv', self.shape[0], self.nnz, m.indptr, m.indices, m.data, x, out)
        return out


# --------------------------------------------------------------------------- #
# local Jacobian blocks emitted by the tape (public names of adstate.py:233-272)
# --------------------------------------------------------------------------- #
def _as_dev_f64(a):
    if isinstance(a, torch.Tensor):
        t = a.to(device=device(), dtype=torch.float64)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64))).to(device())
    return t.reshape(-1).contiguous() if not (t.dim() == 1 and t.is_contiguous()) else t


def _as_dev_i32(a):
    if isinstance(a, torch.Tensor):
        t = a.to(device=device(), dtype=torch.int32)
    else:
        a = np.asarray(a)
        if a.size and (a.max() > INT32_MAX or a.min() < -INT32_MAX):
            raise OverflowError('index out of int32 range')
        t = torch.from_numpy(np.ascontiguousarray(a.astype(np.int32))).to(device())
    return t.reshape(-1).contiguous()


class dia_jac:
    """A Jacobian represented by a diagonal matrix (reference adstate.py:233-250).
    `diag_entries` may be a numpy array or a device tensor; it lives in HBM."""

    def __init__(self, diag_entries):
        self.data = _as_dev_f64(diag_entries)

    @property
    def shape(self):
        return (self.data.numel(), self.data.numel())

    def todev(self):
        n = self.data.numel()
        ptr, idx, dat = _empty(n + 1, torch.int32), _empty(n, torch.int32), _empty(n, torch.float64)
        _call('npb_diag_to_csr', n, self.data, 1.0, ptr, idx, dat)
        return DevCsr((n, n), ptr, idx, dat, (), n, 1 if n else 0)

    def tocsr(self):
        return self.todev().toscipy()


class csr_jac:
    """Jacobian given as COO triplets (reference adstate.py:252-272): general
    block, converted on the device (sort + duplicate sum) on first use."""

    def __init__(self, data, i, j, shape=None):
        self.data, self.i, self.j = data, i, j
        self._shape = None if shape is None else (int(shape[0]), int(shape[1]))
        self._dev = None

    @property
    def shape(self):
        if self._shape is None:
            i, j = _as_dev_i32(self.i), _as_dev_i32(self.j)
            self._shape = (int(i.max()) + 1, int(j.max()) + 1)
        return self._shape

    def todev(self):
        if self._dev is None:
            self._dev = coo_to_csr(self.shape, _as_dev_i32(self.i), _as_dev_i32(self.j),
                                   _as_dev_f64(self.data), prune=False)
        return self._dev

    def tocsr(self):
        return self.todev().toscipy()


class sel_jac:
    """Selection block with at most one entry per row:
    row r has the entry (r, map[r]) with value w[r] (1 if w is None), or is
    empty when map[r] < 0.  Emitted by indexing, assignment, stacking,
    broadcasting (what the reference builds as csr_jac with unit data,
    adarray.py:761-796, 356-384, 593-607, 670-693).

    monotone : map is strictly increasing over its non-negative entries
               (a product with it keeps rows sorted and duplicate-free)
    full     : no row is empty
    """

    def __init__(self, map_, n_cols, weights=None, monotone=None, full=None):
        self.map = map_
        self.n_rows, self.n_cols = int(map_.numel()), int(n_cols)
        self.w = weights
        self._monotone, self._full = monotone, full
        self._dev = None

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    @property
    def full(self):
        if self._full is None:
            self._full = bool((self.map >= 0).all().item())
        return self._full

    @property
    def monotone(self):
        if self._monotone is None:
            kept = self.map[self.map >= 0]
            self._monotone = bool((kept[1:] > kept[:-1]).all().item()) if kept.numel() > 1 else True
        return self._monotone

    def todev(self):
        if self._dev is None:
            n = self.n_rows

            def symbolic():
                cnt, ptr = _empty(n, torch.int32), _empty(n + 1, torch.int32)
                stats = _empty(2, torch.int64)
                _call('npb_sel_to_csr_symbolic', n, self.map, cnt, ptr, _scan_ws(n), stats)
                return ptr, _read_stats(stats)[0]
            ptr, nnz = _symbolic('sel_to_csr', (n, self.n_cols), symbolic, (self.map,))
            idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
            _call('npb_sel_to_csr_numeric', n, self.map, self.w, ptr, idx, dat)
            # unit weights: no stored zero, nothing for pruned() to look for
            self._dev = DevCsr(self.shape, ptr, idx, dat, (), nnz, 1 if nnz else 0,
                               clean=self.w is None)
        return self._dev

    def tocsr(self):
        return self.todev().toscipy()


def tocsr(jac):
    """reference adstate.py:274-278 -- host (scipy) view of a block"""
    if isinstance(jac, (dia_jac, csr_jac, sel_jac)):
        return jac.tocsr()
    if isinstance(jac, (DevCsr, DevEye)):
        return jac.toscipy()
    return jac


# --------------------------------------------------------------------------- #
# kernels behind the products / sums
# --------------------------------------------------------------------------- #
def coo_to_csr(shape, rows, cols, vals, prune):
    nrows, ncols, nnz_in = int(shape[0]), int(shape[1]), int(rows.numel())
    L = _lib.lib()
    ws_bytes = L.cdll.npb_coo_to_csr_workspace(nrows, nnz_in)
    ws = _empty(ws_bytes, torch.uint8)
    ptr = _empty(nrows + 1, torch.int32)
    stats = _empty(2, torch.int64)
    _lib.lib().call('npb_coo_to_csr_symbolic', nrows, ncols, nnz_in, rows, cols, vals,
                    1 if prune else 0, ws, ws_bytes, ptr, stats, stream_ptr())
    nnz, max_row = _read_stats(stats)
    idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
    _call('npb_coo_to_csr_numeric', nrows, ncols, nnz_in, ws, idx, dat)
    return DevCsr(shape, ptr, idx, dat, (), nnz, max_row, clean=bool(prune))


def _zero_cell():
    return _empty(1, torch.int64)


def _finish(C, zc):
    """C was written by a product kernel that counted the zeros it produced in
    `zc`; compact them away if there are any (csr_matmat never stores a zero)"""
    p = _active_plan
    if zc is not None and p is not None and not p.recording:
        # replay: what the recording saw decides, the counters verify it at the end
        pruned = _symbolic('finish', (C.shape, C.nnz), None)
        if pruned is None:
            p.checks.append(zc)              # no zero then: there must be none now
            C.clean = True
            return C
        return C._pruned_into(*pruned)
    if zc is not None and int(zc.item()) > 0:
        C.clean = False
        out = C.pruned(_plan_aware=False)
        if p is not None:
            _symbolic('finish', (C.shape, C.nnz), lambda: (out.indptr, out.nnz, out.max_row))
        return out
    if zc is not None and p is not None:
        _symbolic('finish', (C.shape, C.nnz), lambda: None)
    C.clean = True
    return C


def _csr_times_diag(A, d):
    out, zc = _empty(A.nnz, torch.float64), _zero_cell()
    _call('npb_csr_scale_cols', A.nnz, A.indices, A.data, make_scales(A.scales), d.data, out, zc)
    return _finish(DevCsr(A.shape, A.indptr, A.indices, out, (), A.nnz, A.max_row), zc)


def _diag_times_csr(d, B):
    out, zc = _empty(B.nnz, torch.float64), _zero_cell()
    _call('npb_csr_scale_rows', B.shape[0], B.indptr, B.data, make_scales(B.scales), d.data,
          out, zc)
    return _finish(DevCsr(B.shape, B.indptr, B.indices, out, (), B.nnz, B.max_row), zc)


def _csr_times_sel(A, S):
    """A (m x k) times a selection (k x n): column relabel."""
    shape = (A.shape[0], S.n_cols)
    if A.nnz == 0:
        return DevCsr.empty(shape)
    if S.monotone:
        sc = make_scales(A.scales)
        if S.full:              # per-entry kernel, any row length
            idx = _empty(A.nnz, torch.int32)
            if S.w is None:     # data buffer and pending scales are shared
                _call('npb_csr_relabel_same', A.nnz, A.indices, A.data, sc, S.map, None, idx,
                      None, None)
                return DevCsr(shape, A.indptr, idx, A.data, A.scales, A.nnz, A.max_row, True)
            dat, zc = _empty(A.nnz, torch.float64), _zero_cell()
            _call('npb_csr_relabel_same', A.nnz, A.indices, A.data, sc, S.map, S.w, idx, dat, zc)
            return _finish(DevCsr(shape, A.indptr, idx, dat, (), A.nnz, A.max_row), zc)
        if A.max_row <= LONG_ROW:
            m = A.shape[0]

            def symbolic():
                cnt, ptr = _empty(m, torch.int32), _empty(m + 1, torch.int32)
                stats = _empty(2, torch.int64)
                _call('npb_csr_relabel_symbolic', m, A.indptr, A.indices, S.map, cnt, ptr,
                      _scan_ws(m), stats)
                return (ptr,) + _read_stats(stats)
            ptr, nnz, max_row = _symbolic('relabel', (A.shape, A.nnz, S.n_cols), symbolic, (S.map,))
            # (slack: in a replay whose operand deviated from the recording -- detected at the
            # end -- a row may hold more entries than its recorded slots)
            slack = A.max_row + 8 if _active_plan is not None else 0
            idx, dat = _empty(nnz + slack, torch.int32)[:nnz], _empty(nnz + slack, torch.float64)[:nnz]
            zc = _zero_cell() if S.w is not None else None
            _call('npb_csr_relabel_numeric', m, A.indptr, A.indices, A.data, sc, S.map, S.w,
                  ptr, idx, dat, zc)
            return _finish(DevCsr(shape, ptr, idx, dat, (), nnz, max_row), zc)
    if A.max_row <= 48 and USE_SPGEMM_ROWS:
        # non-monotone map, short rows: relabel + per-row sort + duplicate sums, no device-wide sort
        m = A.shape[0]
        tidx, tval = _empty(A.nnz, torch.int32), _empty(A.nnz, torch.float64)
        cnt, tstart = _empty(m, torch.int32), _empty(m, torch.int32)
        _call('npb_csr_relabel_merge_rows', m, A.indptr, A.indices, A.data, make_scales(A.scales),
              S.map, S.w, 1, tidx, tval, cnt, tstart)
        ptr = _empty(m + 1, torch.int32)
        stats = _empty(2, torch.int64)
        _call('npb_spgemm_rows_indptr', m, cnt, ptr, _scan_ws(m), stats)
        nnz, max_row = _read_stats(stats)
        idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
        _call('npb_spgemm_rows_compact', m, cnt, tstart, ptr, tidx, tval, idx, dat)
        return DevCsr(shape, ptr, idx, dat, (), nnz, max_row, True)
    # general: relabel in COO form, drop, sort, merge duplicates, prune
    Am = A.materialized()
    cols = S.map[Am.indices.long()]
    vals = Am.data if S.w is None else _mul_entries(Am.data, S.w, Am.indices)
    rows = Am.rows()
    if not S.full:
        keep = cols >= 0
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    return coo_to_csr(shape, rows.contiguous(), cols.contiguous(), vals.contiguous(), prune=True)


def _mul_entries(data, w, indices):
    """data[k] * w[indices[k]] through the column-scaling kernel"""
    out = _empty(data.numel(), torch.float64)
    _call('npb_csr_scale_cols', data.numel(), indices, data, NO_SCALES, w, out, None)
    return out


def _sel_times_csr(S, B):
    """selection (m x k) times B (k x n): row gather."""
    m = S.n_rows
    shape = (m, B.shape[1])
    def symbolic():
        cnt, ptr = _empty(m, torch.int32), _empty(m + 1, torch.int32)
        stats = _empty(2, torch.int64)
        _call('npb_csr_gather_rows_symbolic', m, S.map, B.indptr, cnt, ptr, _scan_ws(m), stats)
        return (ptr,) + _read_stats(stats)
    ptr, nnz, max_row = _symbolic('gather_rows', (m, B.shape, B.nnz), symbolic, (S.map,))
    slack = B.max_row + 8 if _active_plan is not None else 0
    idx, dat = _empty(nnz + slack, torch.int32)[:nnz], _empty(nnz + slack, torch.float64)[:nnz]
    zc = _zero_cell() if S.w is not None else None
    _call('npb_csr_gather_rows_numeric', m, S.map, S.w, B.indptr, B.indices, B.data,
          make_scales(B.scales), ptr, idx, dat, zc)
    return _finish(DevCsr(shape, ptr, idx, dat, (), nnz, max_row), zc)


# expansion buffers of the general product are capped at this many triplets;
# larger products are done by row blocks of A and stacked
SPGEMM_CHUNK = 1 << 29


def _row_block(A, r0, r1):
    """rows r0..r1 of a CSR matrix as its own DevCsr (views + rebased indptr)"""
    k0, k1 = int(A.indptr[r0].item()), int(A.indptr[r1].item())
    ptr = (A.indptr[r0:r1 + 1] - k0).contiguous()
    return DevCsr((r1 - r0, A.shape[1]), ptr, A.indices[k0:k1], A.data[k0:k1], A.scales,
                  k1 - k0, A.max_row, A.clean)


def _vstack(blocks, shape):
    ptrs, off = [], 0
    for b in blocks:
        ptrs.append(b.indptr[:-1] + off)
        off += b.nnz
    ptrs.append(torch.tensor([off], dtype=torch.int32, device=device()))
    bm = [b.materialized() for b in blocks]
    return DevCsr(shape, torch.cat(ptrs).to(torch.int32), torch.cat([b.indices for b in bm]),
                  torch.cat([b.data for b in bm]), (), off, max(b.max_row for b in blocks), True)


USE_SPGEMM_ROWS = True
# (rank, world, min_rows) while shard.solve builds its hierarchies: the large general products
# are computed by row blocks, one per rank, and all-gathered (bit-identical to the local product)
SPGEMM_SHARD = None


def _csr_times_csr_sharded(A, B):
    """rows [m q / world, m (q+1) / world) of A B on rank q, then an all-gather of the pieces;
    None when a row has too many partial products on any rank (all ranks take the same path)"""
    import torch.distributed as dist
    rank, world, _ = SPGEMM_SHARD
    m = A.shape[0]
    cuts = [m * q // world for q in range(world + 1)]
    A = A.materialized()
    C = _csr_times_csr_rows(_row_block(A, cuts[rank], cuts[rank + 1]), B)
    meta = torch.tensor([C.nnz if C is not None else -1, C.max_row if C is not None else 0],
                        dtype=torch.int64, device=device())
    allm = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(allm, meta)
    allm = torch.stack(allm).tolist()
    if any(a[0] < 0 for a in allm):
        return None
    nnzs = [int(a[0]) for a in allm]
    koff = [0]
    for v in nnzs:
        koff.append(koff[-1] + v)
    idx, dat = _empty(koff[-1], torch.int32), _empty(koff[-1], torch.float64)
    cnt = _empty(m, torch.int32)
    idx[koff[rank]:koff[rank + 1]] = C.indices
    dat[koff[rank]:koff[rank + 1]] = C.data
    cnt[cuts[rank]:cuts[rank + 1]] = C.indptr[1:] - C.indptr[:-1]
    for q in range(world):
        if cuts[q + 1] > cuts[q]:
            dist.broadcast(cnt[cuts[q]:cuts[q + 1]], q)
        if nnzs[q]:
            dist.broadcast(idx[koff[q]:koff[q + 1]], q)
            dist.broadcast(dat[koff[q]:koff[q + 1]], q)
    ptr = torch.zeros(m + 1, dtype=torch.int32, device=device())
    ptr[1:] = torch.cumsum(cnt, 0)
    return DevCsr((m, B.shape[1]), ptr, idx, dat, (), koff[-1], max(int(a[1]) for a in allm), True)


def _csr_times_csr_rows(A, B):
    """general product, row-batched with shared-memory staging (csrc/spgemm.cu); None when a
    row has too many partial products for a batch (the caller takes the sort-based path).
    Same arithmetic and summation order as the sort-based path."""
    m = A.shape[0]
    L = _lib.lib().cdll
    ub, ubptr = _empty(m, torch.int32), _empty(m + 1, torch.int32)
    stats = _empty(2, torch.int64)
    _call('npb_spgemm_rows_bound', m, A.indptr, A.indices, B.indptr, ub, ubptr, _scan_ws(m), stats)
    total, max_ub = _read_stats(stats)
    if max_ub > L.npb_spgemm_rows_cap() // 2 or total > INT32_MAX - 8:
        return None
    nb = L.npb_spgemm_rows_batches(total, max_ub)
    brow = _empty(nb + 1, torch.int32)
    tidx, tval = _empty(total, torch.int32), _empty(total, torch.float64)
    cnt, tstart = _empty(m, torch.int32), _empty(m, torch.int32)
    _call('npb_spgemm_rows_numeric', m, total, max_ub, A.indptr, A.indices, A.data,
          make_scales(A.scales), B.indptr, B.indices, B.data, make_scales(B.scales), ubptr, 1,
          brow, tidx, tval, cnt, tstart)
    ptr = _empty(m + 1, torch.int32)
    st2 = _empty(2, torch.int64)
    _call('npb_spgemm_rows_indptr', m, cnt, ptr, _scan_ws(m), st2)
    nnz, max_row = _read_stats(st2)
    idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
    _call('npb_spgemm_rows_compact', m, cnt, tstart, ptr, tidx, tval, idx, dat)
    return DevCsr((A.shape[0], B.shape[1]), ptr, idx, dat, (), nnz, max_row, True)


def _csr_times_csr(A, B):
    """general product (scipy csr_matmat semantics: duplicates summed, exact-zero sums
    dropped): row-batched in shared memory when every row fits, else expansion + sort"""
    shape = (A.shape[0], B.shape[1])
    if A.nnz == 0 or B.nnz == 0:
        return DevCsr.empty(shape)
    if USE_SPGEMM_ROWS:
        if SPGEMM_SHARD is not None and A.shape[0] >= SPGEMM_SHARD[2] and not A.scales and not B.scales:
            C = _csr_times_csr_sharded(A, B)
        else:
            C = _csr_times_csr_rows(A, B)
        if C is not None:
            return C
    cnt, off = _empty(A.nnz, torch.int32), _empty(A.nnz + 1, torch.int32)
    stats = _empty(2, torch.int64)
    _call('npb_spgemm_expand_symbolic', A.nnz, A.indices, B.indptr, cnt, off,
          _scan_ws(A.nnz), stats)
    total = int(stats[0].item())
    if total == 0:
        return DevCsr.empty(shape)
    if total > SPGEMM_CHUNK and A.shape[0] > 1:
        nblk = min(A.shape[0], -(-total // (SPGEMM_CHUNK // 2)))
        cuts = [int(round(i * A.shape[0] / nblk)) for i in range(nblk + 1)]
        return _vstack([_csr_times_csr(_row_block(A, cuts[i], cuts[i + 1]), B)
                        for i in range(nblk) if cuts[i + 1] > cuts[i]], shape)
    if total > INT32_MAX:
        raise OverflowError('sparse product with %d partial products in one row block' % total)
    r, c, v = _empty(total, torch.int32), _empty(total, torch.int32), _empty(total, torch.float64)
    _call('npb_spgemm_expand_numeric', A.nnz, A.rows(), A.indices, A.data, make_scales(A.scales),
          B.indptr, B.indices, B.data, make_scales(B.scales), off, r, c, v)
    return coo_to_csr(shape, r, c, v, prune=True)


def _csr_plus_csr(A, B):
    assert A.shape == B.shape
    m = A.shape[0]
    if max(A.max_row, B.max_row) > LONG_ROW:
        Am, Bm = A.materialized(), B.materialized()
        return coo_to_csr(A.shape, torch.cat([Am.rows(), Bm.rows()]),
                          torch.cat([Am.indices, Bm.indices]),
                          torch.cat([Am.data, Bm.data]), prune=True)
    sa, sb = make_scales(A.scales), make_scales(B.scales)

    def symbolic(prune=1):
        cnt, ptr = _empty(m, torch.int32), _empty(m + 1, torch.int32)
        stats = _empty(2, torch.int64)
        _call('npb_spadd_symbolic', m, A.indptr, A.indices, A.data, sa, B.indptr, B.indices,
              B.data, sb, prune, cnt, ptr, _scan_ws(m), stats)
        return (ptr,) + _read_stats(stats)
    p = _active_plan
    if p is not None and not p.recording:
        # replay: fill the recorded (pruned) structure; a row whose entry count differs -- a sum
        # that cancels exactly now but did not then, or the reverse -- fails the replay
        ptr, nnz, max_row = _symbolic('spadd', (A.shape, A.nnz, B.nnz), None)
        idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
        bad = _empty(1, torch.int64)
        _call('npb_spadd_numeric_checked', m, A.indptr, A.indices, A.data, sa, B.indptr, B.indices,
              B.data, sb, 1, ptr, idx, dat, bad)
        p.checks.append(bad)
        return DevCsr(A.shape, ptr, idx, dat, (), nnz, max_row, True)
    ptr, nnz, max_row = _symbolic('spadd', (A.shape, A.nnz, B.nnz), symbolic)
    idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
    _call('npb_spadd_numeric', m, A.indptr, A.indices, A.data, sa, B.indptr, B.indices,
          B.data, sb, 1, ptr, idx, dat)
    return DevCsr(A.shape, ptr, idx, dat, (), nnz, max_row, True)


# --------------------------------------------------------------------------- #
# _multiply_ops / _add_ops                          (reference adstate.py:282-297)
# --------------------------------------------------------------------------- #
def _is0(x):
    """the reference's `x is 0`: the int literal only"""
    return type(x) is int and x == 0


def _as_operand(x):
    """number | DevEye | DevCsr | dia_jac | sel_jac | csr_jac ; foreign sparse
    matrices (user supplied scipy blocks) are uploaded once"""
    if isinstance(x, (numbers.Number, DevEye, DevCsr, dia_jac, sel_jac, csr_jac)):
        return x
    if sp.issparse(x):
        return DevCsr.from_scipy(x)
    if hasattr(x, 'tocsr'):
        return DevCsr.from_scipy(x.tocsr())
    raise TypeError('unsupported Jacobian block %r' % type(x))


def _mat(x):
    """accumulated derivative (DevEye / DevCsr) -> DevCsr"""
    return x.todev()


def _collapse(c):
    """`if op.nnz == 0: op = 0`"""
    if isinstance(c, DevCsr) and c.nnz == 0:
        return 0
    return c


def _multiply_ops(op0, op1):
    if _is0(op0) or _is0(op1):
        return 0
    a, b = _as_operand(op0), _as_operand(op1)
    an, bn = isinstance(a, numbers.Number), isinstance(b, numbers.Number)
    if an and bn:
        return a * b
    if an or bn:
        s, m = (a, b) if an else (b, a)
        if isinstance(m, (DevEye, DevCsr)):
            return _collapse(m.scaled(s))
        return _collapse(m.todev().scaled(s))
    # matrix x matrix: csr_matmat semantics -- operands' explicit zeros never
    # reach the result, so accumulated operands are cleaned first
    if isinstance(a, DevEye) and isinstance(b, DevEye):
        return DevEye(a.n, a.val * b.val)
    if isinstance(a, DevEye):           # val*I x block
        if a.val == 0.0:                # an all-zero dia converts to an empty CSR
            return 0
        c = (b if isinstance(b, DevCsr) else b.todev()).pruned()
        return _collapse(c if a.val == 1.0 else c.scaled(a.val))
    if isinstance(b, DevEye):
        if b.val == 0.0:
            return 0
        c = (a if isinstance(a, DevCsr) else a.todev()).pruned()
        return _collapse(c if b.val == 1.0 else c.scaled(b.val))
    if isinstance(a, DevCsr):           # adjoint order: accumulated x local block
        a = a.pruned()
        if a.nnz == 0:
            return 0
        if isinstance(b, dia_jac):
            return _collapse(_csr_times_diag(a, b))
        if isinstance(b, sel_jac):
            return _collapse(_csr_times_sel(a, b))
        return _collapse(_csr_times_csr(a, b.todev()))
    if isinstance(b, DevCsr):           # tangent order: local block x accumulated
        b = b.pruned()
        if b.nnz == 0:
            return 0
        if isinstance(a, dia_jac):
            return _collapse(_diag_times_csr(a, b))
        if isinstance(a, sel_jac):
            return _collapse(_sel_times_csr(a, b))
        return _collapse(_csr_times_csr(a.todev(), b))
    return _collapse(_csr_times_csr(a.todev(), b.todev()))


def _add_ops(op0, op1):
    if _is0(op0):
        return op1
    if _is0(op1):
        return op0
    a, b = _as_operand(op0), _as_operand(op1)
    if isinstance(a, numbers.Number) or isinstance(b, numbers.Number):
        raise TypeError('cannot add a number to a Jacobian')
    if isinstance(a, DevEye) and isinstance(b, DevEye):
        return DevEye(a.n, a.val + b.val)        # dia + dia keeps zeros
    return _csr_plus_csr(a.todev(), b.todev())
