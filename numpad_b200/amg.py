"""Setup of the multigrid hierarchies and of the least-squares-commutator block
preconditioner whose solve phase lives in csrc/amg.cu.

Everything numeric runs on the device through the C-ABI: the strength graph /
MIS aggregation (npb_amg_aggregate), the prolongation smoothing and the
Galerkin products R A P (the same SpGEMM / SpAdd kernels that accumulate
Jacobians), the dense inverse of the coarsest operator (npb_dense_inverse).
Python only owns the buffers and fills the descriptor structs.
"""
import ctypes

import torch

from . import _lib
from ._dev import (DevCsr, sel_jac, dia_jac, device, stream_ptr, _empty, _call, _scan_ws,
                   _read_stats, _csr_times_sel, _csr_times_csr, _diag_times_csr,
                   _csr_plus_csr, _sel_times_csr, coo_to_csr)


class Halo(ctypes.Structure):            # npb_halo: filled by shard.py for row-sharded operators
    _fields_ = [('dist', ctypes.c_void_p), ('n_own', ctypes.c_int64),
                ('send_idx', ctypes.c_void_p), ('n_send', ctypes.c_int32),
                ('maxb', ctypes.c_int32), ('recv', ctypes.c_void_p),
                ('p2p_data', ctypes.c_int64), ('p2p_flags', ctypes.c_int64),
                ('seq', ctypes.c_void_p)]


class CsrMat(ctypes.Structure):
    _fields_ = [('nrows', ctypes.c_int64), ('ncols', ctypes.c_int64), ('nnz', ctypes.c_int64),
                ('indptr', ctypes.c_void_p), ('indices', ctypes.c_void_p),
                ('data', ctypes.c_void_p), ('halo', Halo), ('data32', ctypes.c_void_p),
                ('dia_vals', ctypes.c_void_p), ('ndiag', ctypes.c_int32), ('dia_ld', ctypes.c_int32),
                ('dia_off', ctypes.c_int32 * 8),
                ('ell_idx', ctypes.c_void_p), ('ell_val', ctypes.c_void_p),
                ('ell_w', ctypes.c_int32), ('pad2', ctypes.c_int32)]


class AmgLevel(ctypes.Structure):
    _fields_ = [('A', CsrMat), ('P', CsrMat), ('R', CsrMat), ('dinv', ctypes.c_void_p),
                ('x', ctypes.c_void_p), ('x2', ctypes.c_void_p), ('b', ctypes.c_void_p),
                ('r', ctypes.c_void_p), ('gather', ctypes.c_int32), ('pad', ctypes.c_int32),
                ('offsets', ctypes.c_void_p), ('g_data', ctypes.c_int64),
                ('g_flags', ctypes.c_int64), ('g_seq', ctypes.c_void_p)]


class AmgStruct(ctypes.Structure):
    _fields_ = [('nlevels', ctypes.c_int32), ('nu_pre', ctypes.c_int32),
                ('nu_post', ctypes.c_int32), ('coarse_n', ctypes.c_int32),
                ('omega', ctypes.c_double), ('coarse_inv', ctypes.c_void_p),
                ('levels', ctypes.POINTER(AmgLevel)), ('dist', ctypes.c_void_p),
                ('gamma', ctypes.c_int32), ('gamma_from', ctypes.c_int32)]


class AmgPrecondStruct(ctypes.Structure):
    _fields_ = [('amg', ctypes.c_void_p), ('A', CsrMat), ('cycles', ctypes.c_int32),
                ('pad', ctypes.c_int32), ('r', ctypes.c_void_p), ('dx', ctypes.c_void_p)]


class LscStruct(ctypes.Structure):
    _fields_ = [('n', ctypes.c_int64), ('nv', ctypes.c_int64), ('np', ctypes.c_int64),
                ('vset', ctypes.c_void_p), ('pset', ctypes.c_void_p),
                ('A', CsrMat), ('G', CsrMat), ('D', CsrMat), ('Lm', CsrMat),
                ('a_dinv', ctypes.c_void_p), ('amg_a', ctypes.c_void_p),
                ('amg_l', ctypes.c_void_p), ('cycles_a', ctypes.c_int32),
                ('cycles_l', ctypes.c_int32), ('wv', ctypes.c_void_p * 5),
                ('wp', ctypes.c_void_p * 8), ('scal', ctypes.c_void_p),
                ('red', ctypes.c_void_p), ('a_gmres', ctypes.c_int32), ('graph_id', ctypes.c_int32),
                ('a_work', ctypes.c_void_p), ('a_work_bytes', ctypes.c_int64)]


# Operators of the preconditioner with at least this many rows are ALSO stored with fp32
# values and applied from those (csrc/spmv_tma.cuh, value type float): the preconditioner is
# an approximation anyway (a few multigrid cycles inside flexible GMRES, whose Krylov basis,
# operator and residuals stay fp64), and its SpMVs are HBM-bound at 12 bytes per entry, 8 of
# them the value.  0 disables.
F32_MIN_ROWS = 200000
_f32_cache = {}


def f32_values(A):
    """fp32 copy of A.data for large operators (None otherwise); cached per data tensor so that
    the LSC blocks and the finest multigrid level share one copy"""
    if not F32_MIN_ROWS or A.shape[0] < F32_MIN_ROWS or A.nnz == 0 or A.nnz / A.shape[0] > 96:
        return None
    key = (A.data.data_ptr(), A.nnz)
    hit = _f32_cache.get(key)
    if hit is not None and hit[0]() is A.data:
        return hit[1]
    import weakref
    v = A.data.to(torch.float32)
    if len(_f32_cache) > 64:
        _f32_cache.clear()
    _f32_cache[key] = (weakref.ref(A.data), v)
    return v


# Stencil operators whose entries sit on at most 8 constant offsets col - row (the pressure
# Laplacian L = -D G of the cavity: 5) are also kept in diagonal storage (fp32, no column
# indices, coalesced reads of x instead of gathers): csrc/krylov.cu k_spmv_dia.  0 disables.
DIA_MIN_ROWS = 200000
# also build hybrid (diagonals + ELL remainder) images, see dia_image.  OFF: measured on the momentum
# block of the 2048^2 cavity (tools/hyb_bench.py), the hybrid kernel takes 212 us per Jacobi sweep
# against 169 us of the TMA-pipelined fp32 CSR kernel (which already runs at 0.88 of the HBM peak
# there): fewer gathers, but two dependent load phases per row at 2 CTAs / SM
HYB_IMAGES = False
_dia_cache = {}


def dia_image(A):
    """-> (offsets list, vals float32[ndiag * ld], ld, ell) or None; ell = None for a pure
    diagonal-storage image, else (width, idx int32[width * ld], val float32[width * ld]) -- the
    hybrid image of an operator MOST of whose entries sit on a few diagonals (the momentum block
    of the cavity: 5 of 9 per row on 7 diagonals, 4 cross-coupling entries with drifting offsets).
    The candidate offsets come from a strided sample of the entries; the conversion kernel checks
    EVERY entry against them."""
    n = A.shape[0]
    if (not DIA_MIN_ROWS or n < DIA_MIN_ROWS or A.shape[0] != A.shape[1] or A.nnz == 0
            or A.nnz > 16 * n):
        return None
    key = (A.data.data_ptr(), A.nnz)
    hit = _dia_cache.get(key)
    if hit is not None and hit[0]() is A.data:
        return hit[1]
    import weakref
    step = max(1, A.nnz // 200000)
    d = (A.indices[::step].to(torch.int64) - A.rows()[::step].to(torch.int64))
    offs, cnt = torch.unique(d, return_counts=True)
    out = None
    csr_bytes = 8.0 * A.nnz + 4.0 * n
    bad = _empty(1, torch.int64)
    if 1 <= offs.numel() <= 8 and 4.0 * offs.numel() * n < 0.8 * csr_bytes:
        off_list = [int(v) for v in offs.tolist()]
        off_h = (ctypes.c_int32 * 8)(*(off_list + [0] * (8 - len(off_list))))
        ld = (n + 31) // 32 * 32                     # every diagonal 128-byte aligned
        vals = _empty(len(off_list) * ld, torch.float32)
        _call('npb_csr_to_dia', n, A.indptr, A.indices, A.data, len(off_list),
              ctypes.cast(off_h, ctypes.c_void_p).value, vals, ld, bad)
        if int(bad.item()) == 0:
            out = (off_list, vals, ld, None)
    elif HYB_IMAGES and offs.numel() > 8:
        # diagonals that hold an entry in at least a quarter of the rows (as the sample sees it)
        rows_in_sample = d.numel() * n / max(A.nnz, 1)
        strong = offs[cnt.to(torch.float64) >= 0.25 * rows_in_sample]
        if 1 <= strong.numel() <= 8:
            off_list = [int(v) for v in strong.tolist()]
            nd = len(off_list)
            off_h = (ctypes.c_int32 * 8)(*(off_list + [0] * (8 - nd)))
            offp = ctypes.cast(off_h, ctypes.c_void_p).value
            maxw = _empty(1, torch.int32)
            _call('npb_csr_dia_remainder', n, A.indptr, A.indices, nd, offp, maxw)
            w = int(maxw.item())
            if 1 <= w <= 8 and (4.0 * nd + 8.0 * w) * n < 0.85 * csr_bytes:
                ld = (n + 31) // 32 * 32
                vals = _empty(nd * ld, torch.float32)
                eidx, eval_ = _empty(w * ld, torch.int32), _empty(w * ld, torch.float32)
                _call('npb_csr_to_hyb', n, A.shape[1], A.indptr, A.indices, A.data, nd, offp, vals, ld,
                      w, eidx, eval_, bad)
                if int(bad.item()) == 0:
                    out = (off_list, vals, ld, (w, eidx, eval_))
    if len(_dia_cache) > 16:
        _dia_cache.clear()
    _dia_cache[key] = (weakref.ref(A.data), out)
    return out


def csr_mat(A, keep=None):
    """descriptor of a device CSR matrix; with `keep` (a list that outlives the descriptor) large
    operators also get their fp32 values (see F32_MIN_ROWS) or, stencil operators, their
    diagonal-storage image (see DIA_MIN_ROWS)"""
    m = CsrMat()
    m.nrows, m.ncols, m.nnz = A.shape[0], A.shape[1], A.nnz
    m.indptr, m.indices, m.data = A.indptr.data_ptr(), A.indices.data_ptr(), A.data.data_ptr()
    if keep is not None:
        d = dia_image(A)
        if d is not None:
            offs, vals, ld, ell = d
            keep.append(vals)
            m.dia_vals, m.ndiag, m.dia_ld = vals.data_ptr(), len(offs), ld
            for k, o in enumerate(offs):
                m.dia_off[k] = o
            if ell is not None:
                keep += [ell[1], ell[2]]
                m.ell_w, m.ell_idx, m.ell_val = ell[0], ell[1].data_ptr(), ell[2].data_ptr()
        v = f32_values(A) if d is None else None
        if v is not None:
            keep.append(v)
            m.data32 = v.data_ptr()
    return m


def diagonal(A):
    d = _empty(A.shape[0], torch.float64)
    _call('npb_csr_diagonal', A.shape[0], A.indptr, A.indices, A.data, d)
    return d


def aggregate(A, theta, max_rounds=12, kind='mis1', with_roots=False):
    """-> (agg int32[n], n_agg).  kind 'mis1': roots = maximal independent set of the
    strength graph (aggregates of ~2.75 nodes on a 5-point stencil); 'mis2': distance-2
    independent set (~7 nodes)"""
    n = A.shape[0]
    if kind == 'mis2':
        state, near = _empty(n, torch.int8), _empty(n, torch.int8)
        keys = _empty(2 * n, torch.int64)
        label, label2, flag = (_empty(n, torch.int32) for _ in range(3))
        root_id, agg = _empty(n + 1, torch.int32), _empty(n, torch.int32)
        und = _empty(2, torch.int32)
        stats = _empty(2, torch.int64)
        _call('npb_amg_aggregate_mis2', n, A.indptr, A.indices, A.data, float(theta),
              int(max(max_rounds, 40)), state, near, keys, label, label2, flag, root_id, agg, und,
              _scan_ws(n), stats)
        n_agg, _ = _read_stats(stats)
        return (agg, n_agg, root_id) if with_roots else (agg, n_agg)
    state, state2 = _empty(n, torch.int8), _empty(n, torch.int8)
    flag, root_id, agg = _empty(n, torch.int32), _empty(n + 1, torch.int32), _empty(n, torch.int32)
    und = _empty(2, torch.int32)
    stats = _empty(2, torch.int64)
    _call('npb_amg_aggregate', n, A.indptr, A.indices, A.data, float(theta), int(max_rounds),
          state, state2, flag, root_id, agg, und, _scan_ws(n), stats)
    n_agg, _ = _read_stats(stats)
    # root_id[i] = number of roots below node i: aggregate ids follow the order of their roots
    return (agg, n_agg, root_id) if with_roots else (agg, n_agg)


def merge_small_aggregates(indptr, indices, data, agg, n_agg, root_id, max_small=2, max_size=6):
    """Merge the aggregates of at most `max_small` nodes into their most strongly connected
    neighbouring aggregate: the distance-1 independent set leaves many singletons and pairs
    (2.75 nodes per aggregate on a 5-point stencil); merged, aggregates hold ~4.5 nodes and the
    operator complexity of the hierarchy drops from ~2.5 to ~1.8 at the same convergence
    (tools/proto_lsc.py).  Pure torch (setup code, runs on whatever device the inputs are on).

    Rules (no chains: a merge target never merges itself): a small aggregate goes to its
    strongest neighbour if that one is large; two small ones that choose each other pair up
    (the lower id survives); a small one whose choice is a small one that stays joins it.
    -> (agg, n_agg, root_id) with ids in the order of the surviving roots."""
    dev = agg.device
    n = agg.numel()
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (indptr[1:] - indptr[:-1]).long())
    a, b = agg.long()[rows], agg.long()[indices.long()]
    off = a != b
    key = a[off] * n_agg + b[off]
    w = data[off].abs()
    ukey, inv = torch.unique(key, return_inverse=True)
    wsum = torch.zeros(ukey.numel(), dtype=w.dtype, device=dev).scatter_add_(0, inv, w)
    ga, gb = ukey // n_agg, ukey % n_agg
    size = torch.bincount(agg.long(), minlength=n_agg)
    small = size <= max_small
    ok = small[ga] & (size[ga] + size[gb] <= max_size)
    ga, gb, wsum = ga[ok], gb[ok], wsum[ok]
    best_w = torch.zeros(n_agg, dtype=w.dtype, device=dev).scatter_reduce_(0, ga, wsum, 'amax', include_self=True)
    is_best = wsum == best_w[ga]
    best = torch.full((n_agg,), n_agg, dtype=torch.long, device=dev)
    best.scatter_reduce_(0, ga[is_best], gb[is_best], 'amin', include_self=True)
    has = best < n_agg                                    # small aggregates with a candidate
    ids = torch.arange(n_agg, device=dev)
    tgt = torch.where(has, best, ids)
    tgt_c = tgt.clamp(max=n_agg - 1)
    to_large = has & ~small[tgt_c]
    mutual = has & small[tgt_c] & (tgt[tgt_c] == ids)
    loser = mutual & (ids > tgt)                          # the higher id of a pair merges
    merging = to_large | loser
    stays = ~merging
    to_staying_small = has & ~merging & ~mutual & small[tgt_c] & stays[tgt_c] & ~has[tgt_c].logical_and(merging[tgt_c])
    # a target chosen in the last rule must not be merging itself (it is not: stays[tgt])
    merging = merging | to_staying_small
    final = torch.where(merging, tgt, ids)
    # the rule above can only point at aggregates that stay; assert the no-chain property cheaply
    final = torch.where(merging[final.clamp(max=n_agg - 1)], ids, final)
    survive = final == ids
    new_id = torch.cumsum(survive.long(), 0) - 1
    agg_new = new_id[final[agg.long()]].to(torch.int32)
    n_new = int(survive.sum().item())
    flag = (root_id[1:] - root_id[:-1]).long()            # 1 at the root node of each old aggregate
    flag = flag * survive[agg.long()].long()
    rid = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    rid[1:] = torch.cumsum(flag, 0).to(torch.int32)
    return agg_new.contiguous(), n_new, rid


def filtered(A, theta):
    """A with its weak off-diagonal entries lumped onto the diagonal"""
    m = A.shape[0]
    cnt, ptr = _empty(m, torch.int32), _empty(m + 1, torch.int32)
    stats = _empty(2, torch.int64)
    _call('npb_csr_filter_symbolic', m, A.indptr, A.indices, A.data, float(theta), cnt, ptr,
          _scan_ws(m), stats)
    nnz = int(stats[0].item())
    idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
    _call('npb_csr_filter_numeric', m, A.indptr, A.indices, A.data, float(theta), ptr, idx, dat)
    return DevCsr(A.shape, ptr, idx, dat, (), nnz, A.max_row, False)


def truncated(P, theta):
    """P without its small entries (|p| < theta * row max), rows rescaled"""
    m = P.shape[0]
    cnt, ptr = _empty(m, torch.int32), _empty(m + 1, torch.int32)
    stats = _empty(2, torch.int64)
    _call('npb_csr_truncate_symbolic', m, P.indptr, P.indices, P.data, float(theta), cnt, ptr,
          _scan_ws(m), stats)
    nnz = int(stats[0].item())
    idx, dat = _empty(nnz, torch.int32), _empty(nnz, torch.float64)
    _call('npb_csr_truncate_numeric', m, P.indptr, P.indices, P.data, float(theta), ptr, idx, dat)
    return DevCsr(P.shape, ptr, idx, dat, (), nnz, P.max_row, True)


# a level of at most this many unknowns ends the hierarchy (dense inverse)
COARSE_MAX = 600


class Hierarchy:
    """Smoothed-aggregation AMG for one operator."""

    def __init__(self, A, theta=0.0, smooth=True, nu=1, omega=0.67, omega_p=0.67,
                 coarse_max=None, max_levels=24, agg_passes=1, p_trunc=0.0, agg='mis1',
                 gamma=1, gamma_from=2):
        A = A.materialized()
        if coarse_max is None:
            coarse_max = COARSE_MAX
        self.mats, self.keep = [], []
        levels = []
        cur = A
        agg_kind = agg
        while True:
            n = cur.shape[0]
            d = diagonal(cur)
            dinv = torch.where(d != 0, 1.0 / d, torch.ones_like(d))
            lev = dict(A=cur, dinv=dinv)
            levels.append(lev)
            if n <= coarse_max or len(levels) >= max_levels:
                break
            base_kind = 'mis1' if agg_kind == 'mis1m' else agg_kind
            agg, n_agg, root_id = aggregate(cur, theta, kind=base_kind, with_roots=True)
            if n_agg > 0.6 * n and theta > 0:      # filtered graph too sparse down here
                agg, n_agg, root_id = aggregate(cur, 0.0, kind=base_kind, with_roots=True)
            if agg_kind == 'mis1m' and 0 < n_agg < 0.9 * n:
                agg, n_agg, root_id = merge_small_aggregates(cur.indptr, cur.indices, cur.data, agg,
                                                             n_agg, root_id)
            if n_agg >= 0.9 * n or n_agg == 0:
                break
            for _ in range(1, agg_passes):
                # aggregate the aggregates (distance-2 style, ~3x3 cells in 2-D):
                # graph of aggregates = rows/cols relabelled by aggregate id,
                # |entries| summed -- one sort
                if n_agg <= coarse_max:
                    break
                rows = agg[cur.rows().long()]
                cols = agg[cur.indices.long()]
                Gagg = coo_to_csr((n_agg, n_agg), rows.contiguous(), cols.contiguous(),
                                  cur.data.abs(), prune=False)
                agg2, n2 = aggregate(Gagg, 0.0, kind='mis1')
                if n2 == 0 or n2 >= 0.9 * n_agg:
                    break
                agg, n_agg, root_id = agg2[agg.long()].contiguous(), n2, None
            lev['root_id'] = root_id       # shard.py: coarse ownership follows the roots
            T = sel_jac(agg, n_agg, full=True, monotone=False)
            if smooth:
                # P = (I - omega D_F^-1 A_F) T with the strength-filtered A_F
                AF = filtered(cur, theta) if theta > 0 else cur
                dF = diagonal(AF) if theta > 0 else d
                dFinv = torch.where(dF != 0, 1.0 / dF, torch.ones_like(dF))
                AT = _csr_times_sel(AF, T)
                P = _csr_plus_csr(T.todev(), _diag_times_csr(dia_jac(-omega_p * dFinv), AT))
                if p_trunc > 0:
                    P = truncated(P.materialized(), p_trunc)
                R = P.transpose()
                Ac = _csr_times_csr(R, _csr_times_csr(cur, P))
            else:
                # piecewise-constant P: the Galerkin product is a relabel of
                # rows and columns by aggregate id + duplicate sum (one sort)
                P = T.todev()
                R = P.transpose()
                rows = agg[cur.rows().long()]
                cols = agg[cur.indices.long()]
                Ac = coo_to_csr((n_agg, n_agg), rows.contiguous(), cols.contiguous(),
                                cur.data, prune=False)
            lev['P'], lev['R'] = P.materialized(), R.materialized()
            cur = Ac.materialized()
        self.levels = levels
        nl = len(levels)
        self.arr = (AmgLevel * nl)()
        for l, lev in enumerate(levels):
            n = lev['A'].shape[0]
            L = self.arr[l]
            L.A = csr_mat(lev['A'], self.keep)
            if 'P' in lev:
                L.P, L.R = csr_mat(lev['P'], self.keep), csr_mat(lev['R'], self.keep)
            L.dinv = lev['dinv'].data_ptr()
            lev['work'] = [torch.zeros(n, dtype=torch.float64, device=device()) for _ in range(4)]
            L.x, L.x2, L.b, L.r = [w.data_ptr() for w in lev['work']]
        self.struct = AmgStruct()
        self.struct.nlevels = nl
        self.struct.nu_pre = self.struct.nu_post = int(nu)
        self.struct.omega = float(omega)
        self.struct.gamma, self.struct.gamma_from = int(gamma), int(gamma_from)
        self.struct.levels = ctypes.cast(self.arr, ctypes.POINTER(AmgLevel))
        last = levels[-1]['A']
        nc = last.shape[0]
        if nc <= 2048:
            W = _empty(2 * nc * nc, torch.float64)
            self.coarse_inv = _empty(nc * nc, torch.float64)
            info = _empty(1, torch.int32)
            _call('npb_dense_inverse', nc, last.indptr, last.indices, last.data, W,
                  self.coarse_inv, info)
            if int(info.item()) == 0:
                self.struct.coarse_n = nc
                self.struct.coarse_inv = self.coarse_inv.data_ptr()
            else:
                self.struct.coarse_n = 0
        else:
            self.struct.coarse_n = 0
        self.sizes = [lev['A'].shape[0] for lev in levels]
        self.nnzs = [lev['A'].nnz for lev in levels]

    @property
    def addr(self):
        return ctypes.addressof(self.struct)

    def vcycle(self, b, x=None):
        if x is None:
            x = torch.empty_like(b)
        _lib.lib().call('npb_amg_vcycle', self.addr, b, x, stream_ptr())
        return x

    def describe(self):
        return 'levels %s, operator complexity %.2f' % (
            self.sizes, sum(self.nnzs) / max(self.nnzs[0], 1))


class AmgPreconditioner:
    """k V-cycles as an npb_apply_fn (C callback, no Python in the loop)."""

    def __init__(self, A, cycles=1, **kw):
        self.A = A.materialized()
        self.h = Hierarchy(self.A, **kw)
        n = A.shape[0]
        self.work = [torch.zeros(n, dtype=torch.float64, device=device()) for _ in range(2)]
        self.struct = AmgPrecondStruct()
        self.struct.amg = self.h.addr
        self.struct.A = csr_mat(self.A)
        self.struct.cycles = int(cycles)
        self.struct.r, self.struct.dx = [w.data_ptr() for w in self.work]
        self.c_apply = ctypes.cast(_lib.lib().cdll.npb_amg_apply_cb, ctypes.c_void_p).value
        self.c_ctx = ctypes.addressof(self.struct)

    def apply(self, r, z):
        rc = _lib.lib().cdll.npb_amg_apply_cb(self.c_ctx, r.data_ptr(), z.data_ptr(), stream_ptr())
        if rc:
            raise RuntimeError('npb_amg_apply_cb failed: %s' % _lib.lib().last_error())


def _submatrix(J, rows, colmap, ncols):
    """J[rows][:, cols] with cols given as a map old column -> new column / -1"""
    Jr = _sel_times_csr(sel_jac(rows, J.shape[1], full=True, monotone=True), J)
    return _csr_times_sel(Jr, sel_jac(colmap, ncols, monotone=True, full=False)).materialized()


class LscPreconditioner:
    """Block preconditioner for J = [[A, G], [D, 0]] split by zero diagonal."""

    def __init__(self, J, cycles_a=1, cycles_l=(3, 4), theta_a=0.25, nu_a=4, nu_l=1, smooth_a=True,
                 reuse=None, a_gmres=0, agg_passes=1, agg_passes_a=None, p_trunc=0.1,
                 agg_l='mis1', agg_a='mis2', gamma_l=1, gamma_a=1, gamma_from=2):
        J = J.materialized()
        n = J.shape[0]
        d = diagonal(J)
        has = d != 0
        vset = torch.nonzero(has).reshape(-1).to(torch.int32)
        pset = torch.nonzero(~has).reshape(-1).to(torch.int32)
        nv, np_ = int(vset.numel()), int(pset.numel())
        self.n, self.nv, self.np = n, nv, np_
        ar_v = torch.arange(nv, dtype=torch.int32, device=device())
        ar_p = torch.arange(np_, dtype=torch.int32, device=device())
        vmap = torch.full((n,), -1, dtype=torch.int32, device=device())
        vmap[vset.long()] = ar_v
        pmap = torch.full((n,), -1, dtype=torch.int32, device=device())
        pmap[pset.long()] = ar_p
        self.vset, self.pset = vset, pset
        self.A = _submatrix(J, vset, vmap, nv)
        self.G = _submatrix(J, vset, pmap, np_)
        self.D = _submatrix(J, pset, vmap, nv)
        # L = -D G: constant across Newton iterations when the constraint
        # blocks are (they are the linear div / grad operators); reuse its
        # hierarchy when the caller hands back an LscPreconditioner whose D, G
        # are bit-identical
        if (reuse is not None and reuse.np == np_ and reuse.nv == nv and
                reuse.D.nnz == self.D.nnz and reuse.G.nnz == self.G.nnz and
                torch.equal(reuse.D.indices, self.D.indices) and
                torch.equal(reuse.D.data, self.D.data) and
                torch.equal(reuse.G.indices, self.G.indices) and
                torch.equal(reuse.G.data, self.G.data)):
            self.Lm, self.amg_l = reuse.Lm, reuse.amg_l
        else:
            self.Lm = _csr_times_csr(self.D, self.G).scaled(-1.0).materialized()
            self.amg_l = Hierarchy(self.Lm, theta=0.0, nu=nu_l, agg_passes=agg_passes,
                                   p_trunc=p_trunc, agg=agg_l, gamma=gamma_l, gamma_from=gamma_from)
        # momentum block: distance-2 aggregation on the strength graph, prolongator
        # smoothed with the strength-filtered operator (keeps the two velocity
        # components apart); operator complexity 1.4 (2.97 with distance-1 aggregates:
        # their 32-entry coarse rows made this hierarchy 60 % of the setup time) paid
        # for with 4 Jacobi sweeps instead of 2 -- same outer iteration count at 2048^2
        if a_gmres > 0:
            # robust mode: no multigrid for the momentum block (see csrc/amg.cu)
            self.amg_a = None
            dA = diagonal(self.A)
            self.a_dinv = torch.where(dA != 0, 1.0 / dA, torch.ones_like(dA))
            wb = _lib.lib().cdll.npb_fgmres_workspace_bytes(nv, int(a_gmres), 0)
            self.a_work = _empty(wb, torch.uint8)
        else:
            self.amg_a = Hierarchy(self.A, theta=theta_a, nu=nu_a, smooth=smooth_a,
                                   agg_passes=agg_passes if agg_passes_a is None else agg_passes_a,
                                   p_trunc=p_trunc, agg=agg_a, gamma=gamma_a, gamma_from=gamma_from)
            self.a_dinv = self.amg_a.levels[0]['dinv']
        self.a_gmres = int(a_gmres)
        self.wv = [torch.zeros(nv, dtype=torch.float64, device=device()) for _ in range(5)]
        self.wp = [torch.zeros(max(np_, 1), dtype=torch.float64, device=device()) for _ in range(8)]
        self.scal = torch.zeros(4, dtype=torch.float64, device=device())
        self.red = torch.zeros(_lib.lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8,
                               device=device())
        s = self.struct = LscStruct()
        s.n, s.nv, s.np = n, nv, np_
        s.vset, s.pset = vset.data_ptr(), pset.data_ptr()
        self.keep = []
        s.A, s.G, s.D, s.Lm = (csr_mat(m, self.keep) for m in (self.A, self.G, self.D, self.Lm))
        s.a_dinv = self.a_dinv.data_ptr()
        s.amg_a = self.amg_a.addr if self.amg_a is not None else None
        s.amg_l = self.amg_l.addr
        s.a_gmres = self.a_gmres
        # the apply's launch sequence is fixed (device-resident CG scalars): csrc/amg.cu replays
        # it as a CUDA graph under this id
        LscPreconditioner._next_graph_id += 1
        s.graph_id = self.graph_id = LscPreconditioner._next_graph_id
        if self.a_gmres > 0:
            s.a_work, s.a_work_bytes = self.a_work.data_ptr(), self.a_work.numel()
        if isinstance(cycles_l, (tuple, list)):      # (first L-solve, second L-solve)
            cycles_l = int(cycles_l[0]) | (int(cycles_l[1]) << 16)
        s.cycles_a, s.cycles_l = int(cycles_a), int(cycles_l)
        for k in range(5):
            s.wv[k] = self.wv[k].data_ptr()
        for k in range(8):
            s.wp[k] = self.wp[k].data_ptr()
        s.scal, s.red = self.scal.data_ptr(), self.red.data_ptr()
        self.c_apply = ctypes.cast(_lib.lib().cdll.npb_lsc_apply_cb, ctypes.c_void_p).value
        self.c_ctx = ctypes.addressof(s)

    _next_graph_id = 0

    def __del__(self):
        try:
            _lib.lib().cdll.npb_lsc_graph_release(int(self.graph_id))
        except Exception:
            pass

    def apply(self, r, z):
        rc = _lib.lib().cdll.npb_lsc_apply_cb(self.c_ctx, r.data_ptr(), z.data_ptr(), stream_ptr())
        if rc:
            raise RuntimeError('npb_lsc_apply_cb failed: %s' % _lib.lib().last_error())

    def describe(self):
        a = self.amg_a.describe() if self.amg_a is not None else 'GMRES(%d)-Jacobi' % self.a_gmres
        return 'LSC: nv=%d np=%d | A: %s | L-AMG %s' % (self.nv, self.np, a, self.amg_l.describe())
