"""Row-sharded Newton linear solve across the GPUs of one node.

One process per GPU.  The Jacobian is permuted to strip-major order (every field of a
grid strip contiguous), each rank keeps the rows of its strip of every operator -- the
Jacobian, the blocks of the least-squares-commutator preconditioner and the large levels
of both multigrid hierarchies -- and the C++ solve phase (csrc/amg.cu, gmres.cu) runs
unchanged on the local rows: an SpMV is preceded by `pack boundary values -> all-gather`
on the library's own NCCL communicator (csrc/dist.cu), an inner product is followed by
an all-reduce of the device scalars.  The small multigrid levels are replicated: the
last sharded level all-gathers its restricted residual and every rank runs the coarse
part of the V-cycle redundantly (latency, not bandwidth, is what costs there).

Numerically this is the single-GPU solver (same hierarchies, same smoother, same Krylov
recurrences; only the order of the floating-point sums in the inner products differs),
so iteration counts do not depend on the number of GPUs -- unlike the block-Jacobi
preconditioner of the reference's MpiJacobian.approx_solve (admpisolve.py:167-172),
which the saddle-point structure of the cavity Jacobian does not tolerate.

This round most of the SETUP is replicated: every rank builds the same global hierarchies
(deterministic kernels) and slices its rows out; only the large general products of the
Galerkin operators are computed by row blocks, one per rank, and all-gathered
(`_dev._csr_times_csr_sharded`, bit-identical to the local product).  The slicing logic
(`localize`) is plain torch and is tested on CPU; the halo numbering it produces is what
`npb_halo` (include/numpad_b200.h) documents.  Exchanges go by direct stores into NVLink peer
memory through a symmetric CUDA-IPC heap when the ranks share a node (<= 8), else NCCL.
"""
import ctypes
import os

import numpy as np
import torch

# levels with fewer rows than this are replicated instead of sharded
REPLICATE_BELOW = 120000

# general products of the (otherwise replicated) multigrid setup with at least this many rows
# are computed by row blocks across the ranks and all-gathered (0: never)
SHARD_PRODUCTS_ABOVE = 500000

# bytes of the symmetric heap for direct-store exchanges (0 / NPB_DIST_P2P=0: NCCL only)
HEAP_BYTES = 64 << 20
# replay the sharded preconditioner as a CUDA graph (needs the direct-store exchanges)
GRAPHS = os.environ.get('NPB_DIST_GRAPHS', '1') != '0'

_handle = {'ptr': None, 'rank': 0, 'world': 1, 'p2p': False, 'heap': None}


def _nccl_path():
    try:
        import nvidia.nccl
        p = os.path.join(os.path.dirname(nvidia.nccl.__file__), 'lib', 'libnccl.so.2')
        return p if os.path.exists(p) else ''
    except Exception:
        return ''


def init(group=None):
    """Collective: create the library's NCCL communicator over the ranks of the default
    torch.distributed process group.  Returns the handle (int address)."""
    import torch.distributed as dist
    from . import _lib
    if _handle['ptr'] is not None:
        return _handle['ptr']
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    L = _lib.lib().cdll
    path = _nccl_path().encode()
    uid = ctypes.create_string_buffer(128)
    if rank == 0:
        rc = L.npb_dist_unique_id(path, uid)
        if rc:
            raise RuntimeError('npb_dist_unique_id failed: %s' % _lib.lib().last_error())
    t = torch.frombuffer(bytearray(uid.raw), dtype=torch.uint8).clone()
    dev = torch.device('cuda', torch.cuda.current_device())
    t = t.to(dev)
    dist.broadcast(t, 0, group=group)
    uid = ctypes.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
    h = ctypes.c_void_p()
    rc = L.npb_dist_init(path, uid, rank, world, ctypes.byref(h))
    if rc:
        raise RuntimeError('npb_dist_init failed: %s' % _lib.lib().last_error())
    _handle.update(ptr=h.value, rank=rank, world=world, p2p=False)
    if HEAP_BYTES and world <= 8 and os.environ.get('NPB_DIST_P2P', '1') != '0':
        ipc = ctypes.create_string_buffer(72)
        # the heap is a torch allocation owned here (the library allocates nothing); its own
        # cudaMalloc segment, so that the IPC mapping exposes nothing else of this process
        _handle['heap'] = torch.empty(HEAP_BYTES, dtype=torch.uint8, device=dev)
        rc = L.npb_dist_heap_create(h, ctypes.c_void_p(_handle['heap'].data_ptr()), HEAP_BYTES, ipc)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            mine = torch.frombuffer(bytearray(ipc.raw), dtype=torch.uint8).clone().to(dev)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine, group=group)
            blob = b''.join(bytes(t.cpu().numpy().tobytes()) for t in allh)
            rc = L.npb_dist_heap_open(h, blob)
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            _handle['p2p'] = int(ok.item()) == 1
    return h.value


def shutdown():
    from . import _lib
    if _handle['ptr'] is not None:
        _lib.lib().cdll.npb_dist_free(ctypes.c_void_p(_handle['ptr']))
        _handle.update(ptr=None, rank=0, world=1, p2p=False, heap=None)


# --------------------------------------------------------------------------- #
# slicing a replicated global CSR operator into the rows of one rank (pure torch)
# --------------------------------------------------------------------------- #
class LocalOp:
    """rows [rcuts[rank], rcuts[rank+1]) of a global CSR matrix with the columns renumbered
    [owned | boundary values of rank 0 | rank 1 | ...] (npb_halo in the C header)."""
    __slots__ = ('shape', 'indptr', 'indices', 'data', 'nnz', 'n_own', 'send_idx', 'maxb',
                 'recv', 'sharded_cols', 'data32')


def localize(indptr, indices, data, shape, rcuts, ccuts, rank, world, replicated_cols=False):
    """-> LocalOp.  rcuts / ccuts: world + 1 increasing python ints (row / column ownership
    ranges).  replicated_cols: the operand vector is replicated (global column numbering is
    kept, no halo)."""
    dev = indices.device
    op = LocalOp()
    r0, r1 = int(rcuts[rank]), int(rcuts[rank + 1])
    kc = indptr[torch.as_tensor(list(rcuts), dtype=torch.int64, device=dev)].to(torch.int64)
    kcl = [int(v) for v in kc.tolist()]
    k0, k1 = kcl[rank], kcl[rank + 1]
    op.indptr = (indptr[r0:r1 + 1] - k0).to(torch.int32).contiguous()
    # a view when the slice starts on a 16-byte boundary (the TMA SpMV needs that), else a copy
    op.data = data[k0:k1] if k0 % 2 == 0 else data[k0:k1].clone()
    op.nnz = k1 - k0
    op.sharded_cols = not replicated_cols
    if replicated_cols:
        op.shape = (r1 - r0, shape[1])
        op.indices = indices[k0:k1] if k0 % 4 == 0 else indices[k0:k1].clone()
        op.n_own, op.maxb = shape[1], 0
        op.send_idx = torch.zeros(0, dtype=torch.int32, device=dev)
        op.recv = torch.zeros(1, dtype=torch.float64, device=dev)
        return op
    n_own = int(ccuts[rank + 1]) - int(ccuts[rank])
    # boundary sets of ALL ranks (every rank derives the same lists from the replicated
    # matrix): columns referenced by a row of another rank.  The entries of rank q's rows
    # are the contiguous range [kc[q], kc[q+1]): two comparisons per entry, no search.
    rem = []
    for q in range(world):
        seg = indices[kcl[q]:kcl[q + 1]]
        m = (seg < int(ccuts[q])) | (seg >= int(ccuts[q + 1]))
        rem.append(seg[m])
    bcols = torch.unique(torch.cat(rem)) if world > 1 else indices[:0]    # sorted
    cc = torch.as_tensor(list(ccuts), dtype=bcols.dtype, device=dev)
    bstart = torch.searchsorted(bcols, cc)     # world + 1 segment bounds
    bstart_l = [int(v) for v in bstart.tolist()]
    maxb = max([bstart_l[q + 1] - bstart_l[q] for q in range(world)] + [0])
    lc = indices[k0:k1]
    own = (lc >= int(ccuts[rank])) & (lc < int(ccuts[rank + 1]))
    new_idx = (lc - int(ccuts[rank])).to(torch.int32)
    ridx = torch.nonzero(~own).reshape(-1)                 # the few entries that read a halo
    if ridx.numel():
        rc_ = lc[ridx]
        lo = torch.searchsorted(cc, rc_, right=True) - 1
        pos = torch.searchsorted(bcols, rc_)
        new_idx[ridx] = (n_own + lo * maxb + (pos - bstart[lo])).to(torch.int32)
    op.indices = new_idx.contiguous()
    b0, b1 = bstart_l[rank], bstart_l[rank + 1]
    op.send_idx = (bcols[b0:b1] - int(ccuts[rank])).to(torch.int32).contiguous()
    op.n_own, op.maxb = n_own, maxb
    op.shape = (r1 - r0, n_own + world * maxb)
    op.recv = torch.zeros(max(world * maxb, 1), dtype=torch.float64, device=dev)
    return op


def local_matvec_reference(op, x_own, gathered):
    """y = op @ [x_own | gathered] with plain torch (CPU tests of the numbering)"""
    xe = torch.cat([x_own, gathered])
    rows = torch.repeat_interleave(torch.arange(op.shape[0], device=x_own.device),
                                   (op.indptr[1:] - op.indptr[:-1]).long())
    y = torch.zeros(op.shape[0], dtype=x_own.dtype, device=x_own.device)
    y.index_add_(0, rows, op.data * xe[op.indices.long()])
    return y


def strip_permutation(owner, world):
    """stable sort of the unknowns by owning rank -> (perm, inv, cuts): new index i holds
    old index perm[i]; cuts = world + 1 range bounds in the new numbering"""
    owner = torch.as_tensor(owner)
    perm = torch.sort(owner.to(torch.int64), stable=True).indices
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), dtype=torch.int64, device=perm.device)
    counts = torch.bincount(owner.to(torch.int64), minlength=world)
    cuts = [0] + [int(v) for v in torch.cumsum(counts, 0).tolist()]
    return perm, inv, cuts


# --------------------------------------------------------------------------- #
# device side: sharded hierarchies, LSC preconditioner, FGMRES
# --------------------------------------------------------------------------- #
def _fill_mat(m, op, handle, f32=False):
    """fill an amg.CsrMat from a LocalOp"""
    m.nrows, m.ncols, m.nnz = op.shape[0], op.shape[1], op.nnz
    m.indptr, m.indices, m.data = op.indptr.data_ptr(), op.indices.data_ptr(), op.data.data_ptr()
    if f32:
        # preconditioner operators: applied from fp32 values when large (amg.F32_MIN_ROWS)
        from . import amg
        if amg.F32_MIN_ROWS and op.shape[0] >= amg.F32_MIN_ROWS and 0 < op.nnz <= 96 * op.shape[0]:
            op.data32 = op.data.to(torch.float32)
            m.data32 = op.data32.data_ptr()
    if op.sharded_cols:
        m.halo.dist = handle
        m.halo.n_own = op.n_own
        m.halo.send_idx = op.send_idx.data_ptr()
        m.halo.n_send = int(op.send_idx.numel())
        m.halo.maxb = op.maxb
        m.halo.recv = op.recv.data_ptr()
        m.halo.p2p_data = m.halo.p2p_flags = -1
        if _handle['p2p'] and op.maxb > 0:
            from . import _lib
            L = _lib.lib().cdll
            world = _handle['world']
            d = L.npb_dist_heap_alloc(ctypes.c_void_p(handle), 2 * world * op.maxb * 8)
            f = L.npb_dist_heap_alloc(ctypes.c_void_p(handle), world * 8)
            if d >= 0 and f >= 0:
                m.halo.p2p_data, m.halo.p2p_flags = d, f
                m.halo.seq = _new_counter()
            else:
                _all_p2p[0] = False        # an exchange of this solve goes through NCCL
        elif op.maxb > 0:
            _all_p2p[0] = False
    return m


_seq_keep = []          # device exchange counters of the current solve's operators
_all_p2p = [True]       # every exchange of the current solve goes by direct stores


def _new_counter():
    """exchange counter of one operator, in device memory (the kernels read and bump it: the
    launch sequence of the solve phase does not depend on it and can be graph-captured)"""
    from ._dev import device
    t = torch.zeros(1, dtype=torch.int64, device=device())
    _seq_keep.append(t)
    return t.data_ptr()


def _heap_reset():
    """collective: recycle the symmetric heap between solves"""
    import torch.distributed as dist
    from . import _lib
    from ._dev import stream_ptr
    if not _handle['p2p']:
        return
    torch.cuda.synchronize()
    dist.barrier()
    _lib.lib().call('npb_dist_heap_reset', ctypes.c_void_p(_handle['ptr']), stream_ptr())
    _seq_keep.clear()
    _all_p2p[0] = True
    dist.barrier()


def _loc(A, rcuts, ccuts, replicated_cols=False):
    A = A.materialized()
    return localize(A.indptr, A.indices, A.data, A.shape, rcuts, ccuts, _handle['rank'],
                    _handle['world'], replicated_cols)


class ShardedHierarchy:
    """An amg.Hierarchy (built on the global operator, replicated) whose large levels are
    applied on the local rows."""

    def __init__(self, h, cuts):
        from . import amg
        from ._dev import device
        rank, world, handle = _handle['rank'], _handle['world'], _handle['ptr']
        self.h, self.keep = h, []
        nl = len(h.levels)
        # cuts of every level: a coarse index belongs to the rank owning its root
        lev_cuts = [list(cuts)]
        for l in range(nl - 1):
            rid = h.levels[l].get('root_id')
            if rid is None:
                raise RuntimeError('hierarchy without root ids cannot be sharded')
            c = rid[torch.as_tensor(lev_cuts[-1], dtype=torch.int64, device=rid.device)]
            lev_cuts.append([int(v) for v in c.tolist()])
        # first replicated level: the first one below REPLICATE_BELOW rows (at least the last)
        first_rep = nl - 1
        for l in range(1, nl):
            if h.sizes[l] < REPLICATE_BELOW:
                first_rep = l
                break
        self.first_rep, self.lev_cuts = first_rep, lev_cuts
        self.arr = (amg.AmgLevel * nl)()
        self.work = []
        for l, lev in enumerate(h.levels):
            L = self.arr[l]
            if l >= first_rep:                 # replicated: the global level as it is
                src = h.arr[l]
                ctypes.memmove(ctypes.addressof(L), ctypes.addressof(src), ctypes.sizeof(amg.AmgLevel))
                continue
            c, cn = lev_cuts[l], lev_cuts[l + 1]
            a = _loc(lev['A'], c, c)
            last = (l + 1 == first_rep)
            p = _loc(lev['P'], c, cn, replicated_cols=last)
            r = _loc(lev['R'], cn, c)
            self.keep += [a, p, r]
            _fill_mat(L.A, a, handle, True); _fill_mat(L.P, p, handle, True); _fill_mat(L.R, r, handle, True)
            dinv = lev['dinv'][c[rank]:c[rank + 1]]
            n = c[rank + 1] - c[rank]
            w = [torch.zeros(max(n, 1), dtype=torch.float64, device=device()) for _ in range(4)]
            self.work.append(w)
            self.keep.append(dinv)
            L.dinv = dinv.data_ptr()
            L.x, L.x2, L.b, L.r = [t.data_ptr() for t in w]
            if last:
                off = (ctypes.c_int64 * (world + 1))(*cn)
                self.keep.append(off)
                L.gather = 1
                L.offsets = ctypes.cast(off, ctypes.c_void_p).value
                L.g_data = L.g_flags = -1
                if _handle['p2p']:
                    from . import _lib
                    Lc = _lib.lib().cdll
                    d = Lc.npb_dist_heap_alloc(ctypes.c_void_p(handle), 2 * cn[world] * 8)
                    f = Lc.npb_dist_heap_alloc(ctypes.c_void_p(handle), world * 8)
                    if d >= 0 and f >= 0:
                        L.g_data, L.g_flags, L.g_seq = d, f, _new_counter()
                    else:
                        _all_p2p[0] = False
                else:
                    _all_p2p[0] = False
        s = self.struct = amg.AmgStruct()
        ctypes.memmove(ctypes.addressof(s), ctypes.addressof(h.struct), ctypes.sizeof(amg.AmgStruct))
        s.levels = ctypes.cast(self.arr, ctypes.POINTER(amg.AmgLevel))
        s.dist = handle

    @property
    def addr(self):
        return ctypes.addressof(self.struct)


class ShardedLsc:
    """amg.LscPreconditioner applied on the rows of this rank (csrc/amg.cu npb_lsc_apply_cb
    with sharded operators)."""

    def __init__(self, M, cuts):
        from . import amg, _lib
        from ._dev import device
        rank, handle = _handle['rank'], _handle['ptr']
        if M.amg_a is None:
            raise RuntimeError('the robust (Krylov) momentum solve is not sharded')
        self.M = M
        dev = M.vset.device
        ct = torch.as_tensor(cuts, dtype=torch.int32, device=dev)
        cv = [int(v) for v in torch.searchsorted(M.vset, ct).tolist()]
        cp = [int(v) for v in torch.searchsorted(M.pset, ct).tolist()]
        self.cuts, self.cv, self.cp = list(cuts), cv, cp
        self.vset = (M.vset[cv[rank]:cv[rank + 1]] - cuts[rank]).contiguous()
        self.pset = (M.pset[cp[rank]:cp[rank + 1]] - cuts[rank]).contiguous()
        nv, np_ = cv[rank + 1] - cv[rank], cp[rank + 1] - cp[rank]
        self.A, self.G = _loc(M.A, cv, cv), _loc(M.G, cv, cp)
        self.D, self.Lm = _loc(M.D, cp, cv), _loc(M.Lm, cp, cp)
        self.amg_a = ShardedHierarchy(M.amg_a, cv)
        self.amg_l = ShardedHierarchy(M.amg_l, cp)
        self.a_dinv = M.a_dinv[cv[rank]:cv[rank + 1]]
        self.wv = [torch.zeros(max(nv, 1), dtype=torch.float64, device=device()) for _ in range(5)]
        self.wp = [torch.zeros(max(np_, 1), dtype=torch.float64, device=device()) for _ in range(8)]
        self.scal = torch.zeros(4, dtype=torch.float64, device=device())
        self.red = torch.zeros(_lib.lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8,
                               device=device())
        s = self.struct = amg.LscStruct()
        s.n, s.nv, s.np = cuts[rank + 1] - cuts[rank], nv, np_
        s.vset, s.pset = self.vset.data_ptr(), self.pset.data_ptr()
        for m, op in ((s.A, self.A), (s.G, self.G), (s.D, self.D), (s.Lm, self.Lm)):
            _fill_mat(m, op, handle, True)
        s.a_dinv = self.a_dinv.data_ptr()
        s.amg_a, s.amg_l = self.amg_a.addr, self.amg_l.addr
        s.a_gmres = 0
        # CUDA-graph replay of the apply (csrc/amg.cu): only when every exchange inside it goes by
        # direct stores with device-side counters (no NCCL call, nothing host-dependent)
        self.graph_id = 0
        if _handle['p2p'] and _all_p2p[0] and GRAPHS:
            amg.LscPreconditioner._next_graph_id += 1
            self.graph_id = amg.LscPreconditioner._next_graph_id
        s.graph_id = self.graph_id
        s.cycles_a, s.cycles_l = M.struct.cycles_a, M.struct.cycles_l
        for k in range(5):
            s.wv[k] = self.wv[k].data_ptr()
        for k in range(8):
            s.wp[k] = self.wp[k].data_ptr()
        s.scal, s.red = self.scal.data_ptr(), self.red.data_ptr()
        self.c_apply = ctypes.cast(_lib.lib().cdll.npb_lsc_apply_cb, ctypes.c_void_p).value
        self.c_ctx = ctypes.addressof(s)

    def __del__(self):
        try:
            if self.graph_id:
                from . import _lib
                _lib.lib().cdll.npb_lsc_graph_release(int(self.graph_id))
        except Exception:
            pass


def permuted(J, perm, inv):
    """Pi J Pi^T for new index i = old index perm[i] (inv = inverse map): rows gathered in the
    new order, columns relabelled and every row re-sorted in place (short rows), instead of a
    device-wide sort of all entries"""
    from ._dev import coo_to_csr, sel_jac, _sel_times_csr, _empty, _call, DevCsr
    J = J.materialized()
    if J.max_row > 64:
        rows = inv[J.rows().long()].to(torch.int32)
        cols = inv[J.indices.long()].to(torch.int32)
        return coo_to_csr(J.shape, rows, cols, J.data, prune=False)
    Jr = _sel_times_csr(sel_jac(perm.to(torch.int32), J.shape[0], full=True, monotone=False), J)
    Jr = Jr.materialized()
    idx, dat = _empty(Jr.nnz, torch.int32), _empty(Jr.nnz, torch.float64)
    _call('npb_csr_relabel_sort_rows', Jr.shape[0], Jr.indptr, Jr.indices, Jr.data,
          inv.to(torch.int32), idx, dat)
    return DevCsr(J.shape, Jr.indptr, idx, dat, (), Jr.nnz, Jr.max_row, True)


last_info = {}
# True: last_info['lsc_digests'] holds checksums of the replicated operators (LSC blocks and
# every multigrid level) so that a multi-rank test can assert they are bit-identical everywhere
DIGESTS = False


def _digest(t):
    v = t.reshape(-1)
    v = v.view(torch.int64) if v.dtype == torch.float64 else v.to(torch.int64)
    w = torch.arange(1, v.numel() + 1, device=v.device, dtype=torch.int64)
    return int(((v * (w % 1000003 + 1)).sum() % 9223372036854775783).item())


def _lsc_digests(M):
    out = []
    mats = [M.A, M.G, M.D, M.Lm]
    for h in (M.amg_a, M.amg_l):
        for lev in h.levels:
            mats.append(lev['A'])
            if 'P' in lev:
                mats.append(lev['P'])
    for m in mats:
        out += [_digest(m.indptr), _digest(m.indices), _digest(m.data)]
    return out


def solve(J, b, owner=None, rtol=1e-10, restart=100, maxiter=600, lsc_options=None):
    """x = J^-1 b with the rows of every operator sharded over the ranks of the process
    group (collective; J and b replicated on every rank, x returned replicated).
    owner[i] = rank that owns unknown i (numpy / torch int array; None: equal contiguous
    ranges of the natural ordering)."""
    import time
    import torch.distributed as dist
    from . import amg, linsolve, _lib
    from ._dev import device, coo_to_csr, DevCsr
    handle = init()
    rank, world = _handle['rank'], _handle['world']
    J = J.materialized()
    n = J.shape[0]
    dev = device()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if owner is None:
        owner = torch.clamp((torch.arange(n, device=dev) * world) // n, max=world - 1)
    owner = torch.as_tensor(owner).to(dev)
    perm, inv, cuts = strip_permutation(owner, world)
    # J' = Pi J Pi^T (strip-major), replicated
    Jp = permuted(J, perm, inv)
    bp = b.reshape(-1)[perm].contiguous()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    from . import _dev
    # the large Galerkin products of the setup are computed by row blocks, one per rank
    _dev.SPGEMM_SHARD = (rank, world, SHARD_PRODUCTS_ABOVE) if world > 1 and SHARD_PRODUCTS_ABOVE else None
    f32_rows, amg.F32_MIN_ROWS = amg.F32_MIN_ROWS, 0     # (fp32 copies are made of the LOCAL slices)
    dia_rows, amg.DIA_MIN_ROWS = amg.DIA_MIN_ROWS, 0     # (diagonal storage: single-GPU operators only)
    try:
        M = amg.LscPreconditioner(Jp, **(lsc_options or {}))
    finally:
        _dev.SPGEMM_SHARD = None
        amg.F32_MIN_ROWS = f32_rows
        amg.DIA_MIN_ROWS = dia_rows
    torch.cuda.synchronize(); t2 = time.perf_counter()
    _heap_reset()
    S = ShardedLsc(M, cuts)
    Jl = _loc(Jp, cuts, cuts)
    jm = amg.CsrMat()
    _fill_mat(jm, Jl, handle)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    L = _lib.lib().cdll
    mv = ctypes.cast(L.npb_mat_matvec_cb, ctypes.c_void_p).value
    ar = ctypes.cast(L.npb_dist_allreduce_cb, ctypes.c_void_p).value
    b_loc = bp[cuts[rank]:cuts[rank + 1]].contiguous()
    x_loc, res = linsolve.fgmres(None, b_loc, precond=S, rtol=rtol, restart=restart, maxiter=maxiter,
                                 matvec_c=(mv, ctypes.addressof(jm)), allreduce_c=(ar, handle))
    torch.cuda.synchronize(); t4 = time.perf_counter()
    # gather the strips and undo the permutation
    xp = torch.empty(n, dtype=torch.float64, device=dev)
    parts = [xp[cuts[q]:cuts[q + 1]] for q in range(world)]
    parts[rank].copy_(x_loc)
    for q in range(world):
        if parts[q].numel():
            dist.broadcast(parts[q], q)
    x = torch.empty_like(xp)
    x[perm] = xp
    torch.cuda.synchronize(); t5 = time.perf_counter()
    res.update(method='fgmres+lsc (row-sharded, %d ranks)' % world, permute_ms=1e3 * (t1 - t0),
               setup_ms=1e3 * (t2 - t1), localize_ms=1e3 * (t3 - t2), krylov_ms=1e3 * (t4 - t3),
               gather_ms=1e3 * (t5 - t4), precond=M.describe(),
               first_replicated_level=(S.amg_a.first_rep, S.amg_l.first_rep),
               halo=(Jl.maxb, S.A.maxb, S.Lm.maxb),
               exchange='direct stores into NVLink peer memory' if _handle['p2p'] else 'NCCL all-gather')
    if DIGESTS:
        res['lsc_digests'] = _lsc_digests(M)
    last_info.clear()
    last_info.update(res)
    if res['status'] != 0:
        raise linsolve.LinearSolveError('sharded linear solve did not converge: %d iterations, '
                                        '|r|/|b| = %.3e' % (res['iters'], res['rnorm'] / max(res['bnorm'], 1e-300)))
    return x
