"""Row-sharded sparse operators and the row-sharded Newton linear solve across
the GPUs of one node (seam S4).

One process per GPU (torchrun); `torch.distributed` (NCCL over NVLink on the
GPUs, gloo in the CPU tests) carries exactly two kinds of traffic, as the
reference's distributed solver does with mpi4py:

  * halo exchange of the operand before an SpMV
    -- replaces the partial-product exchange of `MpiJacobian.matvec`
       (admpisolve.py:146-165): x halos are exchanged instead of partial y,
       which makes the CSR->CSC ownership swap (`_csr_to_csc`, :61-98) and the
       pickled block shipping unnecessary;
  * one all-reduce of a few doubles per batch of inner products
    -- replaces one scalar `Allreduce` per dot in `_dot` / `_norm2`
       (admpisolve.py:177-188).

The preconditioner is block-Jacobi ACROSS ranks (as `MpiJacobian.approx_solve`,
admpisolve.py:167-172, which factorises the local diagonal block) applied
INSIDE the sub-solves of the least-squares-commutator preconditioner: the
Poisson-type and momentum sub-problems are solved by Krylov / stationary
iterations on the GLOBAL operators (distributed SpMV) preconditioned by
multigrid V-cycles on the rank-local diagonal blocks.

HaloPlan / DistCsr are pure torch + torch.distributed, so the plan logic is
tested on CPU with gloo (the local SpMV kernel is injected).
"""
import numpy as np
import torch
import torch.distributed as dist


def strip_owner(n_global, world, device='cpu'):
    """owner[i] for contiguous, equal strips of a 1-D index range"""
    i = torch.arange(n_global, dtype=torch.int64, device=device)
    return torch.clamp((i * world) // max(n_global, 1), max=world - 1).to(torch.int32)


class HaloPlan:
    """Who sends which owned entries to whom so that every rank can fill the
    halo part of its extended vector [owned | halo].

    owner    : int32 tensor [n_global], rank owning each global index
    halo_ids : sorted unique global indices this rank needs but does not own
    Built with one exchange of request lists (the analogue of
    `MpiJacobian._find_to_and_from_ranks`, admpisolve.py:100-142, without its
    Iprobe polling loops).
    """

    def __init__(self, owner, halo_ids, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        dev = owner.device
        self.device = dev
        own = torch.nonzero(owner == self.rank).reshape(-1)
        self.own = own
        self.n_loc = int(own.numel())
        self.n_halo = int(halo_ids.numel())
        self.halo_ids = halo_ids
        # global -> local position among my owned entries
        lut = torch.full((owner.numel(),), -1, dtype=torch.int64, device=dev)
        lut[own] = torch.arange(self.n_loc, dtype=torch.int64, device=dev)
        src = owner[halo_ids].to(torch.int64) if self.n_halo else torch.zeros(0, dtype=torch.int64, device=dev)
        self.recv_slots, requests = {}, {}
        counts = torch.zeros(self.world, dtype=torch.int64)
        for q in range(self.world):
            if q == self.rank:
                continue
            sel = torch.nonzero(src == q).reshape(-1)
            if sel.numel():
                self.recv_slots[q] = sel
                requests[q] = halo_ids[sel].to(torch.int64).contiguous()
                counts[q] = sel.numel()
        counts = counts.to(dev)
        all_counts = [torch.zeros(self.world, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(all_counts, counts, group=group)
        all_counts = torch.stack(all_counts).cpu()
        ops, wanted = [], {}
        for q in range(self.world):
            if q == self.rank:
                continue
            want = int(all_counts[q][self.rank])        # q needs this many of my entries
            if want:
                wanted[q] = torch.empty(want, dtype=torch.int64, device=dev)
                ops.append(dist.P2POp(dist.irecv, wanted[q], q, group=group))
            if q in requests:
                ops.append(dist.P2POp(dist.isend, requests[q], q, group=group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # requested global ids -> my local positions
        self.send_idx = {}
        for q, ids in wanted.items():
            loc = lut[ids]
            assert bool((loc >= 0).all()), 'peer requested an entry this rank does not own'
            self.send_idx[q] = loc
        self.n_exchanges = 0

    def exchange(self, x_loc, x_ext=None):
        """x_ext = [x_loc | halo values fetched from their owners]"""
        if x_ext is None:
            x_ext = torch.empty(self.n_loc + self.n_halo, dtype=x_loc.dtype, device=x_loc.device)
        x_ext[:self.n_loc] = x_loc
        ops, recvs = [], []
        for q, idx in self.send_idx.items():
            ops.append(dist.P2POp(dist.isend, x_loc[idx], q, group=self.group))
        for q, slots in self.recv_slots.items():
            buf = torch.empty(slots.numel(), dtype=x_loc.dtype, device=x_loc.device)
            recvs.append((slots, buf))
            ops.append(dist.P2POp(dist.irecv, buf, q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for slots, buf in recvs:
            x_ext[self.n_loc + slots] = buf
        self.n_exchanges += 1
        return x_ext

    @property
    def bytes_per_exchange(self):
        return 8 * sum(int(v.numel()) for v in self.send_idx.values())


class DistCsr:
    """The rows of a global matrix owned by this rank, columns renumbered to
    [owned | halo] of the COLUMN space (what every rank of the reference's
    `diff_mpi` result holds as `{rank: block}`, admpi.py:226-229, merged into
    one local CSR).

    col_owner            : int32 tensor, owner of every global column
    indptr, indices, data: this rank's rows, GLOBAL column indices (torch)
    spmv(indptr, indices, data, x_ext, out, b, mode) -> out
                         : local kernel; mode 0: A x, 1: b - A x
    """

    def __init__(self, col_owner, indptr, indices, data, spmv, group=None):
        rank = dist.get_rank(group)
        dev = col_owner.device
        idx64 = indices.to(torch.int64)
        mine = col_owner[idx64] == rank
        halo_ids = torch.unique(idx64[~mine])
        self.plan = HaloPlan(col_owner, halo_ids, group=group)
        n_loc = self.plan.n_loc
        lut = torch.full((col_owner.numel(),), -1, dtype=torch.int32, device=dev)
        lut[self.plan.own] = torch.arange(n_loc, dtype=torch.int32, device=dev)
        lut[halo_ids] = n_loc + torch.arange(halo_ids.numel(), dtype=torch.int32, device=dev)
        self.indptr = indptr.to(torch.int32).contiguous()
        self.indices = lut[idx64].contiguous()
        self.data = data.to(torch.float64).contiguous()
        self.nrows = int(indptr.numel()) - 1
        self.nnz = int(indices.numel())
        self.n_loc, self.n_ext = n_loc, n_loc + int(halo_ids.numel())
        self._spmv = spmv
        self._x_ext = torch.empty(self.n_ext, dtype=torch.float64, device=dev)

    def matvec(self, x_loc, out=None, b=None, mode=0):
        x_ext = self.plan.exchange(x_loc, self._x_ext)
        if out is None:
            out = torch.empty(self.nrows, dtype=torch.float64, device=x_loc.device)
        return self._spmv(self.indptr, self.indices, self.data, x_ext, out, b, mode)

    def local_block(self):
        """(indptr, indices, data) of the owned x owned diagonal block, as
        boolean-mask filtered COO -- host-side helper for tests"""
        keep = self.indices < self.n_loc
        rows = torch.repeat_interleave(torch.arange(self.nrows, device=self.indices.device),
                                       (self.indptr[1:] - self.indptr[:-1]).to(torch.int64))
        return rows[keep], self.indices[keep], self.data[keep]


def allreduce_sum(t, group=None):
    """in-place sum over ranks of a small tensor of local inner products"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


# --------------------------------------------------------------------------- #
# GPU side: row-sharded LSC-preconditioned FGMRES
# --------------------------------------------------------------------------- #
def _cuda_spmv(indptr, indices, data, x_ext, out, b, mode):
    from ._dev import _call
    nrows = indptr.numel() - 1
    if mode == 0:
        _call('npb_spmv', nrows, indices.numel(), indptr, indices, data, x_ext, out)
    else:
        _call('npb_spmv_residual', nrows, indices.numel(), indptr, indices, data, x_ext, b, out)
    return out


def _shard_rows(M, rows, col_owner, group):
    """rows `rows` (device int32, increasing) of the replicated DevCsr M as a
    DistCsr over the column partition `col_owner`"""
    from ._dev import sel_jac, _sel_times_csr
    Mr = _sel_times_csr(sel_jac(rows, M.shape[0], full=True, monotone=True), M).materialized()
    return DistCsr(col_owner, Mr.indptr, Mr.indices, Mr.data, _cuda_spmv, group=group)


def _local_devcsr(D):
    """owned x owned block of a DistCsr as a DevCsr (for the local multigrid)"""
    from ._dev import DevCsr, sel_jac, _csr_times_sel
    n_loc = D.n_loc
    full = DevCsr((D.nrows, D.n_ext), D.indptr, D.indices, D.data, (), D.nnz, 64, True)
    m = torch.full((D.n_ext,), -1, dtype=torch.int32, device=D.indices.device)
    m[:n_loc] = torch.arange(n_loc, dtype=torch.int32, device=D.indices.device)
    return _csr_times_sel(full, sel_jac(m, n_loc, monotone=True, full=False)).materialized()


class DistLsc:
    """Row-sharded least-squares-commutator preconditioner (see csrc/amg.cu for
    the single-GPU form).  Every rank holds the replicated global J (assembled
    redundantly: 2 % of a Newton iteration) and extracts its strips."""

    def __init__(self, J, owner, group=None, cycles_a=2, cycles_l=3, theta_a=0.25):
        from . import amg
        from ._dev import device, _empty
        from ._lib import lib
        self.group = group
        rank = dist.get_rank(group)
        dev = device()
        J = J.materialized()
        n = J.shape[0]
        owner = owner.to(dev)
        d = amg.diagonal(J)
        has = d != 0
        vset = torch.nonzero(has).reshape(-1).to(torch.int32)
        pset = torch.nonzero(~has).reshape(-1).to(torch.int32)
        nv, np_ = int(vset.numel()), int(pset.numel())
        vmap = torch.full((n,), -1, dtype=torch.int32, device=dev)
        vmap[vset.long()] = torch.arange(nv, dtype=torch.int32, device=dev)
        pmap = torch.full((n,), -1, dtype=torch.int32, device=dev)
        pmap[pset.long()] = torch.arange(np_, dtype=torch.int32, device=dev)
        A = amg._submatrix(J, vset, vmap, nv)
        G = amg._submatrix(J, vset, pmap, np_)
        D = amg._submatrix(J, pset, vmap, nv)
        from ._dev import _csr_times_csr
        L = _csr_times_csr(D, G).scaled(-1.0).materialized()
        own_v_owner = owner[vset.long()]          # owner of every V-space index
        own_p_owner = owner[pset.long()]
        my_v = torch.nonzero(own_v_owner == rank).reshape(-1).to(torch.int32)
        my_p = torch.nonzero(own_p_owner == rank).reshape(-1).to(torch.int32)
        my_j = torch.nonzero(owner == rank).reshape(-1).to(torch.int32)
        self.A = _shard_rows(A, my_v, own_v_owner, group)
        self.G = _shard_rows(G, my_v, own_p_owner, group)
        self.D = _shard_rows(D, my_p, own_v_owner, group)
        self.L = _shard_rows(L, my_p, own_p_owner, group)
        self.J = _shard_rows(J, my_j, owner, group)
        self.amg_a = amg.Hierarchy(_local_devcsr(self.A), theta=theta_a, nu=2, smooth=True)
        self.amg_l = amg.Hierarchy(_local_devcsr(self.L), theta=0.0, nu=1, smooth=True)
        # positions of my V / P unknowns inside my strip of J's numbering
        jl = torch.full((n,), -1, dtype=torch.int64, device=dev)
        jl[my_j.long()] = torch.arange(my_j.numel(), device=dev)
        self.v_in_j = jl[vset.long()[my_v.long()]]
        self.p_in_j = jl[pset.long()[my_p.long()]]
        self.my_j = my_j
        self.nv_loc, self.np_loc, self.n_loc = int(my_v.numel()), int(my_p.numel()), int(my_j.numel())
        self.cycles_a, self.cycles_l = cycles_a, cycles_l
        z = lambda m: torch.zeros(max(m, 1), dtype=torch.float64, device=dev)
        self.wv = [z(self.nv_loc) for _ in range(5)]
        self.wp = [z(self.np_loc) for _ in range(7)]
        self.scal = torch.zeros(4, dtype=torch.float64, device=dev)
        self.red = torch.zeros(lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8, device=dev)

    # ---- sub-solves ---- #
    def _pcg_l(self, b, x):
        """k CG steps on the GLOBAL L, preconditioned by the local V-cycle"""
        from ._dev import _call
        r, z, p, q = self.wp[3], self.wp[4], self.wp[5], self.wp[6]
        n = self.np_loc
        r.copy_(b)
        for i in range(self.cycles_l):
            rz_new, rz_old = self.scal[i & 1:(i & 1) + 1], self.scal[(i + 1) & 1:((i + 1) & 1) + 1]
            self.amg_l.vcycle(r, z)
            _call('npb_dot_multi', n, 1, r, n, z, rz_new, self.red)
            allreduce_sum(rz_new, self.group)
            _call('npb_cg_dir', n, rz_new, rz_old, 1 if i == 0 else 0, z, p)
            self.L.matvec(p, out=q)
            pq = self.scal[2:3]
            _call('npb_dot_multi', n, 1, p, n, q, pq, self.red)
            allreduce_sum(pq, self.group)
            _call('npb_cg_update', n, rz_new, pq, p, q, x, r, 1 if i == 0 else 0)
        return x

    def _solve_a(self, b, x):
        """k stationary iterations on the GLOBAL A with the local V-cycle"""
        r, dx = self.wv[3], self.wv[4]
        self.amg_a.vcycle(b, x)
        for _ in range(1, self.cycles_a):
            self.A.matvec(x, out=r, b=b, mode=1)
            self.amg_a.vcycle(r, dx)
            x.add_(dx)
        return x

    def apply(self, r, z):
        """z = M^-1 r on this rank's strip (block upper-triangular LSC)"""
        rv, us, t1, t2 = self.wv[0], self.wv[1], self.wv[2], self.wv[3]
        rp, q, p1 = self.wp[0], self.wp[1], self.wp[2]
        torch.index_select(r, 0, self.v_in_j, out=rv[:self.nv_loc])
        if self.np_loc or True:
            torch.index_select(r, 0, self.p_in_j, out=rp[:self.np_loc])
            self._pcg_l(rp, p1)
            self.G.matvec(p1, out=t1)
            self.A.matvec(t1, out=t2)
            self.D.matvec(t2, out=q)
            self._pcg_l(q, p1)
            p1.neg_()
            self.G.matvec(p1, out=t1)
            rv.sub_(t1)
            z[self.p_in_j] = p1[:self.np_loc]
        self._solve_a(rv, us)
        z[self.v_in_j] = us[:self.nv_loc]


def solve_sharded(J, b, owner, group=None, rtol=1e-10, restart=100, maxiter=2000, **kw):
    """Solve J x = b with the rows of J (replicated DevCsr) sharded by `owner`
    across the ranks of `group`; returns the full solution on every rank and an
    info dict."""
    from . import linsolve
    from ._dev import device
    M = DistLsc(J, owner, group=group, **kw)
    dev = device()
    b_loc = b.reshape(-1)[M.my_j.long()].contiguous()
    n_loc = M.n_loc

    def matvec(x, y):
        M.J.matvec(x, out=y)

    def allreduce(ctx, buf, count, stream):
        t = linsolve._tensor_from_ptr(buf, count)
        allreduce_sum(t, group)
        return 0

    x_loc, info = linsolve.fgmres(None, b_loc, precond=M.apply, matvec=matvec, rtol=rtol,
                                  restart=restart, maxiter=maxiter, allreduce=allreduce)
    # gather the strips (padded to a common length) and scatter them into place
    world = dist.get_world_size(group)
    nmax = torch.tensor([n_loc], dtype=torch.int64, device=dev)
    dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
    nmax = int(nmax.item())
    xp = torch.zeros(nmax, dtype=torch.float64, device=dev)
    xp[:n_loc] = x_loc
    ip = torch.full((nmax,), -1, dtype=torch.int32, device=dev)
    ip[:n_loc] = M.my_j
    parts = [torch.empty_like(xp) for _ in range(world)]
    ids = [torch.empty_like(ip) for _ in range(world)]
    dist.all_gather(parts, xp, group=group)
    dist.all_gather(ids, ip, group=group)
    x = torch.empty(J.shape[0], dtype=torch.float64, device=dev)
    for i, p in zip(ids, parts):
        keep = i >= 0
        x[i[keep].long()] = p[keep]
    info['halo_exchanges'] = sum(m.plan.n_exchanges for m in (M.A, M.G, M.D, M.L, M.J))
    info['halo_bytes'] = M.J.plan.bytes_per_exchange
    return x, info
