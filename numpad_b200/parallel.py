"""Row-sharded sparse operators across the GPUs of one node (seam S4).

One process per GPU (torchrun); `torch.distributed` (NCCL over NVLink on the
GPUs, gloo in the CPU tests) carries exactly two kinds of traffic, as the
reference's distributed solver does with mpi4py:

  * halo exchange of the operand before the boundary part of an SpMV
    -- replaces the partial-product exchange of `MpiJacobian.matvec`
       (admpisolve.py:146-165): x halos are exchanged instead of partial y,
       which makes the CSR->CSC ownership swap (`_csr_to_csc`, :61-98) and the
       pickled block shipping unnecessary;
  * one all-reduce of a few doubles per batch of inner products
    -- replaces one scalar `Allreduce` per dot in `_dot` / `_norm2`
       (admpisolve.py:177-188).

Everything here is host-side plan logic + collectives; the arithmetic is the
CUDA SpMV of the C-ABI (injected as `spmv`, so the plan logic is testable on
CPU with gloo).
"""
import numpy as np
import torch
import torch.distributed as dist


def strip_partition(n_global, world):
    """contiguous equal strips of a 1-D index range: owner[i]"""
    bounds = [(n_global * r) // world for r in range(world + 1)]
    owner = np.empty(n_global, np.int32)
    for r in range(world):
        owner[bounds[r]:bounds[r + 1]] = r
    return owner


class HaloPlan:
    """Who sends which owned entries to whom so that every rank can fill the
    halo part of its extended vector [owned | halo].

    owner      : int array [n_global], rank owning each global index
    halo_ids   : sorted global indices this rank needs but does not own
    Built with one exchange of request lists (the analogue of
    `MpiJacobian._find_to_and_from_ranks`, admpisolve.py:100-142, without its
    Iprobe polling loops).
    """

    def __init__(self, owner, halo_ids, rank=None, world=None, group=None, device='cpu'):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.device = torch.device(device)
        owner = np.asarray(owner)
        self.own = np.nonzero(owner == self.rank)[0]
        self.n_loc = int(self.own.size)
        halo_ids = np.asarray(halo_ids, np.int64)
        assert halo_ids.size == 0 or (owner[halo_ids] != self.rank).all()
        self.halo_ids = halo_ids
        self.n_halo = int(halo_ids.size)
        # global id -> position in its owner's owned list
        pos_in_owner = np.empty(owner.size, np.int64)
        for r in range(self.world):
            ids = np.nonzero(owner == r)[0]
            pos_in_owner[ids] = np.arange(ids.size)
        src = owner[halo_ids] if halo_ids.size else np.zeros(0, np.int32)
        # halo slots grouped by source rank (halo_ids sorted => stable grouping)
        self.recv_slots = {}
        requests = {}
        for q in range(self.world):
            if q == self.rank:
                continue
            sel = np.nonzero(src == q)[0]
            if sel.size:
                self.recv_slots[q] = torch.from_numpy(sel.astype(np.int64)).to(self.device)
                requests[q] = pos_in_owner[halo_ids[sel]]
        # tell every peer which of ITS entries I need
        counts = torch.zeros(self.world, dtype=torch.int64)
        for q, ids in requests.items():
            counts[q] = ids.size
        all_counts = [torch.zeros(self.world, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(all_counts, counts, group=group)
        ops, recv_bufs = [], {}
        for q in range(self.world):
            if q == self.rank:
                continue
            want = int(all_counts[q][self.rank])       # q needs this many of my entries
            if want:
                recv_bufs[q] = torch.empty(want, dtype=torch.int64)
                ops.append(dist.P2POp(dist.irecv, recv_bufs[q], q, group=group))
            if q in requests:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(requests[q].copy()), q,
                                      group=group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self.send_idx = {q: b.to(self.device) for q, b in recv_bufs.items()}

    def exchange(self, x_loc, x_ext=None):
        """x_ext = [x_loc | halo values fetched from their owners]"""
        if x_ext is None:
            x_ext = torch.empty(self.n_loc + self.n_halo, dtype=x_loc.dtype, device=x_loc.device)
        x_ext[:self.n_loc] = x_loc
        ops, recvs = [], []
        for q, idx in self.send_idx.items():
            ops.append(dist.P2POp(dist.isend, x_loc[idx].contiguous(), q, group=self.group))
        for q, slots in self.recv_slots.items():
            buf = torch.empty(slots.numel(), dtype=x_loc.dtype, device=x_loc.device)
            recvs.append((slots, buf))
            ops.append(dist.P2POp(dist.irecv, buf, q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for slots, buf in recvs:
            x_ext[self.n_loc + slots] = buf
        return x_ext

    @property
    def bytes_per_exchange(self):
        return 8 * sum(int(v.numel()) for v in self.send_idx.values())


class DistCsr:
    """The rows of a global square matrix owned by this rank, with columns
    renumbered to [owned | halo] (what every rank of the reference's
    `diff_mpi` result holds as `{rank: block}`, admpi.py:226-229, merged into
    one local CSR).

    indptr/indices/data : this rank's rows with GLOBAL column indices (numpy
                          or torch, any device)
    spmv(indptr, indices, data, x_ext) -> y : local kernel (CUDA SpMV in the
                          product; injected so CPU tests can pass a reference)
    """

    def __init__(self, owner, indptr, indices, data, spmv, group=None, device='cpu'):
        self.group = group
        rank = dist.get_rank(group)
        owner = np.asarray(owner)
        indices_np = indices.cpu().numpy() if isinstance(indices, torch.Tensor) else np.asarray(indices)
        mine = owner[indices_np] == rank
        halo_ids = np.unique(indices_np[~mine])
        self.plan = HaloPlan(owner, halo_ids, group=group, device=device)
        n_loc = self.plan.n_loc
        # global column -> extended local column
        lut = np.full(owner.size, -1, np.int64)
        lut[self.plan.own] = np.arange(n_loc)
        lut[halo_ids] = n_loc + np.arange(halo_ids.size)
        dev = torch.device(device)
        as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))).to(dev, dt)
        self.indptr = as_t(indptr, torch.int32)
        self.indices = torch.from_numpy(lut[indices_np].astype(np.int32)).to(dev)
        self.data = as_t(data, torch.float64)
        self.n_loc, self.n_ext = n_loc, n_loc + int(halo_ids.size)
        self._spmv = spmv
        self._x_ext = torch.empty(self.n_ext, dtype=torch.float64, device=dev)

    def matvec(self, x_loc, out=None):
        x_ext = self.plan.exchange(x_loc, self._x_ext)
        return self._spmv(self.indptr, self.indices, self.data, x_ext, out)


def dot_allreduce(values, group=None):
    """sum the small vector of local inner products over all ranks (in place)"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group)
    return values
