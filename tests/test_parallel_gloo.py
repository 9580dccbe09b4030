"""World-size-2/3 gloo tests (CPU) of the multi-rank host logic: halo plans,
the row-sharded SpMV and the batched inner-product all-reduce of
numpad_b200/parallel.py.  The local SpMV kernel is injected (a scipy reference
here; the CUDA kernel in the product), so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_spmv(indptr, indices, data, x_ext, out, b, mode):
    a = sp.csr_matrix((data.numpy(), indices.numpy(), indptr.numpy()),
                      shape=(indptr.numel() - 1, x_ext.numel()))
    y = torch.from_numpy(a @ x_ext.numpy())
    out.copy_(y if mode == 0 else b - y)
    return out


def _worker(rank, world, port, n, seed, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from numpad_b200 import parallel
        rng = np.random.RandomState(seed)
        # banded + a few long-range couplings, same matrix on every rank
        a = sp.diags([rng.standard_normal(n - 2), rng.standard_normal(n - 1), 3 + rng.random(n),
                      rng.standard_normal(n - 1), rng.standard_normal(n - 3)],
                     [-2, -1, 0, 1, 3], format='lil')
        for _ in range(6):
            a[rng.randint(n), rng.randint(n)] = rng.standard_normal()
        a = sp.csr_matrix(a)
        a.sort_indices()
        x = rng.standard_normal(n)
        # interleaved ownership on purpose: not just contiguous strips
        owner = parallel.strip_owner(n, world)
        owner[::7] = (owner[::7] + 1) % world
        own = torch.nonzero(owner == rank).reshape(-1).numpy()
        rows = a[own]
        D = parallel.DistCsr(owner, torch.from_numpy(rows.indptr.astype(np.int32)),
                             torch.from_numpy(rows.indices.astype(np.int32)),
                             torch.from_numpy(rows.data), _cpu_spmv)
        x_loc = torch.from_numpy(x[own])
        y = D.matvec(x_loc)
        ref = (a @ x)[own]
        assert np.allclose(y.numpy(), ref, rtol=1e-13, atol=1e-13), 'dist spmv'
        b_loc = torch.from_numpy(rng.standard_normal(own.size))
        r = D.matvec(x_loc, b=b_loc, mode=1)
        assert np.allclose(r.numpy(), b_loc.numpy() - ref, rtol=1e-13, atol=1e-13)
        # halo accounting: every halo slot is filled from exactly one peer
        slots = torch.cat([v for v in D.plan.recv_slots.values()]) if D.plan.recv_slots else torch.zeros(0)
        assert slots.numel() == D.plan.n_halo == len(set(slots.tolist()))
        # batched inner products: one all-reduce for several dots
        v = torch.tensor([float(x_loc @ x_loc), float(x_loc.sum()), 1.0], dtype=torch.float64)
        parallel.allreduce_sum(v)
        assert abs(v[0].item() - x @ x) < 1e-10 and abs(v[1].item() - x.sum()) < 1e-10
        assert v[2].item() == world
        # the diagonal block seen by the block-Jacobi preconditioner
        rr, cc, vv = D.local_block()
        blk = sp.csr_matrix((vv.numpy(), (rr.numpy(), cc.numpy())), shape=(own.size, own.size))
        assert abs(blk - a[own][:, own]).max() < 1e-15
        ret[rank] = 'ok'
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_halo_plan_and_sharded_spmv(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 61, 3, ret), nprocs=world, join=True)
    assert dict(ret) == {r: 'ok' for r in range(world)}


def test_strip_owner():
    from numpad_b200 import parallel
    o = parallel.strip_owner(10, 3)
    assert o.tolist() == [0, 0, 0, 0, 1, 1, 1, 2, 2, 2]
    assert parallel.strip_owner(5, 8).max().item() <= 7
