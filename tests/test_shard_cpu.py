"""Host logic of the row-sharded solve (numpad_b200/shard.py) on CPU: the slicing of a
replicated operator into local rows with the [owned | boundary slots of every rank]
column numbering of npb_halo, checked by simulating the all-gather, and the strip
permutation.  The device path is exercised by tools/shard_check.py under torchrun."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from numpad_b200 import shard


def _rand(m, n, density, seed):
    rng = np.random.RandomState(seed)
    a = sp.random(m, n, density=density, random_state=rng, format='csr')
    a.data = rng.standard_normal(a.nnz)
    a.sort_indices()
    return a


def _cuts(n, world, rng):
    c = np.sort(rng.choice(np.arange(1, n), world - 1, replace=False)) if world > 1 else np.array([], int)
    return [0] + [int(v) for v in c] + [n]


@pytest.mark.parametrize('world', [1, 2, 3, 5])
@pytest.mark.parametrize('shape', [(60, 60), (90, 40), (35, 120)])
def test_localize_matches_global_product(world, shape):
    rng = np.random.RandomState(world * 7 + shape[0])
    a = _rand(shape[0], shape[1], 0.15, world + shape[1])
    rc, cc = _cuts(shape[0], world, rng), _cuts(shape[1], world, rng)
    if shape[0] == shape[1]:
        cc = rc
    x = rng.standard_normal(shape[1])
    ptr = torch.from_numpy(a.indptr.astype(np.int32))
    idx = torch.from_numpy(a.indices.astype(np.int32))
    dat = torch.from_numpy(a.data)
    ops = [shard.localize(ptr, idx, dat, shape, rc, cc, r, world) for r in range(world)]
    maxb = ops[0].maxb
    assert all(o.maxb == maxb for o in ops)
    # the all-gather: rank q publishes x_own[send_idx] in its slot of world * maxb values
    gathered = torch.zeros(world * maxb, dtype=torch.float64)
    for q, o in enumerate(ops):
        xq = torch.from_numpy(x[cc[q]:cc[q + 1]])
        assert o.send_idx.numel() <= maxb
        gathered[q * maxb:q * maxb + o.send_idx.numel()] = xq[o.send_idx.long()]
    want = a @ x
    for r, o in enumerate(ops):
        assert o.n_own == cc[r + 1] - cc[r]
        assert o.shape == (rc[r + 1] - rc[r], o.n_own + world * maxb)
        y = shard.local_matvec_reference(o, torch.from_numpy(x[cc[r]:cc[r + 1]]), gathered)
        np.testing.assert_allclose(y.numpy(), want[rc[r]:rc[r + 1]], rtol=0, atol=1e-13)
        # a rank never reads its own slot, owned columns stay below n_own
        own_slot = (o.indices >= o.n_own + r * maxb) & (o.indices < o.n_own + (r + 1) * maxb)
        assert not bool(own_slot.any())


def test_localize_replicated_columns():
    a = _rand(50, 30, 0.2, 3)
    rc = [0, 20, 50]
    x = np.random.RandomState(4).standard_normal(30)
    ptr = torch.from_numpy(a.indptr.astype(np.int32))
    idx = torch.from_numpy(a.indices.astype(np.int32))
    dat = torch.from_numpy(a.data)
    for r in range(2):
        o = shard.localize(ptr, idx, dat, (50, 30), rc, None, r, 2, replicated_cols=True)
        assert o.maxb == 0 and o.n_own == 30 and not o.sharded_cols
        y = shard.local_matvec_reference(o, torch.from_numpy(x), torch.zeros(0, dtype=torch.float64))
        np.testing.assert_allclose(y.numpy(), (a @ x)[rc[r]:rc[r + 1]], rtol=0, atol=1e-13)


def test_strip_permutation_of_the_cavity_partition():
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(None, 12)
    world = 3
    owner = P.partition(world)
    perm, inv, cuts = shard.strip_permutation(torch.from_numpy(owner), world)
    assert cuts[0] == 0 and cuts[-1] == P.n and len(cuts) == world + 1
    assert sorted(perm.tolist()) == list(range(P.n))
    assert bool((inv[perm] == torch.arange(P.n)).all())
    for r in range(world):
        assert (owner[perm[cuts[r]:cuts[r + 1]].numpy()] == r).all()
    # stable: the original order is kept inside a strip
    for r in range(world):
        seg = perm[cuts[r]:cuts[r + 1]]
        assert bool((seg[1:] > seg[:-1]).all())


# --------------------------------------------------------------------------- #
# world-size-2/3 gloo: the same numbering with a REAL all-gather between processes
# --------------------------------------------------------------------------- #
def _free_port():
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, ret):
    import os
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from numpad_b200.workloads.cavity import CavityProblem
        # a strip-permuted cavity-like operator: 2-D 9-point stencil on 3 interleaved fields
        n1 = 14
        T = sp.diags([-1.0, 2.5, -1.2], [-1, 0, 1], shape=(n1, n1))
        a = sp.kron(sp.identity(n1), T) + sp.kron(T, sp.identity(n1)) + 0.1 * sp.kron(T, T)
        a = sp.csr_matrix(sp.kron(np.array([[1, .2, 0], [.1, 1, .3], [0, .4, 1]]), a))
        a.sort_indices()
        n = a.shape[0]
        owner = torch.from_numpy(np.tile((np.arange(n1 * n1) // n1 * world) // n1, 3).astype(np.int32))
        perm, inv, cuts = shard.strip_permutation(owner, world)
        ap = a[perm.numpy()][:, perm.numpy()].tocsr()
        ap.sort_indices()
        x = np.random.RandomState(5).standard_normal(n)
        op = shard.localize(torch.from_numpy(ap.indptr.astype(np.int32)),
                            torch.from_numpy(ap.indices.astype(np.int32)),
                            torch.from_numpy(ap.data), ap.shape, cuts, cuts, rank, world)
        x_own = torch.from_numpy(x[cuts[rank]:cuts[rank + 1]])
        mine = torch.zeros(max(op.maxb, 1), dtype=torch.float64)
        mine[:op.send_idx.numel()] = x_own[op.send_idx.long()]
        slots = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(slots, mine)                    # what ncclAllGather / the direct stores do
        gathered = torch.cat([s[:op.maxb] for s in slots]) if op.maxb else torch.zeros(0, dtype=torch.float64)
        y = shard.local_matvec_reference(op, x_own, gathered)
        want = (ap @ x)[cuts[rank]:cuts[rank + 1]]
        assert np.allclose(y.numpy(), want, rtol=0, atol=1e-12), 'sharded SpMV'
        # inner product: local partial sums + all-reduce == global dot
        part = torch.tensor([float(y @ y)], dtype=torch.float64)
        dist.all_reduce(part)
        full = ap @ x
        assert abs(part.item() - full @ full) <= 1e-10 * (full @ full), 'all-reduced dot'
        ret[rank] = 'ok'
    except Exception as e:                              # surface the failure in the parent
        ret[rank] = repr(e)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_spmv_over_gloo(world):
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ['ok'] * world, dict(ret)
