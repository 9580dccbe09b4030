"""The row-sharded solve path (numpad_b200/shard.py + csrc/dist.cu) on ONE GPU: a world of
one rank runs the same code -- library NCCL communicator, sliced operators, sharded V-cycles
with replicated coarse levels, FGMRES with the all-reduce hook -- and must reproduce the
plain single-GPU solve.  Multi-rank runs: tools/shard_check.py under torchrun."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_world_of_one_matches_single_gpu_solve():
    import torch
    import torch.distributed as dist
    import numpad_b200 as npd
    from numpad_b200 import linsolve, shard, precond
    from numpad_b200.workloads.cavity import CavityProblem
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    created = not dist.is_initialized()
    if created:
        dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=0, world_size=1,
                                device_id=torch.device('cuda', torch.cuda.current_device()))
    try:
        P = CavityProblem(npd, 160)
        u0 = P.fixed_state()
        u = npd.array(u0.copy())
        res = P.residual(u, npd.array(u0), 1.0, 0)
        J = npd.diff_dev(res, u).materialized()
        b = res._dev.reshape(-1).contiguous()
        precond._last['lsc'] = None
        linsolve.reset_hints()        # an earlier test may have switched to the robust inner solver
        x1 = linsolve.solve(J, b, rtol=1e-11, precond='lsc', restart=100)
        its1 = linsolve.last_info['iters']
        old = shard.REPLICATE_BELOW
        shard.REPLICATE_BELOW = 5000          # shard two levels, replicate the rest
        try:
            x2 = shard.solve(J, b, owner=P.partition(1), rtol=1e-11, restart=100)
        finally:
            shard.REPLICATE_BELOW = old
        info = dict(shard.last_info)
        assert info['first_replicated_level'][1] >= 2
        rel = float((b - J.matvec(x2)).norm() / b.norm())
        assert rel < 5e-11
        assert abs(info['iters'] - its1) <= 2
        assert float((x2 - x1).norm() / x1.norm()) < 1e-8
    finally:
        shard.shutdown()
        if created:
            dist.destroy_process_group()


def test_symmetric_permutation_matches_scipy():
    """shard.permuted (row gather + column relabel + per-row sort) == P A P^T of scipy"""
    import torch
    import scipy.sparse as sp
    from numpad_b200 import shard, _dev
    rng = np.random.RandomState(3)
    n = 5000
    a = sp.random(n, n, density=0.002, random_state=rng, format='csr') + sp.identity(n, format='csr')
    a = sp.csr_matrix(a); a.sort_indices()
    owner = torch.from_numpy(rng.randint(0, 3, n).astype(np.int32)).cuda()
    perm, inv, cuts = shard.strip_permutation(owner, 3)
    got = shard.permuted(_dev.DevCsr.from_scipy(a), perm, inv).toscipy()
    p = perm.cpu().numpy()
    want = a[p][:, p].tocsr(); want.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
