"""Host-side (pure torch) pieces of the multigrid setup, on CPU."""
import numpy as np
import scipy.sparse as sp
import torch
from scipy.sparse.csgraph import connected_components


def _greedy_mis1(a, rng):
    """serial distance-1 MIS aggregation (roots + strongest root neighbour), ids in root order"""
    n = a.shape[0]
    pri = rng.permutation(n)
    state = np.zeros(n, int)
    for i in np.argsort(-pri):
        if state[i] == 0:
            state[i] = 1
            nb = a.indices[a.indptr[i]:a.indptr[i + 1]]
            state[nb[state[nb] == 0]] = 2
    roots = np.nonzero(state == 1)[0]
    rid = np.concatenate([[0], np.cumsum(state == 1)]).astype(np.int32)
    agg = np.full(n, -1)
    agg[roots] = np.arange(roots.size)
    for i in np.nonzero(state != 1)[0]:
        nb = a.indices[a.indptr[i]:a.indptr[i + 1]]
        w = np.abs(a.data[a.indptr[i]:a.indptr[i + 1]])
        ok = state[nb] == 1
        agg[i] = agg[nb[ok][np.argmax(w[ok])]]
    return agg, roots.size, rid


def test_merge_small_aggregates_properties():
    """amg.merge_small_aggregates on a 2-D 5-point operator: every node keeps an aggregate, ids
    are dense and follow the order of the surviving roots (what shard.py cuts by), aggregates
    stay connected, singletons and pairs mostly disappear (mean size 2.75 -> > 3.8)"""
    from numpad_b200 import amg
    n1 = 48
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n1, n1))
    a = sp.csr_matrix(sp.kron(sp.identity(n1), T) + sp.kron(T, sp.identity(n1)))
    a.sort_indices()
    agg, na, rid = _greedy_mis1(a, np.random.RandomState(1))
    t = lambda v, dt: torch.from_numpy(np.ascontiguousarray(v)).to(dt)
    g, n2, rid2 = amg.merge_small_aggregates(t(a.indptr, torch.int32), t(a.indices, torch.int32),
                                             t(a.data, torch.float64), t(agg, torch.int32), na,
                                             t(rid, torch.int32))
    g = g.numpy()
    n = a.shape[0]
    assert g.min() == 0 and g.max() == n2 - 1 and (np.bincount(g, minlength=n2) > 0).all()
    assert n / na < 3.3 and n / n2 > 3.8
    coo = a.tocoo()
    same = g[coo.row] == g[coo.col]
    sub = sp.csr_matrix((np.ones(same.sum()), (coo.row[same], coo.col[same])), shape=a.shape)
    assert connected_components(sub, directed=False)[0] == n2
    roots = np.nonzero(np.diff(rid2.numpy()))[0]
    assert roots.size == n2 and np.array_equal(g[roots], np.arange(n2))
    # merging is a coarsening of the input aggregation
    for old in range(na):
        assert np.unique(g[agg == old]).size == 1
