"""Kernel-level parity: each C-ABI family against scipy on seeded random
matrices (bit-exact patterns, values to 1e-15 relative -- the kernels use the
same non-fused multiply/add sequence as scipy's loops)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    import torch
    from numpad_b200 import _dev
    _dev.device()
    return _dev


def rand_csr(rng, m, n, density, zero_frac=0.0):
    a = sp.random(m, n, density=density, random_state=rng, format='csr')
    a.data = rng.standard_normal(a.nnz)
    if zero_frac:
        a.data[rng.random(a.nnz) < zero_frac] = 0.0
    a.sort_indices()
    return a


def check(got, want, rtol=1e-14):
    from conftest import assert_csr_parity
    g = got.toscipy()
    assert g.has_canonical_format or True
    assert_csr_parity(g, want, rtol=rtol, raw_nnz=sp.csr_matrix(want).nnz)
    # device result must already be canonical (sorted, no duplicates)
    gi = g.copy(); gi.sort_indices()
    assert np.array_equal(gi.indices, g.indices)


@pytest.mark.parametrize('shape', [(1, 1), (7, 5), (300, 300), (5000, 4000), (3, 20000)])
def test_roundtrip_and_spmv(dev, shape):
    import torch
    rng = np.random.RandomState(0)
    a = rand_csr(rng, shape[0], shape[1], min(1.0, 8.0 / shape[1]))
    A = dev.DevCsr.from_scipy(a)
    check(A, a, 0.0)
    x = rng.standard_normal(shape[1])
    y = A.matvec(torch.from_numpy(x).to(dev.device())).cpu().numpy()
    np.testing.assert_allclose(y, a @ x, rtol=1e-13, atol=1e-13 * np.abs(a @ x).max() if a.nnz else 0)
    check(A.scaled(-2.5).scaled(0.3), (a * -2.5) * 0.3, 0.0)
    check(A.transpose(), a.T.tocsr(), 0.0)


@pytest.mark.parametrize('shape', [(1, 1), (50, 70), (4000, 4000), (2, 9000)])
@pytest.mark.parametrize('mode', ['plain', 'cancel'])
def test_spadd(dev, shape, mode):
    rng = np.random.RandomState(1)
    d = min(1.0, 6.0 / shape[1])
    a = rand_csr(rng, *shape, d, zero_frac=0.1)
    if mode == 'cancel':        # b cancels a exactly on a's pattern, plus extras
        b = ((-a) + rand_csr(rng, *shape, d / 2)).tocsr()
        b.sort_indices()
    else:
        b = rand_csr(rng, *shape, d, zero_frac=0.1)
    A, B = dev.DevCsr.from_scipy(a), dev.DevCsr.from_scipy(b)
    check(dev._add_ops(A, B), a + b, 0.0)
    check(dev._add_ops(A.scaled(2.0), B.scaled(-0.5).scaled(3.0)), (a * 2.0) + ((b * -0.5) * 3.0), 0.0)
    # exact cancellation: A + (-1)A must vanish entirely
    c = dev._add_ops(A, A.scaled(-1))
    assert c.nnz == 0


def test_spadd_long_rows(dev):
    rng = np.random.RandomState(2)
    a = rand_csr(rng, 3, 50000, 0.5)
    b = rand_csr(rng, 3, 50000, 0.5)
    check(dev._add_ops(dev.DevCsr.from_scipy(a), dev.DevCsr.from_scipy(b)), a + b, 0.0)


def test_diag_products(dev):
    import torch
    rng = np.random.RandomState(3)
    a = rand_csr(rng, 600, 500, 0.02, zero_frac=0.1)
    A = dev.DevCsr.from_scipy(a).scaled(1.7)
    dc = rng.standard_normal(500); dc[::7] = 0.0
    dr = rng.standard_normal(600)
    # dia_jac.tocsr keeps explicit zeros (reference adstate.py:244-250)
    dia = lambda d: sp.csr_matrix((d, np.arange(d.size), np.arange(d.size + 1)))
    check(dev._multiply_ops(A, dev.dia_jac(dc)), (a * 1.7) * dia(dc), 0.0)
    check(dev._multiply_ops(dev.dia_jac(dr), A), dia(dr) * (a * 1.7), 0.0)


def sel_scipy(map_, ncols, w=None):
    rows = np.nonzero(map_ >= 0)[0]
    data = np.ones(rows.size) if w is None else w[rows]
    return sp.csr_matrix((data, (rows, map_[rows])), shape=(map_.size, ncols))


@pytest.mark.parametrize('kind', ['slice', 'scatter', 'perm', 'bcast'])
@pytest.mark.parametrize('weights', [False, True])
def test_selection_products(dev, kind, weights):
    import torch
    rng = np.random.RandomState(4)
    k, ncols = 400, 300
    if kind == 'slice':
        k = 200; m = np.arange(50, 250)
    elif kind == 'scatter':
        m = -np.ones(k, int); rows = np.sort(rng.choice(k, 150, replace=False)); m[rows] = np.sort(rng.choice(ncols, 150, replace=False))
    elif kind == 'perm':
        ncols = k; m = rng.permutation(k)
    else:
        m = rng.randint(0, ncols, k)
    w = rng.standard_normal(k) if weights else None
    S = dev.sel_jac(torch.from_numpy(m.astype(np.int32)).to(dev.device()), ncols,
                    weights=None if w is None else torch.from_numpy(w).to(dev.device()))
    s = sel_scipy(m, ncols, w)
    check(S.todev(), s, 0.0)
    a = rand_csr(rng, 350, k, 0.03, zero_frac=0.1)
    A = dev.DevCsr.from_scipy(a).scaled(0.25)
    check(dev._multiply_ops(A, S), (a * 0.25) * s, 1e-15)           # adjoint order
    b = rand_csr(rng, ncols, 280, 0.03, zero_frac=0.1)
    B = dev.DevCsr.from_scipy(b).scaled(-3.0)
    check(dev._multiply_ops(S, B), s * (b * -3.0), 1e-15)           # tangent order
    # identity seeds
    check(dev._multiply_ops(dev.DevEye(k), S), s, 0.0)
    check(dev._multiply_ops(S, dev.DevEye(ncols, 2.0)), s * 2.0, 0.0)


def test_general_product_and_coo(dev):
    import torch
    rng = np.random.RandomState(5)
    a = rand_csr(rng, 300, 200, 0.04, zero_frac=0.1)
    b = rand_csr(rng, 200, 250, 0.04, zero_frac=0.1)
    got = dev._multiply_ops(dev.DevCsr.from_scipy(a), dev.csr_jac(b.tocoo().data, b.tocoo().row, b.tocoo().col, b.shape))
    check(got, a * b, 1e-14)
    # COO with duplicates: summed, explicit zeros kept
    i = rng.randint(0, 40, 500); j = rng.randint(0, 30, 500); v = rng.standard_normal(500); v[::9] = 0
    want = sp.csr_matrix((v, (i, j)), shape=(40, 30))
    check(dev.csr_jac(v, i, j, (40, 30)).todev(), want, 1e-14)
    # inferred shape
    assert dev.csr_jac(v, i, j).shape == (i.max() + 1, j.max() + 1)


def test_zero_rules(dev):
    """adstate.py:282-297 / SURVEY App. B: int 0 annihilates, 0.0 does not;
    an all-zero dia seed converts to an empty matrix"""
    import torch
    A = dev.DevCsr.from_scipy(sp.identity(4, format='csr'))
    assert dev._multiply_ops(0, A) == 0 and dev._multiply_ops(A, 0) == 0
    z = dev._multiply_ops(A, 0.0)
    assert isinstance(z, dev.DevCsr) and z.nnz == 4
    assert dev._add_ops(0, A) is A and dev._add_ops(A, 0) is A
    S = dev.sel_jac(torch.arange(4, dtype=torch.int32, device=dev.device()), 4)
    assert dev._multiply_ops(dev.DevEye(4, 0.0), S) == 0
    e = dev._add_ops(dev.DevEye(4, 3.0), dev.DevEye(4, -3.0))
    assert isinstance(e, dev.DevEye) and e.val == 0.0


def test_fgmres_poisson(dev):
    import torch
    from numpad_b200 import linsolve
    n = 2000
    a = sp.diags([-1, 2.2, -1.1], [-1, 0, 1], shape=(n, n)).tocsr()
    rng = np.random.RandomState(6)
    x = rng.standard_normal(n)
    b = torch.from_numpy(a @ x).to(dev.device())
    A = dev.DevCsr.from_scipy(a)
    got, info = linsolve.fgmres(A, b, rtol=1e-12, restart=80, maxiter=4000)
    assert info['status'] == 0, info
    np.testing.assert_allclose(got.cpu().numpy(), x, rtol=0, atol=1e-8)
    # python-callback preconditioner (Jacobi) and python matvec
    dinv = torch.from_numpy(1.0 / a.diagonal()).to(dev.device())
    got2, info2 = linsolve.fgmres(A, b, precond=lambda r, z: z.copy_(r * dinv), rtol=1e-12,
                                  restart=80, maxiter=4000)
    assert info2['status'] == 0
    np.testing.assert_allclose(got2.cpu().numpy(), x, rtol=0, atol=1e-8)
    got3, info3 = linsolve.fgmres(None, b, matvec=lambda v, y: A.matvec(v, out=y), rtol=1e-12,
                                  restart=80, maxiter=4000)
    assert info3['status'] == 0 and info3['iters'] == info['iters']


@pytest.mark.parametrize('n,density', [(1, 1.0), (40, 0.2), (700, 0.01), (3000, 0.002)])
def test_band_lu(dev, n, density):
    """pivoted banded LU after RCM vs scipy spsolve, incl. zero diagonals"""
    import torch
    import scipy.sparse.linalg as spla
    from numpad_b200 import precond, linsolve
    rng = np.random.RandomState(7)
    a = rand_csr(rng, n, n, density) + sp.diags(rng.standard_normal(n) * 3).tocsr()
    a = a.tolil()
    if n > 10:                      # saddle-like: some zero diagonals, needs pivoting
        for i in range(0, n - 1, 7):
            a[i, i] = 0.0
            a[i, i + 1] = 1.5
            a[i + 1, i] = -2.0
    a = sp.csr_matrix(a)
    a.eliminate_zeros()
    x = rng.standard_normal(n)
    b = a @ x
    A = dev.DevCsr.from_scipy(a)
    lu = precond.BandLU(A)
    got = lu.solve(torch.from_numpy(b).to(dev.device())).cpu().numpy()
    ref = spla.spsolve(a.tocsc(), b) if n > 1 else b / a.toarray()[0, 0]
    scale = np.abs(ref).max()
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9 * scale)
    # through the front door (direct + refinement), and the transposed system
    got2 = linsolve.solve(A, torch.from_numpy(b).to(dev.device()), precond='direct').cpu().numpy()
    np.testing.assert_allclose(got2, ref, rtol=0, atol=1e-10 * scale)
    bt = a.T @ x
    got3 = linsolve.solve(A, torch.from_numpy(bt).to(dev.device()), transpose=True,
                          precond='direct').cpu().numpy()
    np.testing.assert_allclose(got3, x, rtol=0, atol=1e-8 * np.abs(x).max())


def poisson2d(n1):
    T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n1, n1))
    return sp.csr_matrix(sp.kron(sp.identity(n1), T) + sp.kron(T, sp.identity(n1)))


def test_amg_poisson(dev):
    """smoothed-aggregation V-cycle as FGMRES preconditioner: mesh-independent"""
    import torch
    from numpad_b200 import amg, linsolve
    its = []
    for n1 in (48, 96, 192):
        a = poisson2d(n1)
        A = dev.DevCsr.from_scipy(a)
        M = amg.AmgPreconditioner(A, theta=0.0, nu=1)
        rng = np.random.RandomState(8)
        x = rng.standard_normal(a.shape[0])
        b = torch.from_numpy(a @ x).to(dev.device())
        got, info = linsolve.fgmres(A, b, precond=M, rtol=1e-10, restart=50, maxiter=200)
        assert info['status'] == 0, (n1, info, M.h.describe())
        np.testing.assert_allclose(got.cpu().numpy(), x, rtol=0, atol=1e-6 * np.abs(x).max())
        its.append(info['iters'])
        assert M.h.struct.coarse_n > 0 and len(M.h.sizes) >= 2
    print('amg poisson its', its)
    assert its[-1] <= 40 and its[-1] <= its[0] + 15


def test_dense_inverse(dev):
    import torch
    rng = np.random.RandomState(9)
    n = 150
    a = rand_csr(rng, n, n, 0.1) + sp.identity(n, format='csr') * 0.05
    a = sp.csr_matrix(a); a.sort_indices()
    A = dev.DevCsr.from_scipy(a)
    W = torch.empty(2 * n * n, dtype=torch.float64, device=dev.device())
    Minv = torch.empty(n * n, dtype=torch.float64, device=dev.device())
    info = torch.zeros(1, dtype=torch.int32, device=dev.device())
    dev._call('npb_dense_inverse', n, A.indptr, A.indices, A.data, W, Minv, info)
    assert int(info.item()) == 0
    got = Minv.cpu().numpy().reshape(n, n)
    ref = np.linalg.inv(a.toarray())
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9 * np.abs(ref).max())


@pytest.mark.parametrize('shape,density', [((1, 1), 1.0), ((257, 300), 0.03), ((5000, 5000), 0.002),
                                           ((3, 20000), 0.5), ((40001, 3000), 0.003)])
def test_spmv_variants_agree(dev, shape, density):
    """stream (short rows) and vector (long rows) SpMV kernels vs scipy"""
    import torch
    rng = np.random.RandomState(10)
    a = rand_csr(rng, shape[0], shape[1], density)
    A = dev.DevCsr.from_scipy(a)
    x = rng.standard_normal(shape[1])
    xd = torch.from_numpy(x).to(dev.device())
    want = a @ x
    for variant in (1, 2, 3, 4, 5):
        y = torch.full((shape[0],), np.nan, dtype=torch.float64, device=dev.device())
        dev._call('npb_spmv_variant', shape[0], A.nnz, A.indptr, A.indices, A.data, xd, y, variant)
        np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-13,
                                   atol=1e-13 * (np.abs(want).max() if a.nnz else 1))


@pytest.mark.parametrize('nrows,ncols,per_row', [(1, 1, 1), (255, 255, 3), (70001, 70001, 9),
                                                 (300000, 280000, 5), (20011, 9000, 40),
                                                 (5003, 5003, 130), (40, 100000, 9000)])
def test_spmv_tma_pipeline(dev, nrows, ncols, per_row):
    """persistent TMA-pipelined SpMV (variant 5) with every epilogue vs numpy: ragged rows,
    empty rows, units shorter and longer than a pipeline stage, array tails that are not a
    multiple of 16 bytes"""
    import torch
    rng = np.random.RandomState(nrows % 1000 + per_row)
    lens = rng.poisson(per_row, nrows).clip(0, ncols)
    lens[rng.random(nrows) < 0.05] = 0
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nnz = int(ptr[-1])
    idx = np.empty(nnz, np.int32)
    for r in range(nrows) if nrows < 6000 else ():
        idx[ptr[r]:ptr[r + 1]] = np.sort(rng.choice(ncols, lens[r], replace=False))
    if nrows >= 6000:        # banded columns around the diagonal, sorted, distinct
        base = (np.arange(nrows) * (ncols / nrows)).astype(np.int64)
        rows = np.repeat(np.arange(nrows), lens)
        off = np.arange(nnz) - ptr[rows]
        idx[:] = np.clip(base[rows] - lens[rows] // 2, 0, ncols - lens[rows]) + off
    a = sp.csr_matrix((rng.standard_normal(nnz), idx, ptr), shape=(nrows, ncols))
    x = rng.standard_normal(max(ncols, nrows))[:ncols]
    b = rng.standard_normal(nrows)
    dinv = rng.standard_normal(nrows)
    ax = a @ x
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev.device())
    ptr_d, idx_d, dat_d = t(ptr), t(idx), t(a.data)
    xd, bd, dd = t(x), t(b), t(dinv)
    scale = np.abs(a).dot(np.abs(x)).max() + np.abs(b).max() if nnz else 1.0
    for mode in (0, 1, 2, 3):
        if mode == 3 and nrows != ncols:
            continue
        want = [ax, b - ax, b + ax, (x[:nrows] + 0.7 * dinv * (b - ax)) if nrows == ncols else None][mode]
        y = torch.full((nrows,), np.nan, dtype=torch.float64, device=dev.device())
        dev._call('npb_spmv_epilogue', nrows, nnz, ptr_d, idx_d, dat_d, xd, mode, bd, dd, 0.7, y, 5)
        got = y.cpu().numpy()
        tol = 1e-13 * scale * (1 + np.abs(dinv).max() if mode == 3 else 1)
        np.testing.assert_allclose(got, want, rtol=0, atol=tol)
        # the same operator stored with fp32 values (preconditioner SpMV): exact for the
        # rounded matrix
        a32 = sp.csr_matrix((a.data.astype(np.float32).astype(np.float64), idx, ptr), shape=(nrows, ncols))
        ax32 = a32 @ x
        want32 = [ax32, b - ax32, b + ax32, (x[:nrows] + 0.7 * dinv * (b - ax32)) if nrows == ncols else None][mode]
        y = torch.full((nrows,), np.nan, dtype=torch.float64, device=dev.device())
        dev._call('npb_spmv_f32', nrows, nnz, ptr_d, idx_d, t(a.data.astype(np.float32)), xd, mode, bd, dd, 0.7, y)
        np.testing.assert_allclose(y.cpu().numpy(), want32, rtol=0, atol=tol)


@pytest.mark.parametrize('n,k', [(1, 1), (7, 3), (1000, 1), (4097, 5), (200001, 17), (300000, 40),
                                 (200000, 1), (200000, 2), (300000, 5), (65536, 6), (4096, 3),
                                 (200001, 1), (300001, 2), (65537, 6)])
def test_multi_dot_and_axpy(dev, n, k):
    """Gram-Schmidt building blocks of the FGMRES (csrc/krylov.cu): out = V^T w, and
    w += sign * V h with ||w||^2, against numpy (odd lengths, k above and below the
    16-vector pass, padded leading dimension)"""
    import torch
    from numpad_b200 import _lib
    rng = np.random.RandomState(n % 97 + k)
    ld = (n + 31) // 32 * 32
    V = np.zeros((k, ld))
    V[:, :n] = rng.standard_normal((k, n))
    w = rng.standard_normal(n)
    h = rng.standard_normal(k)
    Vd = torch.from_numpy(V).to(dev.device())
    wd = torch.from_numpy(w.copy()).to(dev.device())
    hd = torch.from_numpy(h).to(dev.device())
    scratch = torch.zeros(_lib.lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8,
                          device=dev.device())
    out = torch.zeros(k, dtype=torch.float64, device=dev.device())
    dev._call('npb_dot_multi', n, k, Vd, ld, wd, out, scratch)
    want = V[:, :n] @ w
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=0,
                               atol=1e-13 * np.abs(V[:, :n]).dot(np.abs(w)).max())
    nrm = torch.zeros(1, dtype=torch.float64, device=dev.device())
    dev._call('npb_axpy_multi', n, k, Vd, ld, hd, -1.0, wd, nrm, scratch)
    w2 = w - h @ V[:, :n]
    np.testing.assert_allclose(wd.cpu().numpy(), w2, rtol=0,
                               atol=1e-13 * (np.abs(w) + np.abs(h) @ np.abs(V[:, :n])).max())
    np.testing.assert_allclose(float(nrm.item()), w2 @ w2, rtol=1e-13)
    # a view that starts on an odd element takes the scalar path
    if n > 8:
        w3 = torch.from_numpy(w.copy()).to(dev.device())
        Vo = torch.from_numpy(np.ascontiguousarray(V[:, :])).to(dev.device())
        dev._call('npb_axpy_multi', n - 1, k, Vo.reshape(-1)[1:], ld, hd, 1.0, w3[1:], nrm, scratch)
        w4 = w[1:] + h @ V[:, 1:n]
        np.testing.assert_allclose(w3[1:].cpu().numpy(), w4, rtol=0,
                                   atol=1e-13 * (np.abs(w) .max() + (np.abs(h) @ np.abs(V[:, :n])).max()))


@pytest.mark.parametrize('n,k', [(1, 0), (9, 1), (4097, 3), (65536, 8), (65537, 9), (200001, 17),
                                 (300000, 40), (4096, 0), (300001, 33)])
def test_dcgs2_passes(dev, n, k):
    """the two sweeps of a Gram-Schmidt step with delayed re-orthogonalisation (csrc/krylov.cu,
    used by npb_fgmres): V^T [w0 w1] in one pass, and u <- (u - V c) / rho, w <- w - V a - b u_old
    with ||w||^2 in one pass; against numpy, odd lengths, k around the 8 / 16-vector passes, and
    an unaligned view (scalar path)"""
    import torch
    from numpad_b200 import _lib
    rng = np.random.RandomState(n % 89 + k)
    ld = (n + 31) // 32 * 32
    V = np.zeros((max(k, 1), ld))
    V[:k, :n] = rng.standard_normal((k, n))
    w0, w1 = rng.standard_normal(n), rng.standard_normal(n)
    coef = rng.standard_normal(2 * k + 2)
    D = dev.device()
    Vd = torch.from_numpy(V).to(D)
    scratch = torch.zeros(_lib.lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8, device=D)
    for off in ((0, 1) if n > 8 else (0,)):
        m = n - off
        Vv = Vd.reshape(-1)[off:]
        ud = torch.from_numpy(w0.copy()).to(D)[off:]
        wd = torch.from_numpy(w1.copy()).to(D)[off:]
        Vh = V[:k, off:n]
        if k > 0:
            o0 = torch.zeros(k, dtype=torch.float64, device=D)
            o1 = torch.zeros(k, dtype=torch.float64, device=D)
            dev._call('npb_dot_multi2', m, k, Vv, ld, ud, wd, o0, o1, scratch)
            for got, rhs in ((o0, w0[off:]), (o1, w1[off:])):
                np.testing.assert_allclose(got.cpu().numpy(), Vh @ rhs, rtol=0,
                                           atol=1e-13 * np.abs(Vh).dot(np.abs(rhs)).max())
        cd = torch.from_numpy(coef).to(D)
        nrm = torch.zeros(1, dtype=torch.float64, device=D)
        dev._call('npb_axpy_multi2', m, k, Vv, ld, cd, ud, wd, nrm, scratch)
        c, a, inv_rho, b = coef[:k], coef[k:2 * k], coef[2 * k], coef[2 * k + 1]
        u_new = (w0[off:] - c @ Vh) * inv_rho
        w_new = w1[off:] - a @ Vh - b * w0[off:]
        scale = (np.abs(w0).max() + np.abs(w1).max() + (np.abs(coef[:2 * k]).sum() + 1) * 5.0)
        np.testing.assert_allclose(ud.cpu().numpy(), u_new, rtol=0, atol=1e-13 * scale * abs(inv_rho))
        np.testing.assert_allclose(wd.cpu().numpy(), w_new, rtol=0, atol=1e-13 * scale)
        np.testing.assert_allclose(float(nrm.item()), w_new @ w_new, rtol=1e-12)


def test_spmv_hybrid_dia_ell_storage(dev):
    """hybrid image (diagonals + ELL remainder) of an operator most of whose entries sit on a few
    diagonals: amg.dia_image finds the strong offsets and the ELL width, the conversion places
    every entry, the SpMV matches scipy for every epilogue; rows with fewer remainder entries than
    the width and a remainder-free operator (-> pure DIA image) included"""
    import torch
    import ctypes
    from numpad_b200 import amg
    N = 220
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    lap = (sp.kron(sp.identity(N), T) + sp.kron(T, sp.identity(N))).tocsr()
    n = lap.shape[0]
    rng = np.random.RandomState(5)
    # remainder: 0..3 entries per row at drifting offsets (like the cross-coupling of the cavity's
    # momentum block)
    rows = np.repeat(np.arange(n), 3)
    drift = (np.arange(n) // N)[rows] + np.tile(np.array([7, 3 * N + 1, -5 * N]), n)
    cols = rows + drift
    ok = (cols >= 0) & (cols < n) & (rng.random(rows.size) < 0.8)
    rem = sp.csr_matrix((rng.standard_normal(ok.sum()), (rows[ok], cols[ok])), shape=(n, n))
    a = sp.csr_matrix(lap + rem)
    a.data = a.data.astype(np.float32).astype(np.float64)
    a.sort_indices()
    old, old_h = amg.DIA_MIN_ROWS, amg.HYB_IMAGES
    amg.DIA_MIN_ROWS, amg.HYB_IMAGES = 1000, True
    try:
        A = dev.DevCsr.from_scipy(a)
        img = amg.dia_image(A)
        assert img is not None and img[3] is not None
        offs, vals, ld, (w, eidx, eval_) = img
        assert sorted(offs) == [-N, -1, 0, 1, N] and 1 <= w <= 3
        off_h = (ctypes.c_int32 * 8)(*(offs + [0] * (8 - len(offs))))
        x, b, dinv = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
        t = lambda v: torch.from_numpy(v).to(dev.device())
        xd, bd, dd = t(x), t(b), t(dinv)
        ax = a @ x
        for mode, want in enumerate([ax, b - ax, b + ax, x + 0.7 * dinv * (b - ax)]):
            y = torch.full((n,), np.nan, dtype=torch.float64, device=dev.device())
            dev._call('npb_spmv_hyb', n, n, len(offs), ctypes.cast(off_h, ctypes.c_void_p).value, vals, ld,
                      w, eidx, eval_, xd, mode, bd, dd, 0.7, y)
            np.testing.assert_allclose(y.cpu().numpy(), want, rtol=0, atol=1e-12 * (np.abs(want).max() + 1))
        keep = []
        m = amg.csr_mat(A, keep)
        assert m.ndiag == 5 and m.ell_w == w and m.ell_idx and not m.data32
        # too narrow an ELL part is reported
        bad = torch.empty(1, dtype=torch.int64, device=dev.device())
        if w > 1:
            dev._call('npb_csr_to_hyb', n, n, A.indptr, A.indices, A.data, len(offs),
                      ctypes.cast(off_h, ctypes.c_void_p).value, vals, ld, w - 1, eidx, eval_, bad)
            assert int(bad.item()) > 0
        # an operator without remainder keeps its pure diagonal image
        L = dev.DevCsr.from_scipy(sp.csr_matrix(lap))
        img2 = amg.dia_image(L)
        assert img2 is not None and img2[3] is None
    finally:
        amg.DIA_MIN_ROWS, amg.HYB_IMAGES = old, old_h


@pytest.mark.parametrize('kind', ['mis1', 'mis2'])
def test_aggregation_properties(dev, kind):
    """MIS aggregation on a 2-D 5-point operator: every node gets an aggregate id in
    range, ids are dense, aggregates are connected through strong edges to their root
    within 1 (mis1) / 2 (mis2) steps, and MIS-2 aggregates are ~7 nodes, MIS-1 ~2.75"""
    from numpad_b200 import amg
    n1 = 96
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n1, n1))
    a = sp.csr_matrix(sp.kron(sp.identity(n1), T) + sp.kron(T, sp.identity(n1)))
    a.sort_indices()
    A = dev.DevCsr.from_scipy(a)
    agg, n_agg = amg.aggregate(A, 0.0, kind=kind)
    g = agg.cpu().numpy()
    n = a.shape[0]
    assert g.min() == 0 and g.max() == n_agg - 1
    sizes = np.bincount(g, minlength=n_agg)
    assert sizes.min() >= 1
    mean = n / n_agg
    assert (2.2 < mean < 3.4) if kind == 'mis1' else (5.5 < mean < 9.5), mean
    # connectivity: aggregate-local graph is connected (BFS inside each aggregate)
    from scipy.sparse.csgraph import connected_components
    coo = a.tocoo()
    same = g[coo.row] == g[coo.col]
    sub = sp.csr_matrix((np.ones(same.sum()), (coo.row[same], coo.col[same])), shape=a.shape)
    ncomp, _ = connected_components(sub, directed=False)
    assert ncomp == n_agg


@pytest.mark.parametrize('m,k,n,da,db', [(1, 1, 1, 1.0, 1.0), (300, 200, 250, 0.05, 0.05),
                                         (20000, 15000, 18000, 0.0004, 0.0006),
                                         (5000, 5000, 5000, 0.004, 0.01), (64, 3000, 4000, 0.3, 0.05)])
def test_row_batched_spgemm(dev, m, k, n, da, db):
    """csrc/spgemm.cu against scipy and against the sort-based path (bit-identical: same
    products, same summation order): empty rows, exact cancellations (pruned), duplicate
    columns inside a row, and a case whose rows exceed the batch capacity (falls back)"""
    import torch
    rng = np.random.RandomState(m + n)
    a = rand_csr(rng, m, k, da, zero_frac=0.02)
    b = rand_csr(rng, k, n, db)
    # exact cancellations: duplicate a column of b with the opposite sign in a second row
    if k > 10 and b.nnz > 0:
        b = b.tolil(); b[1] = -b[0]; b = sp.csr_matrix(b); b.sort_indices()
        a = a.tolil(); a[0] = 0; a[0, 0] = 2.0; a[0, 1] = 2.0; a = sp.csr_matrix(a); a.sort_indices()
    A, B = dev.DevCsr.from_scipy(a), dev.DevCsr.from_scipy(b)
    fast = dev._csr_times_csr_rows(A, B)
    # the warp-per-row kernel (rows of <= 512 partial products) and the CTA-batched kernel give
    # the same bits
    from numpad_b200 import _lib as _l
    was = _l.lib().cdll.npb_spgemm_set_warp_rows(0)
    try:
        batched = dev._csr_times_csr_rows(A, B)
    finally:
        _l.lib().cdll.npb_spgemm_set_warp_rows(was)
    assert (fast is None) == (batched is None)
    if fast is not None:
        assert fast.nnz == batched.nnz
        assert torch.equal(fast.indptr, batched.indptr) and torch.equal(fast.indices, batched.indices)
        assert torch.equal(fast.data, batched.data)
    old = dev.USE_SPGEMM_ROWS
    dev.USE_SPGEMM_ROWS = False
    try:
        slow = dev._csr_times_csr(A, B)
    finally:
        dev.USE_SPGEMM_ROWS = old
    want = sp.csr_matrix(a @ b)
    if fast is None:                       # a row with more partial products than half a batch
        ub = (a != 0).astype(np.int64) @ np.maximum(np.diff(b.indptr), 1)
        from numpad_b200 import _lib
        assert ub.max() > _lib.lib().cdll.npb_spgemm_rows_cap() // 2
        check(slow, want)
        return
    assert m != 64
    check(fast, want)
    f, s = fast.toscipy(), slow.toscipy()
    assert np.array_equal(f.indptr, s.indptr) and np.array_equal(f.indices, s.indices)
    assert np.array_equal(f.data, s.data)


def test_spmv_dia_storage(dev):
    """diagonal-storage SpMV (stencil operators on <= 8 constant offsets, fp32 values) with every
    epilogue vs scipy, on a 5-point Laplacian with one unknown removed (the cavity's pressure
    operator); the conversion counts entries that are on none of the offsets"""
    import torch
    import ctypes
    N = 300
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    a = (sp.kron(sp.identity(N), T) + sp.kron(T, sp.identity(N))).tocsr()[1:, 1:]
    a = sp.csr_matrix(a)
    a.data = a.data * (1 + 0.25 * np.random.RandomState(0).random(a.nnz))   # values exact in fp32?
    a.data = a.data.astype(np.float32).astype(np.float64)
    a.sort_indices()
    n = a.shape[0]
    A = dev.DevCsr.from_scipy(a)
    offs = [-N, -1, 0, 1, N]
    off_h = (ctypes.c_int32 * 8)(*(offs + [0, 0, 0]))
    ld = (n + 31) // 32 * 32
    vals = torch.empty(5 * ld, dtype=torch.float32, device=dev.device())
    bad = torch.empty(1, dtype=torch.int64, device=dev.device())
    dev._call('npb_csr_to_dia', n, A.indptr, A.indices, A.data, 5, ctypes.cast(off_h, ctypes.c_void_p).value, vals, ld, bad)
    assert int(bad.item()) == 0
    rng = np.random.RandomState(1)
    x, b, dinv = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    t = lambda v: torch.from_numpy(v).to(dev.device())
    xd, bd, dd = t(x), t(b), t(dinv)
    ax = a @ x
    for mode, want in enumerate([ax, b - ax, b + ax, x + 0.7 * dinv * (b - ax)]):
        y = torch.full((n,), np.nan, dtype=torch.float64, device=dev.device())
        dev._call('npb_spmv_dia', n, n, 5, ctypes.cast(off_h, ctypes.c_void_p).value, vals, ld, xd, mode, bd, dd, 0.7, y)
        np.testing.assert_allclose(y.cpu().numpy(), want, rtol=0, atol=1e-12 * (np.abs(want).max() + 1))
    # fused kernels of the pressure CG: Jacobi sweep + b . y, and p / q update + p . q
    from numpad_b200 import _lib
    red = torch.zeros(_lib.lib().cdll.npb_reduce_scratch_bytes(), dtype=torch.uint8, device=dev.device())
    offp = ctypes.cast(off_h, ctypes.c_void_p).value
    y = torch.full((n,), np.nan, dtype=torch.float64, device=dev.device())
    dot = torch.zeros(1, dtype=torch.float64, device=dev.device())
    dev._call('npb_spmv_dia_jacobi_dot', n, n, 5, offp, vals, ld, xd, bd, dd, 0.7, y, dot, red)
    want = x + 0.7 * dinv * (b - ax)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=0, atol=1e-12 * (np.abs(want).max() + 1))
    np.testing.assert_allclose(float(dot.item()), b @ want, rtol=1e-12, atol=1e-9)
    p0, q0 = rng.standard_normal(n), rng.standard_normal(n)
    rz = t(np.array([3.0, 2.0]))
    for first in (1, 0):
        pd, qd = t(p0.copy()), t(q0.copy())
        dev._call('npb_dia_cg_dir', n, 5, offp, vals, ld, xd, rz[0:], rz[1:], first, pd, qd, dot, red)
        beta = 0.0 if first else 1.5
        pw, qw = x + beta * p0, ax + beta * q0
        np.testing.assert_allclose(pd.cpu().numpy(), pw, rtol=0, atol=1e-12 * (np.abs(pw).max() + 1))
        np.testing.assert_allclose(qd.cpu().numpy(), qw, rtol=0, atol=1e-12 * (np.abs(qw).max() + 1))
        np.testing.assert_allclose(float(dot.item()), pw @ qw, rtol=1e-11, atol=1e-8)
    # a missing offset is reported
    off_h2 = (ctypes.c_int32 * 8)(-N, -1, 0, 1, 0, 0, 0, 0)
    dev._call('npb_csr_to_dia', n, A.indptr, A.indices, A.data, 4, ctypes.cast(off_h2, ctypes.c_void_p).value, vals, ld, bad)
    assert int(bad.item()) == n - N + 1 - 1 or int(bad.item()) > 0
    # and the multigrid descriptor picks the image up for a large enough operator
    from numpad_b200 import amg
    old = amg.DIA_MIN_ROWS
    amg.DIA_MIN_ROWS = 1000
    try:
        keep = []
        m = amg.csr_mat(A, keep)
        assert m.ndiag == 5 and m.dia_vals and sorted(m.dia_off[:5]) == offs and not m.data32
    finally:
        amg.DIA_MIN_ROWS = old
