"""CPU (scipy) prototype of the LSC / smoothed-aggregation solver, for exploring
algorithmic variants (aggregation, smoothers, cycle counts) before they are written
as CUDA.  Test infrastructure only; mirrors numpad_b200/amg.py + csrc/amg.cu.

    python tools/proto_lsc.py --grid 256 --h 2048 --variants base,mis2
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

WORK = [0.0]          # bytes moved by SpMV-shaped kernels (12 B/nnz + 20 B/row)


def mv(A, x):
    WORK[0] += 12.0 * A.nnz + 20.0 * A.shape[0]
    return A @ x


def nbr_max(S, v):
    """max of v over the row pattern of S (empty rows -> -1)"""
    out = np.full(S.shape[0], -1.0)
    nz = np.diff(S.indptr) > 0
    if S.nnz:
        red = np.maximum.reduceat(v[S.indices], S.indptr[:-1][nz])
        out[nz] = red
    return out


def strength(A, theta):
    A = A.tocsr()
    Ao = A - sp.diags(A.diagonal())
    Ao.eliminate_zeros()
    Ao = Ao.tocsr()
    ab = abs(Ao)
    mx = nbr_max(ab, ab.data) if False else np.asarray(ab.max(axis=1).todense()).ravel()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(ab.indptr))
    keep = ab.data >= theta * mx[rows]
    S = sp.csr_matrix((ab.data[keep], ab.indices[keep], np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=A.shape[0]))])), shape=A.shape)
    return S


def aggregate(A, theta, kind, rng):
    n = A.shape[0]
    S = strength(A, theta)
    pri = rng.permutation(n).astype(float)
    state = np.zeros(n, np.int8)          # 0 undecided 1 root 2 excluded
    for _ in range(100):
        und = state == 0
        if not und.any():
            break
        mp = np.where(und, pri, -1.0)
        m = np.maximum(mp, nbr_max(S, mp))
        if kind == 'mis2':
            m = np.maximum(m, nbr_max(S, m))
        root = und & (m == pri)
        state[root] = 1
        r = (state == 1).astype(float)
        near = nbr_max(S, r) > 0
        if kind == 'mis2':
            near = near | (nbr_max(S, np.maximum(r, near.astype(float))) > 0)
        state[(state == 0) & near] = 2
    roots = state == 1
    rid = np.cumsum(roots) - 1
    agg = np.full(n, -1, np.int64)
    agg[roots] = rid[roots]
    # join: strongest connected root
    for pass_ in range(3):
        un = agg < 0
        if not un.any():
            break
        have = (agg >= 0)
        # for each row pick the strongest neighbour that has an aggregate
        w = S.data * have[S.indices]
        rows = np.repeat(np.arange(n), np.diff(S.indptr))
        best = np.full(n, -1.0)
        np.maximum.at(best, rows, w)
        pick = (w == best[rows]) & (w > 0) & un[rows]
        newagg = agg.copy()
        # first match wins
        idxs = np.nonzero(pick)[0][::-1]
        newagg[rows[idxs]] = agg[S.indices[idxs]]
        agg = newagg
    un = agg < 0
    if un.any():                          # isolated nodes: own aggregates
        k = agg.max() + 1
        agg[un] = k + np.arange(un.sum())
    return agg, int(agg.max() + 1)


def merge_small(A, agg, na, theta, max_small=2, max_size=6):
    """merge aggregates of <= max_small nodes into their most strongly connected neighbouring
    aggregate (greedy, sequential): aggregates of ~4-5 nodes instead of the 2.75 of MIS-1"""
    S = strength(A, theta).tocoo()
    G = sp.csr_matrix((S.data, (agg[S.row], agg[S.col])), shape=(na, na))
    G.sum_duplicates()
    G.setdiag(0)
    G.eliminate_zeros()
    G = G.tocsr()
    size = np.bincount(agg, minlength=na).astype(np.int64)
    target = np.arange(na)
    order = np.argsort(size, kind='stable')
    for a in order:
        if size[a] > max_small or target[a] != a:
            continue
        nb, w = G.indices[G.indptr[a]:G.indptr[a + 1]], G.data[G.indptr[a]:G.indptr[a + 1]]
        best, bw = -1, 0.0
        for b, wb in zip(nb, w):
            r = b
            while target[r] != r:
                r = target[r]
            if r != a and size[r] + size[a] <= max_size and wb > bw:
                best, bw = r, wb
        if best >= 0:
            target[a] = best
            size[best] += size[a]
            size[a] = 0
    root = target.copy()
    for _ in range(20):
        root = root[root]
    ids = np.unique(root, return_inverse=True)[1]
    return ids[agg], int(ids.max() + 1)


def merge_small_parallel(A, agg, na):
    """the product's merge rule (numpad_b200/amg.py: merge_small_aggregates, pure torch)"""
    import torch
    from numpad_b200 import amg
    A = A.tocsr()
    first = np.full(na, A.shape[0])
    np.minimum.at(first, agg, np.arange(A.shape[0]))
    order = np.argsort(first)
    ren = np.empty(na, np.int64)
    ren[order] = np.arange(na)
    agg = ren[agg]
    flag = np.zeros(A.shape[0], np.int64)
    flag[first[order]] = 1
    rid = np.concatenate([[0], np.cumsum(flag)]).astype(np.int32)
    t = lambda v, dt: torch.from_numpy(np.ascontiguousarray(v)).to(dt)
    a2, n2, _ = amg.merge_small_aggregates(t(A.indptr, torch.int32), t(A.indices, torch.int32),
                                           t(A.data, torch.float64), t(agg, torch.int32), na,
                                           t(rid, torch.int32))
    return a2.numpy().astype(np.int64), n2


def filtered(A, theta):
    A = A.tocsr()
    d = A.diagonal()
    Ao = (A - sp.diags(d)).tocsr()
    ab = abs(Ao)
    mx = np.asarray(ab.max(axis=1).todense()).ravel()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(Ao.indptr))
    keep = np.abs(Ao.data) >= theta * mx[rows]
    lump = np.bincount(rows[~keep], weights=Ao.data[~keep], minlength=A.shape[0])
    K = sp.csr_matrix((Ao.data[keep], (rows[keep], Ao.indices[keep])), shape=A.shape)
    return (K + sp.diags(d + lump)).tocsr()


def truncate(P, theta):
    P = P.tocsr()
    rows = np.repeat(np.arange(P.shape[0]), np.diff(P.indptr))
    ab = np.abs(P.data)
    mx = np.zeros(P.shape[0])
    np.maximum.at(mx, rows, ab)
    keep = (ab >= theta * mx[rows]) & (P.data != 0)
    s_all = np.bincount(rows, weights=P.data, minlength=P.shape[0])
    s_kept = np.bincount(rows[keep], weights=P.data[keep], minlength=P.shape[0])
    scale = np.where((np.abs(s_kept) >= 0.5 * np.abs(s_all)) & (s_kept != 0), s_all / np.where(s_kept != 0, s_kept, 1), 1.0)
    return sp.csr_matrix((P.data[keep] * scale[rows[keep]], (rows[keep], P.indices[keep])), shape=P.shape)


class Hier:
    def __init__(self, A, theta=0.0, nu=1, omega=0.67, omega_p=0.67, coarse_max=600, agg='mis1',
                 p_trunc=0.1, smoother='jacobi', seed=0, cheb_deg=2, filt=True, gal='full', omega_c=None, max_levels=20, coarse_sweeps=None,
                 gamma=1, gamma_from=1, psmooth=1, nu_post=None):
        rng = np.random.RandomState(seed)
        self.levels = []
        cur = A.tocsr()
        self.nu, self.omega, self.smoother, self.cheb_deg = nu, omega, smoother, cheb_deg
        self.coarse_sweeps = coarse_sweeps
        self.gamma, self.gamma_from = gamma, gamma_from
        self.nu_post = nu if nu_post is None else nu_post
        while True:
            n = cur.shape[0]
            d = cur.diagonal()
            dinv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1), 1.0)
            lev = dict(A=cur, dinv=dinv)
            if smoother == 'cheb':
                # spectral radius estimate of D^-1 A by a few power iterations
                v = rng.standard_normal(n)
                for _ in range(10):
                    v = dinv * (cur @ v)
                    lam = np.linalg.norm(v)
                    v /= lam
                lev['lam'] = 1.1 * lam
            self.levels.append(lev)
            if n <= coarse_max or len(self.levels) >= max_levels:
                break
            kind = 'mis1' if agg in ('mis1m', 'mis1p') else agg
            ag, na = aggregate(cur, theta, kind, rng)
            if na > 0.6 * n and theta > 0:
                ag, na = aggregate(cur, 0.0, kind, rng)
            if agg == 'mis1m':
                ag, na = merge_small(cur, ag, na, theta)
            if agg == 'mis1p':
                ag, na = merge_small_parallel(cur, ag, na)
            if na >= 0.9 * n:
                break
            T = sp.csr_matrix((np.ones(n), (np.arange(n), ag)), shape=(n, na))
            AF = filtered(cur, theta) if (theta > 0 and filt) else cur
            dF = AF.diagonal()
            dFinv = np.where(dF != 0, 1.0 / np.where(dF != 0, dF, 1), 1.0)
            P = (T - sp.diags(omega_p * dFinv) @ (AF @ T)).tocsr()
            for _ in range(1, psmooth):
                P = (P - sp.diags(omega_p * dFinv) @ (AF @ P)).tocsr()
            if p_trunc > 0:
                P = truncate(P, p_trunc)
            lev['P'] = P.tocsr()
            lev['R'] = P.T.tocsr()
            if gal == 'filtered' and theta > 0:
                cur = (lev['R'] @ AF @ P).tocsr()
            elif gal == 'postfilter' and theta > 0:
                cur = filtered((lev['R'] @ cur @ P).tocsr(), theta)
            else:
                cur = (lev['R'] @ cur @ P).tocsr()
        self.cinv = None
        if self.levels[-1]['A'].shape[0] <= 2500:
            try:
                self.cinv = np.linalg.inv(self.levels[-1]['A'].toarray())
            except np.linalg.LinAlgError:
                self.cinv = np.linalg.pinv(self.levels[-1]['A'].toarray())
        self.sizes = [l['A'].shape[0] for l in self.levels]
        self.cx = sum(l['A'].nnz for l in self.levels) / self.levels[0]['A'].nnz

    def smooth(self, l, b, x, nu):
        L = self.levels[l]
        if self.smoother == 'jacobi':
            for s in range(nu):
                if x is None:
                    x = self.omega * L['dinv'] * b
                else:
                    x = x + self.omega * L['dinv'] * (b - mv(L['A'], x))
            return x
        # Chebyshev on D^-1 A over [lam/30 .. lam], degree = nu * cheb_deg
        lam = L['lam']
        a, c = lam / 30.0, lam
        theta_, delta = 0.5 * (c + a), 0.5 * (c - a)
        deg = nu * self.cheb_deg
        r = b if x is None else b - mv(L['A'], x)
        x = np.zeros_like(b) if x is None else x
        sigma = theta_ / delta
        rho = 1.0 / sigma
        dvec = L['dinv'] * r / theta_
        for k in range(deg):
            x = x + dvec
            if k == deg - 1:
                break
            r = r - mv(L['A'], dvec)
            rho_new = 1.0 / (2 * sigma - rho)
            dvec = rho_new * rho * dvec + 2 * rho_new / delta * (L['dinv'] * r)
            rho = rho_new
        return x

    def vcycle(self, b, l=0):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            if self.cinv is not None:
                return self.cinv @ b
            return self.smooth(l, b, None, self.coarse_sweeps or (2 * self.nu + 6))
        x = self.smooth(l, b, None, self.nu)
        reps = self.gamma if (l + 1 >= self.gamma_from and l + 2 < len(self.levels)) else 1
        for _ in range(reps):
            r = b - mv(L['A'], x)
            bc = mv(L['R'], r)
            xc = self.vcycle(bc, l + 1)
            x = x + mv(L['P'], xc)
        return self.smooth(l, b, x, self.nu_post)


def pcg(h, A, b, k):
    x = np.zeros_like(b)
    r = b.copy()
    rz_old = 1.0
    p = None
    for i in range(k):
        z = h.vcycle(r)
        rz = r @ z
        p = z if i == 0 else z + (rz / rz_old) * p
        q = mv(A, p)
        alpha = rz / (p @ q)
        x = x + alpha * p
        r = r - alpha * q
        rz_old = rz
    return x


def fgmres(matvec, prec, b, rtol=1e-10, restart=100, maxiter=300):
    n = b.size
    x = np.zeros(n)
    bn = np.linalg.norm(b)
    its = 0
    while its < maxiter:
        r = b - matvec(x)
        beta = np.linalg.norm(r)
        if beta <= rtol * bn:
            break
        V = [r / beta]
        Z = []
        H = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = beta
        for j in range(restart):
            z = prec(V[j])
            Z.append(z)
            w = matvec(z)
            for i in range(j + 1):
                H[i, j] = V[i] @ w
                w = w - H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            V.append(w / H[j + 1, j])
            its += 1
            y, res, *_ = np.linalg.lstsq(H[:j + 2, :j + 1], g[:j + 2], rcond=None)
            rn = np.linalg.norm(H[:j + 2, :j + 1] @ y - g[:j + 2])
            if rn <= rtol * bn or its >= maxiter:
                break
        x = x + sum(yi * zi for yi, zi in zip(y, Z))
    return x, its, np.linalg.norm(b - matvec(x)) / bn


def build_problem(N, hden):
    from oracle import numpad_oracle as orc
    from numpad_b200.workloads.cavity import CavityProblem
    P = CavityProblem(orc, N, h=1.0 / hden)
    u0 = P.fixed_state()
    u = orc.array(u0.copy())
    res = P.residual(u, orc.array(u0), 1.0, 0)
    J = res.diff(u).tocsr()
    return J, np.ravel(res._value)


VARIANTS = {
    # name: (kwargs_L, kwargs_A, cycles_l, cycles_a)
    'base': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), 4, 1),
    'mis2': (dict(theta=0.0, nu=1, agg='mis2'), dict(theta=0.25, nu=2, agg='mis2'), 4, 1),
    'mis2L': (dict(theta=0.0, nu=1, agg='mis2'), dict(theta=0.25, nu=2), 4, 1),
    'mis2A': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, agg='mis2'), 4, 1),
    'mis2_nu2': (dict(theta=0.0, nu=2, agg='mis2'), dict(theta=0.25, nu=2, agg='mis2'), 4, 1),
    'mis2_nu2_c3': (dict(theta=0.0, nu=2, agg='mis2'), dict(theta=0.25, nu=2, agg='mis2'), 3, 1),
    'mis2_nu2_c2': (dict(theta=0.0, nu=2, agg='mis2'), dict(theta=0.25, nu=2, agg='mis2'), 2, 1),
    'mis2_cheb': (dict(theta=0.0, nu=1, agg='mis2', smoother='cheb'), dict(theta=0.25, nu=1, agg='mis2', smoother='cheb'), 4, 1),
    'mis2_cheb_c2': (dict(theta=0.0, nu=1, agg='mis2', smoother='cheb'), dict(theta=0.25, nu=1, agg='mis2', smoother='cheb'), 2, 1),
    'mis2_cheb3_c2': (dict(theta=0.0, nu=1, agg='mis2', smoother='cheb', cheb_deg=3), dict(theta=0.25, nu=1, agg='mis2', smoother='cheb', cheb_deg=3), 2, 1),
    'base_c2': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), 2, 1),
    'base_c3': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), 3, 1),
    'mis2L_c2': (dict(theta=0.0, nu=1, agg='mis2'), dict(theta=0.25, nu=2), 2, 1),
    'mis2L_c3': (dict(theta=0.0, nu=1, agg='mis2'), dict(theta=0.25, nu=2), 3, 1),
    'mis2L_nu2_c2': (dict(theta=0.0, nu=2, agg='mis2'), dict(theta=0.25, nu=2), 2, 1),
    'exactA': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), 4, -1),
    'exactL': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), -1, 1),
    'exactAL': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), -1, -1),
    'base_a2': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), 4, 2),
    'base_c2_a2': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2), 2, 2),
    'galF': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, gal='filtered'), 4, 1),
    'galP': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, gal='postfilter'), 4, 1),
    'galF_mis2L': (dict(theta=0.0, nu=1, agg='mis2'), dict(theta=0.25, nu=2, gal='filtered'), 4, 1),
    'galF_mis2': (dict(theta=0.0, nu=1, agg='mis2'), dict(theta=0.25, nu=2, gal='filtered', agg='mis2'), 4, 1),
    'galF_mis2_nu3': (dict(theta=0.0, nu=1, agg='mis2'), dict(theta=0.25, nu=3, gal='filtered', agg='mis2'), 4, 1),
    'A2lev': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, max_levels=2), 4, 1),
    'A3lev': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, max_levels=3), 4, 1),
    'A3lev_s4': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, max_levels=3, coarse_sweeps=4), 4, 1),
    'A4lev': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, max_levels=4), 4, 1),
    'A3lev_mis2': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, max_levels=3, agg='mis2'), 4, 1),
    'A2lev_mis2': (dict(theta=0.0, nu=1), dict(theta=0.25, nu=2, max_levels=2, agg='mis2'), 4, 1),
    'mis1mL': (dict(theta=0.0, nu=1, agg='mis1m'), dict(theta=0.25, nu=2), 4, 1),
    'mis1mL_c3': (dict(theta=0.0, nu=1, agg='mis1m'), dict(theta=0.25, nu=2), 3, 1),
    'mis1pL': (dict(theta=0.0, nu=1, agg='mis1p'), dict(theta=0.25, nu=2), 4, 1),
    'mis1pL_34': (dict(theta=0.0, nu=1, agg='mis1p'), dict(theta=0.25, nu=2), 4, 1),
    'mis2L_W': (dict(theta=0.0, nu=1, agg='mis2', gamma=2), dict(theta=0.25, nu=2), 4, 1),
    'mis2L_W2': (dict(theta=0.0, nu=1, agg='mis2', gamma=2, gamma_from=2), dict(theta=0.25, nu=2), 4, 1),
    'mis2L_ps2': (dict(theta=0.0, nu=1, agg='mis2', psmooth=2), dict(theta=0.25, nu=2), 4, 1),
    'mis2L_v12': (dict(theta=0.0, nu=1, agg='mis2', nu_post=2), dict(theta=0.25, nu=2), 4, 1),
    'mis1pL_W2': (dict(theta=0.0, nu=1, agg='mis1p', gamma=2, gamma_from=2), dict(theta=0.25, nu=2), 4, 1),
    'base_W2': (dict(theta=0.0, nu=1, gamma=2, gamma_from=2), dict(theta=0.25, nu=2), 4, 1),
    'mis2_t0': (dict(theta=0.0, nu=1, agg='mis2', p_trunc=0.0), dict(theta=0.25, nu=2, agg='mis2', p_trunc=0.0), 4, 1),
}


def run(J, b, name):
    kl, ka, cl, ca = VARIANTS[name]
    n = J.shape[0]
    d = J.diagonal()
    vs, ps = np.nonzero(d != 0)[0], np.nonzero(d == 0)[0]
    A, G, D = J[vs][:, vs].tocsr(), J[vs][:, ps].tocsr(), J[ps][:, vs].tocsr()
    L = (-(D @ G)).tocsr()
    t0 = time.time()
    hl = Hier(L, **kl)
    ha = Hier(A, **ka)

    luL = spla.splu(L.tocsc()) if cl < 0 else None
    luA = spla.splu(A.tocsc()) if ca < 0 else None

    def prec(r):
        rv, rp = r[vs], r[ps]
        p1 = pcg(hl, L, rp, cl) if cl > 0 else luL.solve(rp)
        q = mv(D, mv(A, mv(G, p1)))
        p = -(pcg(hl, L, q, cl) if cl > 0 else luL.solve(q))
        rv = rv - mv(G, p)
        if ca < 0:
            u = luA.solve(rv)
        else:
            u = ha.vcycle(rv)
        for _ in range(1, ca):
            u = u + ha.vcycle(rv - mv(A, u))
        z = np.empty(n)
        z[vs], z[ps] = u, p
        return z

    WORK[0] = 0.0
    x, its, rr = fgmres(lambda v: mv(J, v), prec, b)
    w = WORK[0]
    jb = 12.0 * J.nnz + 20.0 * n
    print('%-14s its %3d relres %.1e  work/it = %6.1f J-SpMV  total = %7.0f J-SpMV | L %s cx %.2f | A %s cx %.2f | %.0fs'
          % (name, its, rr, w / jb / max(its, 1), w / jb, hl.sizes, hl.cx, ha.sizes, ha.cx, time.time() - t0), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--grid', type=int, default=128)
    ap.add_argument('--h', type=float, default=0.0)
    ap.add_argument('--variants', default='base,mis2')
    a = ap.parse_args()
    J, b = build_problem(a.grid, a.h or a.grid)
    print('grid %d h=1/%g n=%d nnz=%d' % (a.grid, a.h or a.grid, J.shape[0], J.nnz), flush=True)
    for v in a.variants.split(','):
        run(J, b, v)
