"""The numpy model of the FGMRES loop of csrc/gmres.cu (tools/proto_dcgs2.py): classical
Gram-Schmidt with DELAYED re-orthogonalisation keeps the basis as orthogonal as two full passes
per step (CGS2), satisfies the flexible Arnoldi relation A Z = V H to rounding, and converges in
the same number of iterations -- while a single pass loses orthogonality.  CPU only."""
import importlib.util
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('proto_dcgs2', os.path.join(ROOT, 'tools', 'proto_dcgs2.py'))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


def _problem(N=48):
    rng = np.random.RandomState(1)
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    A = (sp.kron(sp.eye(N), T) + sp.kron(T, sp.eye(N))
         + 0.8 * sp.kron(sp.eye(N), sp.diags([-1.0, 1.0], [-1, 1], shape=(N, N)))).tocsr()
    A = sp.diags(10.0 ** rng.uniform(-1, 1, A.shape[0])) @ A
    b = rng.standard_normal(A.shape[0])
    Lo = sp.tril(A).tocsr()

    def M(v, j):        # a preconditioner that changes from step to step (flexible GMRES)
        z = spla.spsolve_triangular(Lo, v, lower=True)
        for _ in range(j % 2):
            z = z + spla.spsolve_triangular(Lo, v - A @ z, lower=True)
        return z
    return A, b, M


def test_delayed_reorthogonalisation_matches_cgs2():
    A, b, M = _problem()
    x2, its2, V2, Z2, H2 = proto.fgmres_cgs2(A, b, M, 200, 1e-12)
    xd, itsd, Vd, Zd, Hd = proto.fgmres_dcgs2(A, b, M, 200, 1e-12)
    assert itsd == its2 and its2 > 20
    for x in (x2, xd):
        assert np.linalg.norm(b - A @ x) <= 1.5e-12 * np.linalg.norm(b)
    # all basis vectors but the last one (which has had its first pass only) are orthonormal to
    # rounding, as with two passes per step
    G = Vd[:, :-1].T @ Vd[:, :-1]
    assert np.linalg.norm(np.eye(G.shape[0]) - G) < 1e-13
    # the flexible Arnoldi relation holds with the corrected Hessenberg columns
    assert np.linalg.norm(A @ Zd - Vd @ Hd) <= 1e-14 * np.linalg.norm(A @ Zd)
    # and the solutions agree
    assert np.linalg.norm(xd - x2) <= 1e-9 * np.linalg.norm(x2)


def test_single_pass_loses_orthogonality():
    A, b, M = _problem()
    _, its, V1, _, _ = proto.fgmres_cgs2(A, b, M, 200, 1e-12, passes=1)
    G = V1[:, :-1].T @ V1[:, :-1]
    assert np.linalg.norm(np.eye(G.shape[0]) - G) > 1e-6      # the reason a second pass is needed
