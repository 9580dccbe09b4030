"""CPU prototype (numpy) of the flexible GMRES with DELAYED re-orthogonalisation that
csrc/gmres.cu implements: classical Gram-Schmidt applied twice, the second pass over vector j
fused with the first pass over vector j + 1, so the basis is read twice per iteration instead
of four times.  Compared with the plain CGS2 Arnoldi on a badly scaled non-symmetric matrix and
a variable (non-linear) preconditioner: orthogonality of V, Arnoldi residual ||A Z - V H||,
iterations to 1e-12.

  python tools/proto_dcgs2.py
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def hess_ls(H, beta, j):
    """min || beta e1 - H[:j+2,:j+1] y ||: returns y and the residual norm"""
    Hj = H[:j + 2, :j + 1]
    g = np.zeros(j + 2); g[0] = beta
    y, *_ = np.linalg.lstsq(Hj, g, rcond=None)
    return y, np.linalg.norm(g - Hj @ y)


def fgmres_cgs2(A, b, M, m, tol, passes=2):
    n = b.size
    beta = np.linalg.norm(b)
    V = np.zeros((n, m + 1)); Z = np.zeros((n, m)); H = np.zeros((m + 1, m))
    V[:, 0] = b / beta
    for j in range(m):
        Z[:, j] = M(V[:, j], j)
        w = A @ Z[:, j]
        h = V[:, :j + 1].T @ w; w = w - V[:, :j + 1] @ h
        h2 = 0.0
        if passes == 2:
            h2 = V[:, :j + 1].T @ w; w = w - V[:, :j + 1] @ h2
        H[:j + 1, j] = h + h2
        H[j + 1, j] = np.linalg.norm(w)
        V[:, j + 1] = w / H[j + 1, j]
        y, res = hess_ls(H, beta, j)
        if res <= tol * beta:
            break
    return Z[:, :j + 1] @ y, j + 1, V[:, :j + 2], Z[:, :j + 1], H[:j + 2, :j + 1]


def fgmres_dcgs2(A, b, M, m, tol):
    n = b.size
    beta = np.linalg.norm(b)
    V = np.zeros((n, m + 1)); Z = np.zeros((n, m)); H = np.zeros((m + 1, m))
    V[:, 0] = b / beta          # slot j holds u_j until pass 2 of step j makes it v_j
    eta = 0.0
    for j in range(m):
        u = V[:, j].copy()
        Z[:, j] = M(u, j)
        w = A @ Z[:, j]
        # pass 1: ONE sweep over v_0 .. v_{j-1}, u, w
        s = V[:, :j].T @ u
        t = V[:, :j].T @ w
        gamma = u @ w
        rho = np.sqrt(max(u @ u - s @ s, 0.0))
        gp = (gamma - s @ t) / rho
        a = t - (gp / rho) * s
        bb = gp / rho
        if j > 0:                # the column of the previous step now refers to the final v_j
            H[:j, j - 1] += eta * s
            H[j, j - 1] = eta * rho
        # pass 2: ONE sweep: v_j = (u - V s) / rho ;  w' = w - V a - bb u
        V[:, j] = (u - V[:, :j] @ s) / rho
        w = w - V[:, :j] @ a - bb * u
        eta = np.linalg.norm(w)
        H[:j, j] = t; H[j, j] = gp; H[j + 1, j] = eta
        V[:, j + 1] = w / eta
        y, res = hess_ls(H, beta, j)
        if res <= tol * beta:
            break
    return Z[:, :j + 1] @ y, j + 1, V[:, :j + 2], Z[:, :j + 1], H[:j + 2, :j + 1]


def main():
    rng = np.random.RandomState(0)
    N = 80
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N))
    A = (sp.kron(sp.eye(N), T) + sp.kron(T, sp.eye(N))
         + 0.8 * sp.kron(sp.eye(N), sp.diags([-1.0, 1.0], [-1, 1], shape=(N, N)))).tocsr()
    n = A.shape[0]
    A = sp.diags(10.0 ** rng.uniform(-1, 1, n)) @ A          # bad row scaling
    b = rng.standard_normal(n)
    Dinv = 1.0 / A.diagonal()
    Lo = sp.tril(A).tocsr()

    def M(v, j):                 # variable preconditioner: Gauss-Seidel sweeps, their number depends on j
        z = spla.spsolve_triangular(Lo, v, lower=True)
        for _ in range(j % 2):
            z = z + spla.spsolve_triangular(Lo, v - A @ z, lower=True)
        return z

    for name, fn in (('CGS1 ', lambda *a: fgmres_cgs2(*a, passes=1)), ('CGS2 ', fgmres_cgs2), ('DCGS2', fgmres_dcgs2)):
        x, its, V, Z, H = fn(A, b, M, 300, 1e-12)
        print('%s: %3d iterations, true relres %.2e, ||I - V^T V|| (all but the last vector) %.2e, ||A Z - V H|| / ||A Z|| %.2e' % (
            name, its, np.linalg.norm(b - A @ x) / np.linalg.norm(b),
            np.linalg.norm(np.eye(V.shape[1] - 1) - V[:, :-1].T @ V[:, :-1]), np.linalg.norm(A @ Z - V @ H) / np.linalg.norm(A @ Z)))


if __name__ == '__main__':
    main()
