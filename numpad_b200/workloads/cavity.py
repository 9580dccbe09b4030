"""Lid-driven cavity, staggered (MAC) grid, implicit Euler in time.

Same discretisation and the same operation sequence as the reference's
``test/cavity.py:31-94`` (config 1/2 of BASELINE.json), restated with the grid
parameters carried by an object instead of module globals so that it can be
instantiated at any N and against any array namespace.

Unknown layout (field-major, reference test/cavity.py:31-37):
    [ p (N*N) | u_x interior ((N-1)*N) | u_y interior (N*(N-1)) ],  n = 3N^2 - 2N

Provenance: a workload definition, not engine code -- the operation sequence of the
reference's test/cavity.py:31-94 (qiqi/numpad, GPLv3) is reproduced on purpose, because parity
is defined on the tape it records (same states, same local Jacobians, same Newton trajectory).
"""
import numpy as np


class CavityProblem:
    def __init__(self, xp, N, Re=10000.0, h=None):
        self.xp, self.N, self.Re = xp, int(N), Re
        self.dx = self.dy = (1. / N) if h is None else h

    @property
    def n(self):
        return self.N * (3 * self.N - 2)

    # -- config-2(ii) fixed state of SURVEY.md section 8(d) ------------------ #
    def fixed_state(self, seed=0, amp=1e-2):
        return amp * np.random.RandomState(seed).standard_normal(self.n)

    def partition(self, world):
        """owner[i] (numpy int32) for a strip decomposition along the first grid
        index: every field of a cell column goes to the strip of that column
        (SURVEY section 8e), so rows of J reference only their own strip plus
        a one-cell halo."""
        N = self.N
        strip = lambda i: np.minimum((i * world) // N, world - 1)
        ip = np.repeat(np.arange(N), N)              # p[i, j]
        iux = np.repeat(np.arange(N - 1), N)         # u_x interior face i+1 | cell i
        iuy = np.repeat(np.arange(N), N - 1)         # u_y[i, j]
        return np.concatenate([strip(ip), strip(iux), strip(iuy)]).astype(np.int32)

    def fields(self, q):
        """Split the unknown vector and pad the velocities with ghost faces
        that carry the wall / lid boundary conditions."""
        xp, N = self.xp, self.N
        n_p, n_ux = N * N, N * (N - 1)
        p = q[:n_p].reshape([N, N])
        ux = xp.zeros([N + 1, N + 2])
        uy = xp.zeros([N + 2, N + 1])
        ux[1:-1, 1:-1] = q[n_p:n_p + n_ux].reshape([N - 1, N])
        uy[1:-1, 1:-1] = q[n_p + n_ux:].reshape([N, N - 1])
        ux[:, 0] = 1 - ux[:, 1]        # moving lid: tangential velocity 1
        ux[:, -1] = -ux[:, -2]
        uy[0, :] = -uy[1, :]
        uy[-1, :] = -uy[-2, :]
        return p, ux, uy

    def residual(self, q, q_old, dt, f=0):
        xp, N, dx, dy, Re = self.xp, self.N, self.dx, self.dy, self.Re
        p, ux, uy = self.fields(q)

        # momentum fluxes at cell centres and at grid nodes
        ux_c = 0.5 * (ux[1:, 1:-1] + ux[:-1, 1:-1])
        uy_c = 0.5 * (uy[1:-1, 1:] + uy[1:-1, :-1])
        uxx_c = ux_c ** 2
        uyy_c = uy_c ** 2
        ux_n = 0.5 * (ux[:, 1:] + ux[:, :-1])
        uy_n = 0.5 * (uy[1:, :] + uy[:-1, :])
        uxy_n = ux_n * uy_n

        d_uxx_dx = (uxx_c[1:, :] - uxx_c[:-1, :]) / dx
        d_uyy_dy = (uyy_c[:, 1:] - uyy_c[:, :-1]) / dy
        d_uxy_dx = (uxy_n[1:, 1:-1] - uxy_n[:-1, 1:-1]) / dx
        d_uxy_dy = (uxy_n[1:-1, 1:] - uxy_n[1:-1, :-1]) / dy
        dp_dx = (p[1:, :] - p[:-1, :]) / dx
        dp_dy = (p[:, 1:] - p[:, :-1]) / dy

        adv_x = d_uxx_dx + d_uxy_dy + dp_dx
        adv_y = d_uyy_dy + d_uxy_dx + dp_dy

        lap_x = (ux[2:, 1:-1] - 2 * ux[1:-1, 1:-1] + ux[:-2, 1:-1]) / dx ** 2 \
              + (ux[1:-1, 2:] - 2 * ux[1:-1, 1:-1] + ux[1:-1, :-2]) / dy ** 2
        lap_y = (uy[2:, 1:-1] - 2 * uy[1:-1, 1:-1] + uy[:-2, 1:-1]) / dx ** 2 \
              + (uy[1:-1, 2:] - 2 * uy[1:-1, 1:-1] + uy[1:-1, :-2]) / dy ** 2

        n_p, n_ux = N * N, N * (N - 1)
        ux_old = q_old[n_p:n_p + n_ux].reshape([N - 1, N])
        uy_old = q_old[n_p + n_ux:].reshape([N, N - 1])
        ddt_x = (ux[1:-1, 1:-1] - ux_old) / dt
        ddt_y = (uy[1:-1, 1:-1] - uy_old) / dt

        # continuity; the pressure level is pinned in the corner cell
        div = (ux[1:, 1:-1] - ux[:-1, 1:-1]) / dx \
            + (uy[1:-1, 1:] - uy[1:-1, :-1]) / dy
        div[0, 0] = p[0, 0]

        parts = [div, ddt_x + adv_x - lap_x / Re, ddt_y + adv_y - lap_y / Re]
        return xp.hstack([xp.ravel(r) for r in parts]) - f
