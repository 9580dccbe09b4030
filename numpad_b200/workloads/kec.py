"""2-D compressible Euler / Navier-Stokes residuals, kinetic-energy-conserving
finite-volume scheme on a curvilinear duct mesh, implicit Euler in time.

Same discretisation and the same operation sequence as the reference's
``test/euler_kec.py:7-139`` (config 4 of BASELINE.json) and
``test/navierstokes.py:7-177`` (config 5), restated with the mesh size and the
boundary data carried by an object instead of module globals.

Unknown layout (variable-major, reference euler_kec.py:11-12):
    w = [rho | rho*u | rho*v | E], each Ni x Nj,   n = 4 * Ni * Nj
Synthetic mesh of SURVEY.md section 8(d) config 4: a straight 10 x 1 duct with
a smooth distortion, so that any Ni x Nj can be generated.

Provenance: a workload definition, not engine code -- the operation sequence of the
reference's test/euler_kec.py:7-139 and test/navierstokes.py:7-177 (qiqi/numpad, GPLv3) is reproduced on purpose, because parity
is defined on the tape it records (same states, same local Jacobians, same Newton trajectory).
"""
import numpy as np


def synthetic_mesh(Ni, Nj):
    x = np.linspace(0, 10, Ni + 1)
    y = np.linspace(0, 1, Nj + 1)
    y, x = np.meshgrid(y, x)
    x = x + 0.02 * np.sin(2 * np.pi * y)
    y = y * (1 + 0.1 * np.sin(0.2 * np.pi * x))
    return x, y


class DuctGeometry:
    """cell areas, face vectors and wall normals (reference geo2d,
    euler_kec.py:121-139)"""

    def __init__(self, xp, xy):
        xy = xp.array(xy)
        self.xy = xy
        self.xyc = (xy[:, 1:, 1:] + xy[:, :-1, 1:] +
                    xy[:, 1:, :-1] + xy[:, :-1, :-1]) / 4
        self.dxy_i = xy[:, :, 1:] - xy[:, :, :-1]
        self.dxy_j = xy[:, 1:, :] - xy[:, :-1, :]
        self.L_j = xp.sqrt(self.dxy_j[0] ** 2 + self.dxy_j[1] ** 2)
        self.normal_j = xp.array([self.dxy_j[1] / self.L_j,
                                  -self.dxy_j[0] / self.L_j])
        self.area = self._tri(self.dxy_i[:, :-1, :], self.dxy_j[:, :, 1:]) \
            + self._tri(self.dxy_i[:, 1:, :], self.dxy_j[:, :, :-1])

    @staticmethod
    def _tri(a, b):
        return 0.5 * (a[1] * b[0] - a[0] * b[1])


class KecProblem:
    def __init__(self, xp, Ni, Nj, viscous=False, mu=1.0, pt_in=None, p_out=1E5, mesh=None):
        self.xp, self.Ni, self.Nj = xp, int(Ni), int(Nj)
        self.viscous, self.mu = viscous, mu
        self.pt_in = (1.2E5 if viscous else 1.05E5) if pt_in is None else pt_in
        self.p_out = p_out
        x, y = synthetic_mesh(Ni, Nj) if mesh is None else mesh
        self.geo = DuctGeometry(xp, np.array([x, y]))

    @property
    def n(self):
        return 4 * self.Ni * self.Nj

    def fixed_state(self, seed=0, amp=1e-3):
        """uniform subsonic flow with a seeded 0.1 % perturbation"""
        Ni, Nj = self.Ni, self.Nj
        w = np.empty([4, Ni, Nj])
        w[0] = 1.0
        w[1] = 0.3 * np.sqrt(1.4E5)
        w[2] = 0.0
        w[3] = 1E5 / 0.4 + 0.5 * w[1] ** 2
        w *= 1 + amp * np.random.RandomState(seed).standard_normal(w.shape)
        return w.ravel()

    # ---- reference euler_kec.py:55-64 ---- #
    @staticmethod
    def primitive(w):
        rho = w[0]
        u = w[1] / rho
        v = w[2] / rho
        E = w[3]
        p = 0.4 * (E - 0.5 * (u * w[1] + v * w[2]))
        return rho, u, v, E, p

    # ---- reference euler_kec.py:7-53: ghost cells from boundary conditions ---- #
    def extend(self, w_interior):
        xp, Ni, Nj, geo = self.xp, self.Ni, self.Nj, self.geo
        w = xp.zeros([4, Ni + 2, Nj + 2])
        w[:, 1:-1, 1:-1] = w_interior.reshape([4, Ni, Nj])

        # subsonic inlet: total pressure and density prescribed
        rho, u, v, E, p = self.primitive(w[:, 1, 1:-1])
        c2 = 1.4 * p / rho
        c = xp.sqrt(c2)
        # (the two reference scripts differ in three places of this routine:
        # navierstokes.py:18,38,42-48 vs euler_kec.py:18,38,42-51)
        mach2 = u ** 2 / c2 if self.viscous else (u ** 2 + v ** 2) / c2
        rhot = rho * (1 + 0.2 * mach2) ** 2.5      # noqa: F841 (recorded like the reference)
        pt = p * (1 + 0.2 * mach2) ** 3.5
        d_rho = 1 - rho
        d_pt = self.pt_in - pt
        d_u = d_pt / (rho * (u + c))
        d_p = rho * c * d_u
        rho = rho + d_rho
        u = u + d_u
        p = p + d_p
        w[0, 0, 1:-1] = rho
        w[1, 0, 1:-1] = rho * u
        w[2, 0, 1:-1] = 0
        w[3, 0, 1:-1] = p / 0.4 + 0.5 * rho * u ** 2

        # outlet: static pressure prescribed
        w[:, -1, 1:-1] = w[:, -2, 1:-1]
        rho, u, v, E, p = self.primitive(w[:, -1, 1:-1])
        w[3, -1, 1:-1] = self.p_out / (1.4 - 1) + 0.5 * rho * (u ** 2 + v ** 2)

        # slip walls: reflect the normal momentum
        w[:, :, 0] = w[:, :, 1]
        if self.viscous:
            rhoU_n = xp.sum(w[1:3, 1:-1, 0] * geo.normal_j[:, :, 0], 0)
            w[1:3, 1:-1, 0] -= 2 * rhoU_n * geo.normal_j[:, :, 0]
        else:
            nwall = geo.normal_j[:, :, 0]
            nwall = xp.hstack([nwall[:, :1], nwall, nwall[:, -1:]])
            rhoU_n = xp.sum(w[1:3, :, 0] * nwall, 0)
            w[1:3, :, 0] -= 2 * rhoU_n * nwall

        w[:, :, -1] = w[:, :, -2]
        if self.viscous:
            rhoU_n = xp.sum(w[1:3, 1:-1, -1] * geo.normal_j[:, :, -1], 0)
            w[1:3, 1:-1, -1] -= 2 * rhoU_n * geo.normal_j[:, :, -1]
        else:
            nwall = geo.normal_j[:, :, -1]
            nwall = xp.hstack([nwall[:, :1], nwall, nwall[:, -1:]])
            rhoU_n = xp.sum(w[1:3, :, -1] * nwall, 0)
            w[1:3, :, -1] -= 2 * rhoU_n * nwall
        return w

    # ---- reference euler_kec.py:66-69 / navierstokes.py:86-100 ---- #
    def flux(self, rho, u, v, E, p, grad_u=None):
        xp = self.xp
        if grad_u is None:
            F = xp.array([rho * u, rho * u ** 2 + p, rho * u * v, u * (E + p)])
            G = xp.array([rho * v, rho * u * v, rho * v ** 2 + p, v * (E + p)])
            return F, G
        dudx, dudy, dvdx, dvdy = grad_u
        mu, mu_t = self.mu, 0
        sigma_xx = (mu + mu_t) * (2 * dudx - 2. / 3 * (dudx + dvdy))
        sigma_yy = (mu + mu_t) * (2 * dvdy - 2. / 3 * (dudx + dvdy))
        sigma_xy = (mu + mu_t) * (dudy + dvdx)
        F = xp.array([rho * u, rho * u ** 2 + p - sigma_xx,
                      rho * u * v - sigma_xy,
                      u * (E + p) - sigma_xx * u - sigma_xy * v])
        G = xp.array([rho * v, rho * u * v - sigma_xy,
                      rho * v ** 2 + p - sigma_yy,
                      v * (E + p) - sigma_xy * u - sigma_yy * v])
        return F, G

    # ---- reference euler_kec.py:71-86 ---- #
    def sponge_flux(self, c_ext, w_ext):
        xp, geo = self.xp, self.geo
        ci = 0.5 * (c_ext[1:, 1:-1] + c_ext[:-1, 1:-1])
        cj = 0.5 * (c_ext[1:-1, 1:] + c_ext[1:-1, :-1])
        a = geo.area
        ai = xp.vstack([a[:1, :], (a[1:, :] + a[:-1, :]) / 2, a[-1:, :]])
        aj = xp.hstack([a[:, :1], (a[:, 1:] + a[:, :-1]) / 2, a[:, -1:]])
        wxx = (w_ext[:, 2:, 1:-1] + w_ext[:, :-2, 1:-1] - 2 * w_ext[:, 1:-1, 1:-1]) / 3.
        wyy = (w_ext[:, 1:-1, 2:] + w_ext[:, 1:-1, :-2] - 2 * w_ext[:, 1:-1, 1:-1]) / 3.
        # second-order dissipation at the boundary faces, fourth-order inside
        Fi = -0.5 * ci * ai * (w_ext[:, 1:, 1:-1] - w_ext[:, :-1, 1:-1])
        Fi[:, 1:-1, :] = 0.5 * (ci * ai)[1:-1, :] * (wxx[:, 1:, :] - wxx[:, :-1, :])
        Fj = -0.5 * cj * aj * (w_ext[:, 1:-1, 1:] - w_ext[:, 1:-1, :-1])
        Fj[:, :, 1:-1] = 0.5 * (cj * aj)[:, 1:-1] * (wyy[:, :, 1:] - wyy[:, :, :-1])
        return Fi, Fj

    # ---- reference navierstokes.py:63-84: nodal gradient on the dual mesh ---- #
    def grad_dual(self, phi):
        xp, geo = self.xp, self.geo
        dxy_i = 0.5 * (geo.dxy_i[:, 1:, :] + geo.dxy_i[:, :-1, :])
        dxy_j = 0.5 * (geo.dxy_j[:, :, 1:] + geo.dxy_j[:, :, :-1])
        phi_i = xp.array([phi * dxy_i[1], -phi * dxy_i[0]])
        phi_j = xp.array([-phi * dxy_j[1], phi * dxy_j[0]])
        grad_phi = xp.zeros(geo.xy.shape)
        grad_phi[:, :-1, :-1] += phi_i + phi_j
        grad_phi[:, 1:, :-1] += -phi_i + phi_j
        grad_phi[:, :-1, 1:] += phi_i - phi_j
        grad_phi[:, 1:, 1:] += -phi_i - phi_j
        area = xp.zeros(geo.xy.shape[1:])
        area[:-1, :-1] += geo.area
        area[1:, :-1] += geo.area
        area[:-1, 1:] += geo.area
        area[1:, 1:] += geo.area
        return 2 * grad_phi / area

    # ---- reference euler_kec.py:88-117 / navierstokes.py:119-157 ---- #
    def residual(self, w, w0, dt):
        xp, geo = self.xp, self.geo
        w_ext = self.extend(w)
        rho, u, v, E, p = self.primitive(w_ext)
        c = xp.sqrt(1.4 * p / rho)
        if self.viscous:
            dudx, dudy = self.grad_dual(u[1:-1, 1:-1])
            dvdx, dvdy = self.grad_dual(v[1:-1, 1:-1])
            duv_dxy = xp.array([dudx, dudy, dvdx, dvdy])
        # interface averages
        rho_i = 0.5 * (rho[1:, 1:-1] + rho[:-1, 1:-1])
        rho_j = 0.5 * (rho[1:-1, 1:] + rho[1:-1, :-1])
        u_i = 0.5 * (u[1:, 1:-1] + u[:-1, 1:-1])
        u_j = 0.5 * (u[1:-1, 1:] + u[1:-1, :-1])
        v_i = 0.5 * (v[1:, 1:-1] + v[:-1, 1:-1])
        v_j = 0.5 * (v[1:-1, 1:] + v[1:-1, :-1])
        E_i = 0.5 * (E[1:, 1:-1] + E[:-1, 1:-1])
        E_j = 0.5 * (E[1:-1, 1:] + E[1:-1, :-1])
        p_i = 0.5 * (p[1:, 1:-1] + p[:-1, 1:-1])
        p_j = 0.5 * (p[1:-1, 1:] + p[1:-1, :-1])
        if self.viscous:
            duv_dxy_i = 0.5 * (duv_dxy[:, :, 1:] + duv_dxy[:, :, :-1])
            duv_dxy_j = 0.5 * (duv_dxy[:, 1:, :] + duv_dxy[:, :-1, :])
            duv_dxy_i[:, [0, -1], :] = 0       # no viscous stress at inlet / outlet
            F_i, G_i = self.flux(rho_i, u_i, v_i, E_i, p_i, duv_dxy_i)
            F_j, G_j = self.flux(rho_j, u_j, v_j, E_j, p_j, duv_dxy_j)
        else:
            F_i, G_i = self.flux(rho_i, u_i, v_i, E_i, p_i)
            F_j, G_j = self.flux(rho_j, u_j, v_j, E_j, p_j)
        Fi = + F_i * geo.dxy_i[1] - G_i * geo.dxy_i[0]
        Fj = - F_j * geo.dxy_j[1] + G_j * geo.dxy_j[0]
        Fi_s, Fj_s = self.sponge_flux(c, w_ext)
        Fi += 0.5 * Fi_s
        Fj += 0.5 * Fj_s
        divF = (Fi[:, 1:, :] - Fi[:, :-1, :] + Fj[:, :, 1:] - Fj[:, :, :-1]) / geo.area
        return (w - w0) / dt + xp.ravel(divF)
