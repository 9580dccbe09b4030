"""Small differentiable programs exercising every tape op of SURVEY Appendix A.

Each case is ``fn(xp, rng) -> (f, u)``: it builds inputs from the seeded
``rng``, records a short program on the namespace ``xp`` and returns the
output array ``f`` and the independent array ``u``.  The same cases run on

  * the unmodified reference (tests/golden/make_golden.py -> fixtures),
  * the CPU oracle (tests/test_oracle_golden.py),
  * the CUDA engine through the C-ABI (tests/test_gpu_parity.py).

They restate what the reference's embedded unit tests cover
(adarray.py:929-1388: manipulation, indexing, operations, Poisson stencils).
"""
import numpy as np


def _u(xp, rng, shape):
    return xp.array(rng.standard_normal(shape))


def c_add_same(xp, rng):
    u = _u(xp, rng, (4, 3))
    b = _u(xp, rng, (4, 3))
    return u + b + 2.5, u


def c_add_self(xp, rng):
    u = _u(xp, rng, (7,))
    return u + u, u


def c_sub_rsub(xp, rng):
    u = _u(xp, rng, (5,))
    return (3.0 - u) - (u - 1.0), u


def c_mul_scalar(xp, rng):
    u = _u(xp, rng, (6,))
    return -1.5 * u * 4, u


def c_mul_same(xp, rng):
    u = _u(xp, rng, (3, 4))
    b = _u(xp, rng, (3, 4))
    return u * b * u, u


def c_mul_ndarray(xp, rng):
    u = _u(xp, rng, (3, 4))
    b = rng.standard_normal((3, 4))
    return b * u + u * b, u


def c_div(xp, rng):
    u = _u(xp, rng, (5,))
    b = xp.array(rng.uniform(1, 2, 5))
    return u / b + 2.0 / (u * u + 1.0), u


def c_pow(xp, rng):
    u = xp.array(rng.uniform(0.5, 2, (2, 5)))
    return u ** 3 + xp.sqrt(u) + u ** -1, u


def c_pow_nonfinite(xp, rng):
    v = rng.uniform(0.5, 2, 6)
    v[2] = 0.0
    u = xp.array(v)
    return u ** 0.5, u


def c_rpow(xp, rng):
    u = _u(xp, rng, (4,))
    return 2.0 ** u, u


def c_unary(xp, rng):
    u = xp.array(rng.uniform(0.5, 1.5, (3, 3)))
    return xp.exp(u) * xp.sin(u) + xp.cos(u) * xp.log(u) + xp.tanh(u), u


def c_smooth(xp, rng):
    u = _u(xp, rng, (6,))
    b = _u(xp, rng, (6,))
    return xp.maximum_smooth(u, b) + xp.minimum_smooth(u, b, 0.3) + xp.sigmoid(u), u


def c_bcast_add(xp, rng):
    u = _u(xp, rng, (4,))
    b = _u(xp, rng, (3, 4))
    return u + b, u


def c_bcast_add_big(xp, rng):
    u = _u(xp, rng, (3, 4))
    b = _u(xp, rng, (4,))
    return (u + b) + b, u


def c_bcast_mul(xp, rng):
    u = _u(xp, rng, (3, 1))
    b = _u(xp, rng, (3, 4))
    return u * b, u


def c_bcast_mul_big(xp, rng):
    u = _u(xp, rng, (2, 3, 4))
    b = _u(xp, rng, (3, 4))
    return b * u * b, u


def c_scalar_bcast(xp, rng):
    u = _u(xp, rng, ())
    b = _u(xp, rng, (5,))
    return u * b + u, u


def c_getitem_slices(xp, rng):
    u = _u(xp, rng, (5, 6))
    return u[1:, 1:-1] + u[:-1, 1:-1] - 2 * u[1:, 2:], u


def c_getitem_step_neg(xp, rng):
    u = _u(xp, rng, (9,))
    return u[::2] + u[::-2], u


def c_getitem_int(xp, rng):
    u = _u(xp, rng, (4, 5))
    return u[2] * u[:, 3][:4].reshape((4, 1))[:, 0].sum() + u[1, 2], u


def c_getitem_list(xp, rng):
    u = _u(xp, rng, (7,))
    return u[[0, 3, 3, 6]] * u[[1, 1, 2, 5]], u


def c_getitem_bool(xp, rng):
    u = _u(xp, rng, (8,))
    m = np.array([1, 0, 1, 1, 0, 0, 1, 0], bool)
    return u[m] ** 2, u


def c_getitem_newaxis(xp, rng):
    u = _u(xp, rng, (4,))
    return u[:, xp.newaxis] * u[xp.newaxis, :], u


def c_setitem_slices(xp, rng):
    u = _u(xp, rng, (12,))
    w = xp.zeros([5, 6])
    w[1:-1, 1:-1] = u.reshape([3, 4])
    w[:, 0] = 1 - w[:, 1]
    w[0, :] = -w[1, :]
    w[-1, :] = w[-2, :] * 3
    return w, u


def c_setitem_scalar(xp, rng):
    u = _u(xp, rng, (3, 4))
    w = u * 2
    w[1, 1:3] = 0
    w[2, 0] = u[0, 0]
    return w, u


def c_setitem_bcast(xp, rng):
    u = _u(xp, rng, (4,))
    w = xp.zeros([3, 4])
    w[:, :] = u
    w[1] = u * u
    return w, u


def c_setitem_list(xp, rng):
    u = _u(xp, rng, (2, 5, 3))
    w = u * 1.0
    w[:, [0, -1], :] = 0
    return w * w, u


def c_iadd_imul(xp, rng):
    u = _u(xp, rng, (5,))
    b = _u(xp, rng, (5,))
    w = u * 1
    w += b
    w += u
    w *= 3
    w *= u
    w -= 2 * b
    return w, u


def c_aug_sub(xp, rng):
    u = _u(xp, rng, (2, 4))
    w = xp.zeros([3, 5])
    w[:-1, :-1] += u
    w[1:, 1:] += u * u
    w[:2, 1:] -= 2 * u
    return w, u


def c_reshape_ravel_copy(xp, rng):
    u = _u(xp, rng, (3, 4))
    return xp.ravel(u.reshape([2, 6]) * 2).copy() + xp.copy(u).reshape([12]), u


def c_transpose(xp, rng):
    u = _u(xp, rng, (3, 4))
    return u.T * 2 + xp.transpose(u * u), u


def c_transpose_axes(xp, rng):
    u = _u(xp, rng, (2, 3, 4))
    return xp.transpose(u, (2, 0, 1)), u


def c_concatenate(xp, rng):
    u = _u(xp, rng, (3, 2))
    b = _u(xp, rng, (3, 3))
    return xp.concatenate([u, b, u * 2], axis=1), u


def c_hstack_vstack(xp, rng):
    u = _u(xp, rng, (4,))
    a = xp.hstack([u, u * u])
    b = xp.vstack([u, 2 * u, u ** 2])
    return xp.hstack([a, xp.ravel(b)]), u


def c_vstack_2d(xp, rng):
    u = _u(xp, rng, (2, 3))
    return xp.vstack([u[:1, :], (u[1:, :] + u[:-1, :]) / 2, u[-1:, :]]), u


def c_array_list(xp, rng):
    u = _u(xp, rng, (3,))
    return xp.array([u, u * u, [u * 2, u + 1][0]]), u


def c_array_nested(xp, rng):
    u = _u(xp, rng, (2, 3))
    return xp.array([u * u, -u]), u


def c_sum_all(xp, rng):
    u = _u(xp, rng, (3, 4))
    return xp.sum(u * u), u


def c_sum_axis(xp, rng):
    u = _u(xp, rng, (2, 3, 4))
    return xp.sum(u, 0) * xp.sum(u * u, axis=0), u


def c_sum_axis1(xp, rng):
    u = _u(xp, rng, (3, 4))
    return xp.sum(u, axis=1) + u.sum(1), u


def c_mean(xp, rng):
    u = _u(xp, rng, (4, 5))
    return xp.mean(u, 0) + xp.mean(u), u


def c_roll(xp, rng):
    u = _u(xp, rng, (3, 4))
    return xp.roll(u, 1, axis=1) - xp.roll(u, -1, axis=0) + xp.roll(u, 2), u


def c_rollaxis(xp, rng):
    u = _u(xp, rng, (2, 3, 4))
    return xp.rollaxis(u, 2), u


def c_dot(xp, rng):
    u = _u(xp, rng, (3, 4))
    b = _u(xp, rng, (4, 2))
    return xp.dot(u, b), u


def c_dot_vec(xp, rng):
    u = _u(xp, rng, (4,))
    b = _u(xp, rng, (3, 4))
    return xp.dot(b, u), u


def c_sort(xp, rng):
    u = _u(xp, rng, (6,))
    w = u * u + u
    w.sort()
    return w, u


def c_meshgrid(xp, rng):
    u = _u(xp, rng, (3,))
    b = _u(xp, rng, (4,))
    x, y = xp.meshgrid(u, b)
    return x * y + x, u


def c_no_dependence(xp, rng):
    u = _u(xp, rng, (4,))
    b = _u(xp, rng, (3,))
    return b * b, u


def c_identity(xp, rng):
    u = _u(xp, rng, (5,))
    return u, u


def c_scaled_identity(xp, rng):
    u = _u(xp, rng, (5,))
    b = _u(xp, rng, (5,))
    return u + b, u


def c_cancel(xp, rng):
    """x - x: the sum prunes the exactly-cancelled entries."""
    u = _u(xp, rng, (6,))
    return (u * 3 - 3 * u) + u[::-1] * 0.0, u


def c_zero_diag(xp, rng):
    """zero diagonal entries: the blocks carry explicit zeros, scipy's product
    and sum both drop them (value-dependent pattern)."""
    u = xp.array(np.array([0.0, 1.0, 0.0, 2.0]))
    w = u * u
    return w + w[::-1] * u, u


def c_poisson1d(xp, rng):
    u = _u(xp, rng, (9,))
    dx = 0.1
    res = (u[2:] + u[:-2] - 2 * u[1:-1]) / dx ** 2
    return xp.hstack([u[:1], res, u[-1:]]), u


def c_poisson2d(xp, rng):
    N, M = 5, 4
    u = _u(xp, rng, (N * M,))
    w = xp.zeros([N + 2, M + 2])
    w[1:-1, 1:-1] = u.reshape([N, M])
    res = (w[2:, 1:-1] + w[:-2, 1:-1] + w[1:-1, 2:] + w[1:-1, :-2]
           - 4 * w[1:-1, 1:-1]) * 16.0
    return xp.ravel(res) + xp.exp(u), u


def c_poisson3d(xp, rng):
    N = 3
    u = _u(xp, rng, (N, N, N))
    w = xp.zeros([N + 2] * 3)
    w[1:-1, 1:-1, 1:-1] = u
    res = (w[2:, 1:-1, 1:-1] + w[:-2, 1:-1, 1:-1] + w[1:-1, 2:, 1:-1]
           + w[1:-1, :-2, 1:-1] + w[1:-1, 1:-1, 2:] + w[1:-1, 1:-1, :-2]
           - 6 * w[1:-1, 1:-1, 1:-1])
    return res, u


def c_burgers(xp, rng):
    u = _u(xp, rng, (8,))
    f = 0.5 * u ** 2
    up = xp.maximum_smooth(u[1:], u[:-1])
    return (f[1:] - f[:-1]) * up + xp.hstack([u[:1], u[-1:]]).sum(), u


CASES = {k[2:]: v for k, v in sorted(globals().items()) if k.startswith('c_')}
