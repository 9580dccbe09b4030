"""1-D interpolation on adarrays (reference adtools.py:30-91): piecewise linear, or the C2
cubic spline whose knot slopes solve the natural-spline continuity equations through
`solve` -- so interpolated values are differentiable w.r.t. knots, data and abscissae."""
import numpy as np

from .adarray import array, value, zeros, hstack, ravel, transpose
from .adsolve import solve


def _spline_continuity(slopes, x, y):
    """second derivative jumps of the Hermite interpolant at the knots (zero curvature at
    both ends): the residual whose root are the spline's knot slopes"""
    h = x[1:] - x[:-1]
    secant = (y[1:] - y[:-1]) / h
    left = (6 * secant - 4 * slopes[:-1] - 2 * slopes[1:]) / h      # y'' at the left end of each interval
    right = (4 * slopes[1:] + 2 * slopes[:-1] - 6 * secant) / h     # ... at the right end
    return hstack([left[:1], left[1:] - right[:-1], -right[-1:]])


class interp:
    """y = interp(x0, y0, type='linear' | 'cubic');  y(x);  y.derivative(x)"""

    def __init__(self, x0, y0, type='linear'):
        knots = np.asarray(value(x0))
        assert (knots[1:] > knots[:-1]).all()
        x0, y0 = array(x0), array(y0)
        self.x0 = x0.copy()
        if type == 'linear':
            self.y0 = y0[:, np.newaxis].copy()
        elif type == 'cubic':
            slopes = solve(_spline_continuity, zeros(y0.size), (x0, y0), verbose=False)
            self.y0 = transpose([y0, slopes])
        else:
            raise ValueError('interp: unknown type {0}'.format(type))

    cspline_resid = staticmethod(_spline_continuity)

    def find(self, x):
        hi = np.clip(np.searchsorted(np.asarray(value(self.x0)), np.asarray(value(x))),
                     1, self.x0.size - 1)
        return self.x0[hi - 1], self.x0[hi], self.y0[hi - 1], self.y0[hi]

    def __call__(self, x):
        x = array(x)
        shape = x.shape
        x = ravel(x)
        xa, xb, ya, yb = self.find(x)
        s = (x - xa) / (xb - xa)
        t = 1 - s
        y = s * yb[:, 0] + t * ya[:, 0]
        if self.y0.shape[1] == 2:
            secant = (yb[:, 0] - ya[:, 0]) / (xb - xa)
            y += (s - 2 * s ** 2 + s ** 3) * (ya[:, 1] - secant) * (xb - xa) \
                + (t - 2 * t ** 2 + t ** 3) * (yb[:, 1] - secant) * (xa - xb)
        return y.reshape(shape)

    def derivative(self, x):
        x = array(x)
        xa, xb, ya, yb = self.find(x)
        dy = (yb[:, 0] - ya[:, 0]) / (xb - xa)
        if self.y0.shape[1] == 2:
            s = (x - xa) / (xb - xa)
            t = 1 - s
            secant = (yb[:, 0] - ya[:, 0]) / (xb - xa)
            dy += (1 - 4 * s + 3 * s ** 2) * (ya[:, 1] - secant) \
                + (1 - 4 * t + 3 * t ** 2) * (yb[:, 1] - secant)
        return dy
