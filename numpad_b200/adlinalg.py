"""`numpad.linalg` (reference adlinalg.py:14-21): dense solve, differentiable through the
solution by the implicit function theorem (the residual A x - b recorded on the tape)."""
import torch

from .adarray import adarray, array, dot
from .adsolve import adsolution


def solve(A, b):
    """AD counterpart of numpy.linalg.solve: x with A x = b for adarrays A (n x n), b (n or
    n x k).  The dense LU runs on the device (library call: nothing here is on the Newton
    hot path); d x / d A and d x / d b follow from `adsolution` like any converged solve."""
    A, b = array(A), array(b)
    assert A.ndim == 2 and b.shape[0] == A.shape[0]
    x = adarray(torch.linalg.solve(A._dev, b._dev))
    return adsolution(x, dot(A, x) - b, 1)
