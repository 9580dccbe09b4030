"""numpad_b200 -- B200-native engine behind numpad's hot path.

Drop-in for `from numpad import *` (reference __init__.py:1-7): the same names
are exported; Jacobian assembly (`diff`) and the Newton linear solve (`solve`)
run as hand-written sm_100a CUDA kernels behind the C-ABI of
include/numpad_b200.h.  Importing works without a GPU (so that the ABI can be
checked); computing does not -- there is no CPU fallback.
"""
from . import _lib as _libmod
_libmod.lib()           # fail loudly at import if the CUDA library is missing

from .adarray import *          # noqa: F401,F403,E402
from .adsolve import *          # noqa: F401,F403,E402
from . import adarray as _adarray_mod   # noqa: E402
from .adarray import (np, sp, value, array, diff, diff_dev, adarray, sum, exp, sin, cos, log,  # noqa: F401,E402
                      tanh, sqrt, dot, copy, ravel, roll, rollaxis, transpose, mean,
                      zeros, ones, empty, eye, linspace, load, loadtxt, concatenate,
                      hstack, vstack, meshgrid, newaxis, sigmoid, gt_smooth, lt_smooth,
                      maximum_smooth, minimum_smooth, diff_func, replace__globals__,
                      adarray_count, adstate_count)
from .adsolve import (solve, solve_newton_with_dt, psuedo_time_continuation,  # noqa: F401,E402
                      adsolution, ResidualState, SolutionState)
from .adstate import _add_ops, _multiply_ops  # noqa: F401,E402
from .adtools import interp  # noqa: F401,E402
from . import adrandom as random  # noqa: F401,E402      (reference __init__.py:4-7)
from . import advisual as visual  # noqa: F401,E402
from . import adlinalg as linalg  # noqa: F401,E402
from . import adsparse as sparse  # noqa: F401,E402


def launch_count():
    """kernels launched through the C-ABI so far (bench.py: gpu_launches)"""
    return int(_libmod.lib().cdll.npb_launch_count())
