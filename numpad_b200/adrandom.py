"""`numpad.random` (reference adrandom.py:27-29): numpy's generators returning adarrays."""
import numpy as np

from .adarray import array


def _wrap(name):
    fn = getattr(np.random, name)

    def gen(*args, **kargs):
        return array(np.asarray(fn(*args, **kargs), np.float64))
    gen.__name__ = name
    gen.__doc__ = 'Overloaded by numpad_b200, returns adarray.\n' + (fn.__doc__ or '')
    return gen


random = _wrap('random')            # the one generator the reference exports
rand = _wrap('rand')
randn = _wrap('randn')
standard_normal = _wrap('standard_normal')
seed = np.random.seed
RandomState = np.random.RandomState
