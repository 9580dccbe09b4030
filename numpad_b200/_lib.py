"""ctypes binding of libnumpad_b200.so (the C-ABI declared in include/numpad_b200.h).

The prototypes are read from the header itself, so the binding cannot drift
from the declared ABI.  There is NO fallback: if the shared library is missing
or a call fails, an exception is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, 'libnumpad_b200.so')
HEADER_PATH = os.path.join(_ROOT, 'include', 'numpad_b200.h')

MAX_SCALES = 6


class Scales(ctypes.Structure):
    _fields_ = [('n', ctypes.c_int), ('s', ctypes.c_double * MAX_SCALES)]


class KrylovOps(ctypes.Structure):
    _fields_ = [('matvec', ctypes.c_void_p), ('matvec_ctx', ctypes.c_void_p),
                ('precond', ctypes.c_void_p), ('precond_ctx', ctypes.c_void_p),
                ('allreduce', ctypes.c_void_p), ('allreduce_ctx', ctypes.c_void_p)]


class CsrView(ctypes.Structure):
    _fields_ = [('nrows', ctypes.c_int64), ('nnz', ctypes.c_int64),
                ('indptr', ctypes.c_void_p), ('indices', ctypes.c_void_p),
                ('data', ctypes.c_void_p)]


class KrylovInfo(ctypes.Structure):
    _fields_ = [('iters', ctypes.c_int32), ('restarts', ctypes.c_int32),
                ('status', ctypes.c_int32), ('pad', ctypes.c_int32),
                ('bnorm', ctypes.c_double), ('rnorm0', ctypes.c_double),
                ('rnorm', ctypes.c_double)]


APPLY_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                            ctypes.c_void_p, ctypes.c_void_p)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_int, ctypes.c_void_p)

_SCALAR = {'int': ctypes.c_int, 'int64_t': ctypes.c_int64,
           'int32_t': ctypes.c_int32, 'double': ctypes.c_double,
           'unsigned long long': ctypes.c_ulonglong}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every npb_* prototype."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'//[^\n]*', ' ', src)
    protos = {}
    pat = re.compile(r'(?:^|;|\})\s*((?:const\s+)?(?:unsigned long long|int64_t|int|void|char|double)\s*\*?)\s*'
                     r'(npb_\w+)\s*\(([^()]*)\)\s*(?=;)', re.S)
    for m in pat.finditer(src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if 'char' in ret:
            restype = ctypes.c_char_p
        elif ret == 'void':
            restype = None
        else:
            restype = _SCALAR[ret]
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a or a.startswith('npb_apply_fn') or a.startswith('npb_allreduce_fn'):
                    argtypes.append(ctypes.c_void_p)
                else:
                    t = a.rsplit(' ', 1)[0].replace('const ', '').strip()
                    argtypes.append(_SCALAR[t])
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'numpad_b200: %s is missing -- build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` or '
                'numpad_b200/csrc/build.sh. There is no CPU fallback.' % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)      # AttributeError if not exported
            fn.restype = restype
            fn.argtypes = argtypes
        self.csr_matvec_cb_addr = ctypes.cast(self.cdll.npb_csr_matvec_cb,
                                              ctypes.c_void_p).value

    def last_error(self):
        return self.cdll.npb_last_error().decode()

    def call(self, name, *args):
        """Call an int-returning entry point; torch tensors -> device pointers,
        ctypes structures -> pointers.  Raises on a non-zero status."""
        conv = []
        for a in args:
            if a is None:
                conv.append(None)
            elif hasattr(a, 'data_ptr'):
                conv.append(a.data_ptr())
            elif isinstance(a, ctypes.Structure):
                conv.append(ctypes.addressof(a))
            else:
                conv.append(a)
        rc = getattr(self.cdll, name)(*conv)
        if rc != 0:
            raise RuntimeError('%s failed (%d): %s' % (name, rc, self.last_error()))
        return rc

    def raw(self, name):
        return getattr(self.cdll, name)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def make_scales(chain):
    sc = Scales()
    sc.n = len(chain)
    for i, v in enumerate(chain):
        sc.s[i] = float(v)
    return sc


NO_SCALES = make_scales(())
