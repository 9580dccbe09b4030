"""CPU oracle for the numpad hot path (TEST INFRASTRUCTURE -- NOT A PRODUCT PATH).

This file is a from-scratch restatement, in numpy + scipy.sparse, of the
algorithm the reference (qiqi/numpad, pure Python) runs on its hot path:

  * every array operation records its local sparse Jacobian on a tape
    (reference adarray.py:470-807, adstate.py:42-140),
  * ``diff`` accumulates those blocks by the chain rule in tangent or
    adjoint order with scipy's CSR product / sum
    (adstate.py:142-229, 282-297; adarray.py:846-862),
  * ``solve`` is Newton's method whose linear solve is SuperLU ``spsolve``
    and whose result differentiates through the solve by the implicit
    function theorem (adsolve.py:36-264).

All arithmetic of the reference on this path lives in third-party code that is
not vendored in /root/reference (no version pinned there): scipy.sparse
``_sparsetools`` (csr_matmat, csr_plus_csr, coo_tocsr) and SuperLU via
``scipy.sparse.linalg``; this oracle calls the very same routines (the image
has numpy 2.3 / scipy 1.18), so its sparsity rules (sums and products drop
exactly-zero results, `number * csr` keeps them) are the reference's by
construction.

Parity pinning: ``tests/golden/make_golden.py`` imports the UNMODIFIED
reference from /root/reference (through a 3-line numpy-2 shim) and stores its
Jacobians / Newton iterates / adjoint vectors as fixtures; ``tests/
test_oracle_golden.py`` checks this oracle against those fixtures and against
the reference's own known-answer unit tests (adarray.py:929-1388,
adsolve.py:275-356).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.
"""
import itertools
import numbers
import weakref

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

newaxis = np.newaxis
_ids = itertools.count()

# counters used by bench.py / tests to report what one assembly costs
stats = {'mul': 0, 'add': 0, 'spsolve': 0}


# --------------------------------------------------------------------------- #
# lazy local-Jacobian blocks                      (reference adstate.py:233-278)
# --------------------------------------------------------------------------- #
class Diag:
    """n x n diagonal block; CSR built on first use and cached (adstate.py:233-250)."""

    def __init__(self, d):
        self.d = d
        self._m = None

    @property
    def shape(self):
        return (self.d.size, self.d.size)

    def tocsr(self):
        if self._m is None:
            n = self.d.size
            self._m = sp.csr_matrix((self.d, np.arange(n), np.arange(n + 1)))
        return self._m


class Coo:
    """(data, row, col) block; CSR (duplicates summed) on first use (adstate.py:252-272)."""

    def __init__(self, data, row, col, shape=None):
        self.data, self.row, self.col = data, row, col
        self._shape = shape
        self._m = None

    @property
    def shape(self):
        if self._shape is None:
            self._shape = (self.row.max() + 1, self.col.max() + 1)
        return self._shape

    def tocsr(self):
        if self._m is None:
            self._m = sp.csr_matrix((self.data, (self.row, self.col)),
                                    shape=self._shape)
        return self._m


def _csr(b):
    return b.tocsr() if isinstance(b, (Diag, Coo)) else b


def _int0(x):
    """The reference tests ``x is 0`` (adstate.py:283-291): true for the int
    literal only -- a float 0.0 multiplier keeps its pattern."""
    return type(x) is int and x == 0


def jac_add(a, b):
    """adstate.py:282-288 -- int 0 is the additive identity; CSR+CSR prunes zeros."""
    if _int0(a):
        return b
    if _int0(b):
        return a
    stats['add'] += 1
    return _csr(a) + _csr(b)


def jac_mul(a, b):
    """adstate.py:290-297 -- 0 annihilates; an all-empty product collapses to 0."""
    if _int0(a) or _int0(b):
        return 0
    stats['mul'] += 1
    c = _csr(a) * _csr(b)
    if hasattr(c, 'nnz') and c.nnz == 0:
        c = 0
    return c


# --------------------------------------------------------------------------- #
# tape nodes                                        (reference adstate.py:42-177)
# --------------------------------------------------------------------------- #
class Node:
    """One state of an array's history.

    unary  op : ``mult`` = d(self)/d(prev), ``other`` is None
    binary op : d(self)/d(prev) = I, ``mult`` = d(self)/d(other)
    Forward edges (``nxt``, ``_to``) are weak so dead branches vanish from
    the adjoint sweep exactly as in the reference (adstate.py:99-107).
    """

    def __init__(self, owner, prev=None, mult=None, other=None, tag=''):
        self.id = next(_ids)
        self.owner = weakref.ref(owner)
        self.size = owner.size
        self.tag = tag
        self.prev = prev
        self.nxt = None
        self._to = []
        self.other = None
        if prev is not None:
            prev.nxt = weakref.ref(self)
        if mult is None:
            assert prev is None and other is None
        else:
            self.mult = mult
            if other is None:
                if not isinstance(mult, numbers.Number):
                    assert mult.shape == (self.size, self.size)
            else:
                if not isinstance(mult, numbers.Number):
                    assert mult.shape == (self.size, other.size)
                other._to.append(weakref.ref(self))
                self.other = other

    def __hash__(self):
        return self.id

    def __lt__(self, o):
        return self.id < o.id

    def __eq__(self, o):
        return self.id == o.id

    def dependers(self):
        if self.nxt is not None and self.nxt() is not None:
            yield self.nxt()
        for r in self._to:
            if r() is not None:
                yield r()

    def dependees(self):
        if self.prev:
            yield self.prev
        if self.other:
            yield self.other

    def forget(self):
        """adstate.py:118-129"""
        for s in self.dependees():
            s._to = [r for r in s._to if r() is not self]
        self.prev = None
        self.other = None
        if hasattr(self, 'mult') and self.mult:
            del self.mult

    def push(self, mult, other=None, tag=''):
        return Node(self.owner(), self, mult, other, tag)

    # tangent / adjoint rules, adstate.py:142-177
    def tangent(self, dep_du):
        if self.prev is None:
            return 0
        if self.other is None:
            (p,) = dep_du
            return jac_mul(self.mult, p)
        p, o = dep_du
        return jac_add(p, jac_mul(self.mult, o))

    def adjoint(self, f_d_dependers):
        acc = 0
        for s, f_ds in zip(self.dependers(), f_d_dependers):
            if s.other is self:
                local = s.mult
            elif s.prev is self and s.other:
                local = 1
            else:
                assert s.prev is self
                local = s.mult
            acc = jac_add(acc, jac_mul(f_ds, local))
        return acc


def sweep_tangent(f, u):
    """adstate.py:181-204"""
    d = {}
    todo = [f]
    while todo:
        s = todo.pop(0)
        if s not in d:
            d[s] = 0
            todo.extend(s.dependees())
    for s in sorted(d):
        if s is u:
            d[s] = sp.eye(u.size, u.size)
        else:
            d[s] = s.tangent(d[t] for t in s.dependees())
    return d[f]


def sweep_adjoint(f, u):
    """adstate.py:206-229"""
    d = {}
    todo = [u]
    while todo:
        s = todo.pop(0)
        if s not in d:
            d[s] = 0
            todo.extend(s.dependers())
    for s in sorted(d, reverse=True):
        if s is f:
            d[s] = sp.eye(f.size, f.size)
        else:
            d[s] = s.adjoint(d[t] for t in s.dependers())
    return d[u]


# --------------------------------------------------------------------------- #
# the array front end                             (reference adarray.py:467-807)
# --------------------------------------------------------------------------- #
def value(a):
    if isinstance(a, (numbers.Number, np.ndarray, list)):
        return a
    return a._value


class oarray:
    dtype = np.dtype('float64')
    __array_ufunc__ = None          # ndarray (op) oarray -> oarray.__r(op)__

    def __init__(self, a):
        self._value = np.asarray(value(a), np.float64)
        self._ind = np.arange(self.size).reshape(self.shape)
        self._current_state = Node(self)

    def __array__(self, *a, **k):
        return self._value

    size = property(lambda s: s._value.size)
    shape = property(lambda s: s._value.shape)
    ndim = property(lambda s: s._value.ndim)
    T = property(lambda s: transpose(s))

    def __len__(self):
        return len(self._value)

    @property
    def _initial_state(self):
        s = self._current_state
        while s.prev:
            s = s.prev
        return s

    def _ind_as(self, shape):
        """adarray.py:478-484"""
        ind = np.zeros(shape, dtype=int)
        if ind.ndim:
            ind[:] = self._ind
        else:
            ind = self._ind
        return ind

    def push(self, mult, other=None, tag=''):
        if other is None:
            self._current_state = self._current_state.push(mult, None, tag)
        elif isinstance(other, oarray):
            self._current_state = self._current_state.push(
                mult, other._current_state, tag)
        else:
            raise NotImplementedError

    def obliviate(self):
        self._current_state.forget()

    def copy(self):
        return copy(self)

    def transpose(self, axes=None):
        return transpose(self, axes)

    def reshape(self, shape):
        out = oarray(self._value.reshape(shape))
        if self.size > 0:
            out.push(1, self, 'reshape')
        return out

    def sort(self, axis=-1, kind='quicksort'):
        order = np.argsort(self._value, axis, kind)
        self._value.sort(axis, kind)
        j = np.ravel(self._ind[order])
        self.push(Coo(np.ones(j.size), np.arange(j.size), j,
                      shape=(j.size, self.size)), tag='sort')

    # ---- arithmetic, adarray.py:580-757 ---- #
    def __add__(self, a):
        if isinstance(a, numbers.Number):
            out = oarray(self._value + a)
            out.push(1, self, '+')
            return out
        b = self
        out = oarray(value(a) + value(b))
        if a.shape == b.shape:
            if hasattr(a, '_value'):
                out.push(1, a, '+')
            if hasattr(b, '_value'):
                out.push(1, b, '+')
        else:
            ones_ = np.ones(out.shape)
            rows = np.arange(out.size)
            for x in (a, b):
                if hasattr(x, '_value') and ones_.size > 0:
                    cols = np.ravel(x._ind_as(out.shape))
                    out.push(Coo(np.ravel(ones_), rows, cols,
                                 shape=(out.size, x.size)), x, '+')
        return out

    __radd__ = __add__

    def __iadd__(self, a):
        if isinstance(a, (numbers.Number, np.ndarray)):
            self._value += a
        else:
            self._value += a._value
            if a.shape == self.shape:
                self.push(1, a, '+')
            else:
                raise NotImplementedError
        return self

    def __neg__(self):
        out = oarray(-self._value)
        out.push(-1, self, '-')
        return out

    def __pos__(self):
        return self

    def __sub__(self, a):
        return self.__add__(-a)

    def __rsub__(self, a):
        return (-self) + a

    def __isub__(self, a):
        return self.__iadd__(-a)

    def __mul__(self, a):
        if isinstance(a, numbers.Number):
            out = oarray(self._value * a)
            out.push(a, self, '*')
            return out
        b = self
        out = oarray(value(a) * value(b))
        if a.shape == b.shape:
            if hasattr(a, '_value'):
                out.push(Diag(value(b).ravel()), a, '*')
            if hasattr(b, '_value'):
                out.push(Diag(value(a).ravel()), b, '*')
        else:
            wa = np.zeros(out.shape)
            wb = np.zeros(out.shape)
            if wa.ndim:
                wa[:] = value(b)
            else:
                wa = value(b).copy()
            if wb.ndim:
                wb[:] = value(a).copy()
            else:
                wb = value(a).copy()
            rows = np.arange(out.size)
            for x, w in ((a, wa), (b, wb)):
                if hasattr(x, '_value') and out.size > 0:
                    cols = np.ravel(x._ind_as(out.shape))
                    out.push(Coo(np.ravel(w), rows, cols,
                                 shape=(out.size, x.size)), x, '*')
        return out

    __rmul__ = __mul__

    def __imul__(self, a):
        if isinstance(a, numbers.Number):
            self._value *= a
            self.push(a, tag='*')
        else:
            self.push(Diag(np.ravel(a._value.copy())), tag='*')
            self.push(Diag(np.ravel(self._value.copy())), a, '*')
            self._value *= a._value
        return self

    def __truediv__(self, a):
        return self * a ** (-1)

    def __rtruediv__(self, a):
        return a * self ** (-1)

    def __pow__(self, p):
        if not isinstance(p, numbers.Number):
            return NotImplemented
        out = oarray(self._value ** p)
        d = p * np.ravel(self._value) ** (p - 1)
        if d.size > 0:
            d[~np.isfinite(d)] = 0
            out.push(Diag(d), self, '**')
        return out

    def __rpow__(self, a):
        return exp(self * log(a))

    def sum(self, axis=None, dtype=None, out=None):
        return sum(self, axis)

    def mean(self, axis=None, dtype=None, out=None):
        return mean(self, axis)

    # ---- indexing, adarray.py:761-796 ---- #
    def __getitem__(self, ind):
        out = oarray(self._value[ind])
        j = np.ravel(self._ind[ind])
        if j.size > 0:
            out.push(Coo(np.ones(j.size), np.arange(j.size), j,
                         shape=(j.size, self.size)), self, '[]')
        return out

    def __setitem__(self, ind, a):
        a = array(a)
        keep = np.ones(self.size)
        keep[np.ravel(self._ind[ind])] = 0
        self.push(Diag(keep), tag='[]=0')
        self._value.__setitem__(ind, value(a))
        if hasattr(a, '_value'):
            i = self._ind[ind]
            if i.size > 0:
                j = a._ind_as(i.shape)
                i, j = np.ravel(i), np.ravel(j)
                self.push(Coo(np.ones(j.size), i, j,
                              shape=(self.size, a.size)), a, '[]')

    def __str__(self):
        return str(self._value)

    def __repr__(self):
        return 'o' + repr(self._value)

    def diff(self, u, mode='auto'):
        return diff(self, u, mode)


adarray = oarray      # so workload code written against the numpad names runs here


# ---- constructors and elementwise functions, adarray.py:148-289 ---- #
def array(a):
    """adarray.py:293-323"""
    if isinstance(a, oarray):
        return a
    if isinstance(a, (numbers.Number, np.ndarray)):
        return oarray(a)
    if isinstance(a, (list, tuple)):
        parts = [array(x) for x in a]
        out = oarray(np.array([p._value for p in parts]))
        for k, p in enumerate(parts):
            cols = np.arange(p.size)
            out.push(Coo(np.ones(p.size), k * p.size + cols, cols,
                         shape=(out.size, p.size)), p, 'array')
        return out


def zeros(*a, **k):
    return array(np.zeros(*a, **k))


def ones(*a, **k):
    return array(np.ones(*a, **k))


def empty(*a, **k):
    return array(np.empty(*a, **k))


def eye(*a, **k):
    return array(np.eye(*a, **k))


def linspace(*a, **k):
    return array(np.linspace(*a, **k))


def _elementwise(fn, dfn, tag):
    def op(x, out=None):
        x = array(x)
        if out is None:
            out = oarray(fn(x._value))
        else:
            fn(x._value, out._value)
            out.push(0, tag='0')
        out.push(Diag(dfn(np.ravel(x._value))), x, tag)
        return out
    op.__name__ = tag
    return op


exp = _elementwise(np.exp, np.exp, 'exp')
sin = _elementwise(np.sin, np.cos, 'sin')
cos = _elementwise(np.cos, lambda v: -np.sin(v), 'cos')
log = _elementwise(np.log, lambda v: 1. / v, 'log')
tanh = _elementwise(np.tanh, lambda v: 1 - np.tanh(v) ** 2, 'tanh')


def sqrt(x):
    return x ** (0.5)


def sigmoid(x):
    return (tanh(x) + 1) / 2


def gt_smooth(a, b, c=0.1):
    return sigmoid((a - b) / c)


def lt_smooth(a, b, c=0.1):
    return sigmoid((b - a) / c)


def maximum_smooth(a, b, c=0.1):
    g = gt_smooth(a, b, c)
    return a * g + b * (1. - g)


def minimum_smooth(a, b, c=0.1):
    return -maximum_smooth(-a, -b, c)


# ---- shape operations, adarray.py:325-463 ---- #
def ravel(a):
    a = array(a)
    return a.reshape((a.size,))


def copy(a):
    out = oarray(np.copy(value(a)))
    if isinstance(a, oarray):
        out.push(1, a, 'cpy')
    else:
        assert isinstance(a, np.ndarray)
    return out


def transpose(a, axes=None):
    a = array(a)
    out = oarray(np.transpose(a._value, axes))
    i = np.arange(a.size).reshape(a.shape)
    j = np.transpose(i, axes)
    out.push(Coo(np.ones(i.size), np.ravel(i), np.ravel(j)), a, 'T')
    return out


def concatenate(arrs, axis=0):
    arrs = [array(a) for a in arrs]
    vals, marks = [], []
    for a in arrs:
        if a.ndim == 0:
            a = a[np.newaxis]
        marks.append(len(vals) * np.ones_like(a._value))
        vals.append(value(a))
    out = oarray(np.concatenate(vals, axis))
    mark = np.ravel(np.concatenate(marks, axis))
    for k, a in enumerate(arrs):
        rows = (mark == k).nonzero()[0]
        out.push(Coo(np.ones(rows.size, int), rows, np.arange(rows.size),
                     shape=(mark.size, rows.size)), a, 'cat')
    return out


def hstack(arrs):
    nd = max(array(a).ndim for a in arrs)
    return concatenate(arrs, axis=1 if nd > 1 else 0)


def vstack(arrs):
    nd = max(array(a).ndim for a in arrs)
    if nd <= 1:
        return array(arrs)
    return concatenate(arrs, axis=0)


def meshgrid(x, y):
    ix, iy = np.meshgrid(x._ind, y._ind)
    return x[ix], y[iy]


def sum(a, axis=None, dtype=None, out=None, keepdims=False):
    assert dtype is None and out is None
    a = array(a)
    res = oarray(np.sum(a._value, axis, keepdims=keepdims))
    kshape = np.sum(a._value, axis, keepdims=True).shape
    rows = np.arange(res.size).reshape(kshape)
    rows = np.ravel(rows + np.zeros_like(a._value, int))
    cols = np.ravel(a._ind)
    res.push(Coo(np.ones(rows.size, int), rows, cols,
                 shape=(res.size, a.size)), a, 'sum')
    return res


def mean(a, axis=None, dtype=None, out=None, keepdims=False):
    s = sum(a, axis, dtype, out, keepdims)
    return s * (float(s.size) / a.size)


def rollaxis(a, axis, start=0):
    out = oarray(np.rollaxis(a._value, axis, start))
    out.push(Coo(np.ones(a.size), np.ravel(np.rollaxis(a._ind, axis, start)),
                 np.ravel(a._ind)), a, 'rollaxis')
    return out


def roll(a, shift, axis=None):
    out = oarray(np.roll(a._value, shift, axis))
    out.push(Coo(np.ones(a.size), np.ravel(a._ind),
                 np.ravel(np.roll(a._ind, shift, axis))), a, 'roll')
    return out


def dot(a, b):
    ax = a.ndim - 1
    if b.ndim > 1:
        a = a.reshape(a.shape + ((1,) * (b.ndim - 1)))
        b = rollaxis(b, -2)
    b = b.reshape(((1,) * (a.ndim - b.ndim)) + b.shape)
    return sum(a * b, axis=ax)


# ---- differentiation front door, adarray.py:846-897 ---- #
def diff(f, u, mode='auto'):
    if mode == 'auto':
        mode = 'tangent' if u.size < f.size else 'adjoint'
    if mode == 'tangent':
        d = sweep_tangent(f._current_state, u._initial_state)
    elif mode == 'adjoint':
        d = sweep_adjoint(f._current_state, u._initial_state)
    else:
        raise NotImplementedError
    if isinstance(d, numbers.Number) and d == 0:
        d = sp.csr_matrix((f.size, u.size), dtype=float)
    return d


def diff_func(func, u, args=(), kargs={}):
    u = oarray(value(u).copy())
    return func(u, *args, **kargs).diff(u)


# --------------------------------------------------------------------------- #
# Newton + implicit differentiation                (reference adsolve.py:36-264)
# --------------------------------------------------------------------------- #
def _is_identity_pattern(m):
    n = m.shape[0]
    try:
        return (m.indices.size == n and m.indptr.size == n + 1 and
                all(m.indices == np.arange(n)) and
                all(m.indptr == np.arange(n + 1)))
    except (TypeError, AttributeError):
        return False


class ResidualNode(Node):
    """adsolve.py:36-86"""

    def __init__(self, prev):
        Node.__init__(self, prev.owner(), prev, 1, None)
        self.solution = None

    def dependers(self):
        if self.solution is not None and self.solution() is not None:
            yield self.solution()
        yield from Node.dependers(self)

    def adjoint(self, f_d_dependers):
        acc = 0
        it = iter(f_d_dependers)
        if self.solution is not None and self.solution() is not None:
            g = next(it)
            if not _int0(g):
                jt = self.solution().jacobian.T
                n = jt.shape[0]
                if _is_identity_pattern(jt):
                    assert g.shape[-1] == n
                    inv = jt.copy()
                    inv.data = 1. / np.array(inv.data)
                    acc = -g * inv
                else:
                    if hasattr(g, 'todense'):
                        g = g.todense()
                    g = np.array(g)
                    stats['spsolve'] += 1
                    lu = spla.factorized(jt.tocsc())
                    acc = np.array([-lu(row) for row in g]).reshape(g.shape)
                    acc = sp.csr_matrix(acc)
        return jac_add(acc, Node.adjoint(self, it))


class SolutionNode(Node):
    """adsolve.py:89-147"""

    def __init__(self, owner, residual_node, jacobian):
        Node.__init__(self, owner)
        assert isinstance(residual_node, ResidualNode)
        assert residual_node.size == self.size
        residual_node.solution = weakref.ref(self)
        self.residual = residual_node
        if not isinstance(jacobian, numbers.Number):
            assert jacobian.shape == (self.size, self.size)
        self.jacobian = jacobian

    def forget(self):
        Node.forget(self)
        self.residual = None
        self.jacobian = None

    def dependees(self):
        if getattr(self, 'residual', None):
            yield self.residual
        assert self.prev is None and self.other is None

    def tangent(self, dep_du):
        if not getattr(self, 'residual', None):
            return 0
        (r_du,) = dep_du
        if _int0(r_du):
            return 0
        jac = self.jacobian
        if _is_identity_pattern(jac):
            inv = jac.copy()
            inv.data = 1 / np.array(inv.data)
            return -inv * r_du
        if hasattr(r_du, 'todense'):
            r_du = r_du.todense()
        r_du = np.array(r_du)
        stats['spsolve'] += 1
        lu = spla.factorized(jac.tocsc())
        cols = np.transpose([-lu(c) for c in r_du.T]).reshape(r_du.shape)
        return sp.csr_matrix(cols)


class osolution(oarray):
    """adsolve.py:150-168"""

    def __init__(self, solution, residual, n_newton):
        assert isinstance(solution, oarray) and isinstance(residual, oarray)
        residual._current_state = ResidualNode(residual._current_state)
        oarray.__init__(self, solution._value)
        self._current_state = SolutionNode(self, residual._current_state,
                                           residual.diff(solution))
        self._n_Newton = n_newton
        self._res_norm = np.linalg.norm(residual._value.reshape(residual.size))

    def obliviate(self):
        oarray.obliviate(self)
        del self._n_Newton
        del self._res_norm


adsolution = osolution

# optional hook: called as trace(i_newton, u_value, J_csr, rhs, minus_du) from
# inside Newton; tests use it to record per-step goldens
newton_trace = None


def solve_newton_with_dt(func, u0, args, kargs, dt, max_iter, abs_tol, rel_tol,
                         verbose):
    """adsolve.py:171-205"""
    u = oarray(value(u0).copy())
    for it in range(max_iter):
        if dt == np.inf:
            res = func(u, *args, **kargs)
        else:
            res = (u - u0) / dt + func(u, *args, **kargs)
        rn = np.linalg.norm(res._value.reshape(res.size), np.inf)
        if verbose:
            print('    ', it, rn)
        if it == 0:
            rn0 = rn
        if rn < max(abs_tol, rel_tol * rn0):
            return osolution(u, res, it + 1)
        if not np.isfinite(rn) or rn > rn0 * 1E6:
            break
        J = res.diff(u).tocsr()
        if J.shape[0] > 1:
            stats['spsolve'] += 1
            step = spla.spsolve(J, np.ravel(res._value), use_umfpack=False)
        else:
            step = res._value / J.toarray()[0, 0]
        if newton_trace is not None:
            newton_trace(it, u._value.copy(), J, np.ravel(res._value).copy(), step)
        u._value -= step.reshape(u.shape)
        u = oarray(u._value)
    u = oarray(value(u0).copy())
    return osolution(u, func(u, *args, **kargs), np.inf)


def psuedo_time_continuation(func, u0, args, kargs, dt_min, dt_max, dt_ratio,
                             max_iter, abs_tol, rel_tol, verbose):
    """adsolve.py:207-223"""
    dt = dt_min
    while np.isfinite(dt):
        if np.abs(dt) > np.abs(dt_max):
            dt = np.inf
        if verbose:
            print('continuation step with dt = ', dt)
        u = solve_newton_with_dt(func, u0, args, kargs, dt, max_iter, abs_tol,
                                 rel_tol, verbose)
        if not np.isfinite(u._n_Newton):
            return dt, None
        dt *= dt_ratio
        u0 = u
    if np.isfinite(u._n_Newton):
        return np.inf, u


def solve(func, u0, args=(), kargs={}, max_iter=10, abs_tol=1E-6, rel_tol=1E-6,
          verbose=True):
    """adsolve.py:225-264"""
    u = solve_newton_with_dt(func, u0, args, kargs, np.inf, max_iter, abs_tol,
                             rel_tol, verbose)
    if np.isfinite(u._n_Newton):
        return u
    if verbose:
        print('Newton failed to converge, starting psuedo time continuation')
    dt_min, dt_max, dt_ratio = 1, 4, 2
    u = None
    while u is None:
        dp, u = psuedo_time_continuation(func, u0, args, kargs, dt_min, dt_max,
                                         dt_ratio, max_iter, abs_tol, rel_tol,
                                         verbose)
        if u is not None:
            return u
        dm, u = psuedo_time_continuation(func, u0, args, kargs, -dt_min,
                                         -dt_max, dt_ratio, max_iter, abs_tol,
                                         rel_tol, verbose)
        if u is not None:
            return u
        if np.abs(dp) == dt_min and np.abs(dm) == dt_min:
            dt_min /= 2
        elif np.isinf(dp) or np.isinf(dm):
            dt_max *= 2
        else:
            dt_min /= dt_ratio
            dt_ratio = dt_ratio ** 0.5


# --------------------------------------------------------------------------- #
# comparison helpers used by the parity tests
# --------------------------------------------------------------------------- #
def canonical(m):
    """CSR with duplicates summed and column indices sorted (int64 indices)."""
    m = sp.csr_matrix(m, copy=True)
    m.sum_duplicates()
    m.sort_indices()
    return m


def same_pattern(a, b):
    a, b = canonical(a), canonical(b)
    return (a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and
            np.array_equal(a.indices, b.indices))


def max_rel_diff(a, b):
    """max |a-b| / max |b| over the union pattern."""
    d = abs(sp.csr_matrix(a) - sp.csr_matrix(b))
    scale = abs(sp.csr_matrix(b)).max() if b.nnz else 1.0
    return (d.max() if d.nnz else 0.0) / (scale if scale else 1.0)
