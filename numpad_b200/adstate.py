"""Tape of intermediate states and the chain-rule sweeps over it.

Public surface and semantics follow the reference's adstate.py (class and
attribute names are part of numpad's extension protocol: admpi/adsolve
subclass IntermediateState; advisual/adgarbagecollect read .prev .other
.host() .op_name ._state_id .size), but every Jacobian that flows through the
sweeps lives in HBM: blocks are numbers, `dia_jac`, `sel_jac`, `csr_jac`
(all device resident) and accumulated derivatives are `DevEye` / `DevCsr`;
`_multiply_ops` / `_add_ops` dispatch to CUDA kernels through the C-ABI.

Provenance: `IntermediateState` (fields, method names, generator order of froms()/tos(),
the asserts) and the order of the two sweeps deliberately restate the reference's adstate.py
(qiqi/numpad, GPLv3 -- see the reference's LICENSE): they ARE the extension protocol that
admpi / adsolve and user subclasses program against.  The block algebra, the
device kernels and the early freeing of intermediates are new.

reference: adstate.py:34-229 (tape + sweeps), :233-297 (blocks, add/mul).
"""
import numbers
import weakref

from ._dev import (dia_jac, csr_jac, sel_jac, tocsr, DevEye, DevCsr,  # noqa: F401
                   _add_ops, _multiply_ops, _is0)

_g_state_count = 0


def InitialState(host):
    """A state that depends on nothing (reference adstate.py:36-40)."""
    return IntermediateState(host, None, None, None)


class IntermediateState:
    """One state in the history of an array.

    Unary operation:  `multiplier` is d(self)/d(prev_state); other_state None.
    Binary operation: d(self)/d(prev_state) is the identity and `multiplier`
    is d(self)/d(other_state).  States order (and hash) by creation id, which
    is a topological order of the tape.  (reference adstate.py:42-97)
    """

    def __init__(self, host, prev_state, multiplier, other_state, op_name=''):
        global _g_state_count
        self._state_id = _g_state_count
        _g_state_count += 1
        self.host = weakref.ref(host)
        self.size = host.size
        self.op_name = op_name
        self._to_refs = []
        self.prev = prev_state
        self.other = None
        if prev_state is not None:
            assert isinstance(prev_state, IntermediateState)
            prev_state.next = weakref.ref(self)
        if multiplier is None:
            assert prev_state is None and other_state is None
            return
        self.multiplier = multiplier
        is_number = isinstance(multiplier, numbers.Number)
        if other_state is None:
            assert is_number or multiplier.shape == (self.size, self.size)
        else:
            assert isinstance(other_state, IntermediateState)
            assert is_number or multiplier.shape == (self.size, other_state.size)
            other_state._to_refs.append(weakref.ref(self))
            self.other = other_state

    def __hash__(self):
        return self._state_id

    def __lt__(self, other):
        return self._state_id < other._state_id

    def __eq__(self, other):
        return self._state_id == other._state_id

    def tos(self):
        """live states that immediately depend on this one (adstate.py:99-107)"""
        nxt = getattr(self, 'next', None)
        if nxt is not None and nxt() is not None:
            yield nxt()
        for ref in self._to_refs:
            s = ref()
            if s is not None:
                yield s

    def froms(self):
        """states this one immediately depends on (adstate.py:109-116)"""
        if self.prev:
            yield self.prev
        if self.other:
            yield self.other

    def obliviate(self):
        """cut this state off its dependees and free its block (adstate.py:118-129)"""
        for s in self.froms():
            s._to_refs = [r for r in s._to_refs if r() is not self]
        self.prev = None
        self.other = None
        if hasattr(self, 'multiplier') and not isinstance(self.multiplier, numbers.Number):
            del self.multiplier
        elif hasattr(self, 'multiplier') and self.multiplier:
            del self.multiplier

    def next_state(self, multiplier, other_state=None, op_name=''):
        return IntermediateState(self.host(), self, multiplier, other_state, op_name)

    # ---- local chain-rule steps (adstate.py:142-177) ---- #
    def diff_tangent(self, dependees_diff_u):
        if self.prev is None:
            return 0
        if self.other is None:
            prev_diff_u, = dependees_diff_u
            return _multiply_ops(self.multiplier, prev_diff_u)
        prev_diff_u, other_diff_u = dependees_diff_u
        return _add_ops(prev_diff_u, _multiply_ops(self.multiplier, other_diff_u))

    def diff_adjoint(self, f_diff_dependers):
        f_diff_self = 0
        for state, f_diff_state in zip(self.tos(), f_diff_dependers):
            if state.other is self:
                state_diff_self = state.multiplier
            elif state.prev is self and state.other:
                state_diff_self = 1
            else:
                assert state.prev is self
                state_diff_self = state.multiplier
            f_diff_self = _add_ops(f_diff_self,
                                   _multiply_ops(f_diff_state, state_diff_self))
        return f_diff_self


# ---- sweeps over the tape (adstate.py:181-229) ---- #
def _reach(start, neighbours):
    seen = {}
    todo = [start]
    while todo:
        s = todo.pop(0)
        if s not in seen:
            seen[s] = 0
            todo.extend(neighbours(s))
    return seen


def diff_tangent(f, u):
    """d f / d u accumulated forward from u.  Returns 0, a DevEye or a DevCsr."""
    diff_u = _reach(f, lambda s: s.froms())
    # how many dependers inside the swept set still need each derivative
    pending = {}
    for s in diff_u:
        for d in s.froms():
            pending[d] = pending.get(d, 0) + 1
    for state in sorted(diff_u):
        if state is u:
            diff_u[state] = DevEye(u.size)
        else:
            dependees = list(state.froms())
            diff_u[state] = state.diff_tangent(diff_u[s] for s in dependees)
            for d in dependees:          # release HBM as soon as nothing needs it
                pending[d] -= 1
                if pending[d] == 0 and d is not f:
                    diff_u[d] = 0
    return diff_u[f]


def diff_adjoint(f, u):
    """d f / d u accumulated backward from f.  Returns 0, a DevEye or a DevCsr."""
    f_diff = _reach(u, lambda s: s.tos())
    pending = {}
    for s in f_diff:
        for d in s.tos():
            if d in f_diff:
                pending[d] = pending.get(d, 0) + 1
    for state in sorted(f_diff, reverse=True):
        if state is f:
            f_diff[state] = DevEye(f.size)
        else:
            dependers = list(state.tos())
            f_diff[state] = state.diff_adjoint(f_diff[s] for s in dependers)
            for d in dependers:
                pending[d] -= 1
                if pending[d] == 0 and d is not u:
                    f_diff[d] = 0
    return f_diff[u]
