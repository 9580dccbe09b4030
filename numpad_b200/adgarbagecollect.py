"""Tape garbage collection (reference adgarbagecollect.py:27-66): cut the edges to states
whose host arrays are gone and on which nothing alive depends, so that their device-resident
Jacobian blocks are freed."""
import gc


def _edges(state):
    yield 'other', state.other
    yield 'prev', state.prev
    yield 'residual', getattr(state, 'residual', None)


def collect(state):
    gc.collect()
    memo = {}

    def dead(s):
        """True when s and everything below it can go (its host array no longer exists)"""
        if id(s) in memo:
            return memo[id(s)]
        memo[id(s)] = ok = s.host() is None
        for name, child in _edges(s):
            if child:
                if dead(child):
                    setattr(s, name, None)
                else:
                    ok = False
        memo[id(s)] = ok
        return ok
    dead(state)
