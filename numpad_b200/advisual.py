"""Graphviz image of a tape (reference advisual.py:31-69): one node per state `S<id>_<size>`,
one edge per dependency labelled with the operation; the edge from a solution to the residual
it solves is red."""
from .adgarbagecollect import collect


def dot(f):
    collect(f._current_state)
    lines, seen = [], set()
    stack = [f._current_state]
    name = lambda s: 'S{0}_{1}'.format(s._state_id, s.size)
    while stack:
        s = stack.pop()
        if id(s) in seen:
            continue
        seen.add(id(s))
        for child, attr in ((s.other, ''), (s.prev, ''), (getattr(s, 'residual', None), 'color=red')):
            if child:
                lines.append('{0} -> {1} [label="{2}", {3}];\n'.format(name(s), name(child), s.op_name, attr))
                stack.append(child)
    return 'digraph G{\n' + ''.join(lines) + '}\n'
