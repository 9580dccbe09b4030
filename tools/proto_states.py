"""Generate convection-dominated test systems for solver prototyping (CPU, oracle):
the reference's driver loop (test/cavity.py:97-114) from rest, every Newton linear
system (J, rhs, du) of the first `--steps` time steps saved to /tmp/proto_states_N.npz.

    python tools/proto_states.py --grid 128 --steps 6
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import numpad_oracle as orc                       # noqa: E402
from numpad_b200.workloads.cavity import CavityProblem        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--grid', type=int, default=128)
ap.add_argument('--steps', type=int, default=6)
ap.add_argument('--h', type=float, default=0.0)
ap.add_argument('--out', default='')
a = ap.parse_args()
P = CavityProblem(orc, a.grid, h=(1.0 / a.h) if a.h else None)
out = {}
cur = {'step': 0}


def trace(it, u, J, rhs, du):
    k = 's%d_i%d' % (cur['step'], it)
    J = sp.csr_matrix(J)
    out[k + '.indptr'], out[k + '.indices'], out[k + '.data'] = J.indptr, J.indices, J.data
    out[k + '.rhs'], out[k + '.du'], out[k + '.u'] = rhs, du, u
    print('   saved', k, 'n', J.shape[0], 'nnz', J.nnz, flush=True)


orc.newton_trace = trace
q = orc.zeros(P.n)
force = orc.zeros(P.n)
t, dt = 0, 1.0
for step in range(a.steps):
    cur['step'] = step
    t0 = time.time()
    print('t =', t, 'dt =', dt, flush=True)
    q = orc.solve(P.residual, q, args=(q, dt, force), rel_tol=0.05, abs_tol=1e-8, verbose=False)
    out['s%d.dt' % step] = dt
    out['s%d.nnewton' % step] = q._n_Newton
    out['s%d.q' % step] = np.ravel(q._value).copy()
    print('  -> %s Newton its, %.1fs, |q| = %.12e' % (q._n_Newton, time.time() - t0,
                                                    np.linalg.norm(q._value)), flush=True)
    if q._n_Newton == 1:
        break
    elif q._n_Newton < 4:
        dt *= 2
    t += dt
    q.obliviate()
np.savez(a.out or '/tmp/proto_states_%d.npz' % a.grid, **out)
