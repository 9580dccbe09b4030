"""record an assembly plan at the fixed state, replay it at the zero state (deviating replay)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import numpad_b200 as npd
from numpad_b200 import _dev
from numpad_b200.workloads.cavity import CavityProblem
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
P = CavityProblem(npd, N)
for state in (P.fixed_state(), np.zeros(P.n), P.fixed_state(seed=3)):
    u = npd.array(state.copy())
    res = P.residual(u, npd.array(state.copy()), 1.0, 0)
    J = res.diff(u)
    torch.cuda.synchronize()
    print('nnz', J.nnz, _dev.plan_stats, flush=True)
