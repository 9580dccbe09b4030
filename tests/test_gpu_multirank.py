"""Multi-rank runs of the distributed layers UNDER PYTEST (VERDICT r1 item 1c / ADVICE):
torchrun is spawned with one rank per visible GPU (2 at least, 8 at most); the test is
skipped on a box with a single device.

  * tools/admpi_check.py  -- the reference's own distributed tests (admpi.py:248-327:
    testSendRecv, testPoisson1DResidual) and its 2-D Poisson demo (admpisolve.py:462-533)
    through COMM_WORLD.Send/Recv, diff_mpi, MpiJacobian, _lgmres, _dot, _norm2;
  * tools/shard_check.py  -- the row-sharded Newton linear solve against the single-GPU
    solve on the same cavity Jacobian, plus bit-identical replicated operators across ranks
    (the design relies on every rank building the same hierarchy).
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nranks():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs (found %d)' % n)
    return 8 if n >= 8 else (4 if n >= 4 else 2)


def _torchrun(script, n, *args, timeout=900):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                          '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
                          '--master-port', str(port), os.path.join(ROOT, 'tools', script), *args],
                         capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    log = out.stdout + '\n' + out.stderr[-3000:]
    logdir = os.environ.get('NPB_TEST_LOG_DIR')           # evidence for profiles/
    if logdir:
        with open(os.path.join(logdir, 'multirank_%s_%dgpu.txt' % (script.replace('.py', ''), n)), 'w') as f:
            f.write(' '.join([script, *args]) + '\n' + out.stdout)
    assert out.returncode == 0, log
    return out.stdout


def test_admpi_reference_tests_multirank():
    n = _nranks()
    out = _torchrun('admpi_check.py', n)
    assert out.count('testSendRecv ok') == n
    assert out.count('testPoisson1DResidual ok') == n
    assert out.count('MpiJacobian + _lgmres ok') == n


def test_sharded_solve_matches_single_gpu_multirank():
    n = _nranks()
    out = _torchrun('shard_check.py', n, '--grid', '384', '--reps', '1', '--strict')
    assert 'sharded x%d' % n in out
    assert 'replicated operators identical on all ranks' in out
    assert 'STRICT OK' in out
