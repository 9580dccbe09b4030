"""CPU-side checks of the C-ABI: the shared library loads, exports every
symbol include/numpad_b200.h declares, and the package refuses to compute
without a device (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

from numpad_b200 import _lib


def test_header_symbols_exported():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, missing


def test_exports_match_header():
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l and 'npb_' in l}
    exported = {e for e in exported if e.startswith('npb_')}
    declared = set(_lib.parse_header())
    assert declared <= exported
    # nothing public is left undeclared
    assert exported - declared == set(), exported - declared


def test_version_and_error_string():
    L = _lib.lib()
    assert L.cdll.npb_version() >= 100
    assert isinstance(L.last_error(), str)


def test_sm100a_sass_present():
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in out.stdout


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('device present')
    import numpad_b200 as npd
    with pytest.raises(RuntimeError):
        npd.zeros(4)


def test_rcm_host():
    """npb_rcm_host is host code: a valid permutation whose bandwidth matches
    scipy's reverse Cuthill-McKee on a 2-D grid operator"""
    import numpy as np
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    n1 = 30
    T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n1, n1))
    A = sp.csr_matrix(sp.kron(sp.identity(n1), T) + sp.kron(T, sp.identity(n1)))
    rng = np.random.RandomState(0)
    p0 = rng.permutation(n1 * n1)
    A = A[p0][:, p0].tocsr()
    A.sort_indices()
    n = A.shape[0]
    ptr, idx = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    perm, bw = np.empty(n, np.int32), np.zeros(2, np.int32)
    rc = _lib.lib().cdll.npb_rcm_host(n, ptr.ctypes.data, idx.ctypes.data,
                                      perm.ctypes.data, bw.ctypes.data)
    assert rc == 0 and sorted(perm) == list(range(n))
    B = A[perm][:, perm].tocoo()
    assert (B.row - B.col).max() == bw[0] and (B.col - B.row).max() == bw[1]
    q = reverse_cuthill_mckee(A, symmetric_mode=True)
    C = A[q][:, q].tocoo()
    assert bw[0] <= 1.5 * (C.row - C.col).max()


def test_struct_layouts_match_the_header(tmp_path):
    """the ctypes mirrors of the host structs (amg.py, _lib.py) have the size and the field
    offsets the C header gives them: compile a probe against include/numpad_b200.h with gcc"""
    import shutil
    if shutil.which('gcc') is None:
        pytest.skip('gcc unavailable')
    from numpad_b200 import amg, precond
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = [('npb_halo', amg.Halo), ('npb_csr_mat', amg.CsrMat), ('npb_amg_level', amg.AmgLevel),
             ('npb_amg', amg.AmgStruct), ('npb_amg_precond', amg.AmgPrecondStruct),
             ('npb_lsc', amg.LscStruct), ('npb_csr_view', _lib.CsrView),
             ('npb_krylov_ops', _lib.KrylovOps), ('npb_krylov_info', _lib.KrylovInfo),
             ('npb_scales', _lib.Scales), ('npb_block_jacobi', precond._BlockJacobiCtx)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "numpad_b200.h"', 'int main(void) {']
    for name, cls in pairs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, fname, name, fname))
    lines += ['return 0; }']
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.check_call(['gcc', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in pairs:
        assert int(got[name]) == ctypes.sizeof(cls), name
        for fname, _ in cls._fields_:
            assert int(got['%s.%s' % (name, fname)]) == getattr(cls, fname).offset, (name, fname)
