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
