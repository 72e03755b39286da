"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol include/slsb200.h
declares.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

import slc_b200
from slc_b200 import _capi


def header_symbols():
    src = open(os.path.join(ROOT, "include", "slsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sls_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for s in ("sls_create", "sls_destroy", "sls_set_plant", "sls_set_supports", "sls_plan", "sls_reduce", "sls_reduce_of",
              "sls_solve", "sls_solve_device", "sls_last_error"):
        assert s in syms


def test_library_builds_and_exports_every_header_symbol():
    from slc_b200 import build as b
    lib = b.build()
    dll = ctypes.CDLL(lib)
    missing = [s for s in header_symbols() if not hasattr(dll, s)]
    assert not missing, missing
    assert set(header_symbols()) == set(_capi.SIGNATURES), "ctypes binding out of sync with include/slsb200.h"
    dll.sls_abi_version.restype = ctypes.c_int
    assert dll.sls_abi_version() == _capi.ABI_VERSION


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every compute entry point must fail loudly, not fall back."""
    api = _capi.default_api()
    if api.sls_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_capi.SlsError):
        slc_b200.Synthesizer()
    msg = api.sls_last_error(None).decode()
    assert "no usable CUDA device" in msg


def test_sass_is_sm100a_only():
    import subprocess
    lib = os.path.join(ROOT, "systemlevelcontrol.jl_b200", "lib", "libslsb200.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
