"""The C ABI: every symbol declared in include/estdepth_b200.h is exported by the library and bound by the ctypes
layer; argument validation returns error codes (no compute, no GPU needed)."""
import ctypes
import os
import re

import pytest

from estdepth_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "estdepth_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"ESTD_API[^;(]*?\b(estd_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 18
    assert sorted(_lib.SIGNATURES.keys()) == names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n


def test_library_loads_and_reports_version():
    lib = _lib.get()
    assert lib.estd_version() == 100
    assert isinstance(_lib.launch_count(), int)


def test_bad_arguments_return_error_codes_without_launching():
    lib = _lib.get()
    n0 = _lib.launch_count()
    assert lib.estd_premix(None, None, None, None, 32, 32, 8, 8, None) == -1
    assert b"null pointer" in lib.estd_last_error()
    assert lib.estd_warp_cost(None, None, None, None, None, 32, 8, 8, 8, 0, None) == -1
    assert lib.estd_head_softargmin(None, None, None, None, None, None, None, None, None, 8, 8, 8, 4, None) == -1
    assert lib.estd_conv3d(None, None) == -1
    d = _lib.ConvDesc()
    d.in0_chunks, d.in1_chunks, d.cout_pad, d.D, d.H, d.W = 5, 0, 24, 8, 8, 8
    assert lib.estd_conv3d_num_ctas(ctypes.byref(d)) == -3           # ESTD_EUNSUPPORTED: no such specialisation
    assert b"no kernel" in lib.estd_last_error()
    assert _lib.launch_count() == n0
    with pytest.raises(RuntimeError, match="no kernel"):
        _lib.check(-3, "estd_conv3d_num_ctas")


def test_conv_grid_is_persistent_and_bounded():
    lib = _lib.get()
    d = _lib.ConvDesc()
    d.in0_chunks, d.in1_chunks, d.cout_pad, d.D, d.H, d.W = 8, 0, 32, 64, 120, 160
    n = lib.estd_conv3d_num_ctas(ctypes.byref(d))
    assert 1 <= n <= 2400 and n <= 160            # one CTA per SM (148 on B200; 148 assumed without a device)
    d.D, d.H, d.W = 2, 8, 32
    assert lib.estd_conv3d_num_ctas(ctypes.byref(d)) == 1
