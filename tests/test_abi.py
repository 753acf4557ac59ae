"""The C-ABI library loads and exports every symbol include/b200wgcna.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from pywgcna_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200wgcna.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200w_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200wgcna.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes prototype in pywgcna_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(syms)


def test_version_and_pure_functions():
    lib = _lib.load()
    assert lib.b200w_version() == 100
    assert lib.b200w_padded_k(200) == 208 and lib.b200w_padded_k(16) == 16
    assert lib.b200w_padded_n(20000) == 20096 and lib.b200w_padded_n(128) == 128
    assert lib.b200w_tom_operand_bytes(256, _lib.PREC_F64) == 256 * 256 * 8
    assert lib.b200w_tom_operand_bytes(256, _lib.PREC_I8) == _lib.I8_SLICES * 256 * 256


def test_fails_loudly_without_device():
    """No CPU fallback: on a box without a B200 every compute entry point reports NO_DEVICE."""
    if _lib.device_count() > 0:
        pytest.skip("a device is present")
    import numpy as np

    import pywgcna_b200 as bw

    with pytest.raises(_lib.B200WError) as ei:
        bw.adjacency(np.random.default_rng(0).random((8, 12)))
    assert ei.value.code == _lib.E_NO_DEVICE


def test_header_is_valid_c(tmp_path):
    """include/b200wgcna.h is a plain C header (extern "C" ABI): it must compile as C99 and as C++."""
    import subprocess

    src = tmp_path / "t.c"
    src.write_text('#include "b200wgcna.h"\nint main(void) { return B200W_VERSION == 100 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True)


def test_argument_validation_returns_codes_not_crashes():
    """Bad arguments are rejected with B200W_E_INVALID and a message before any device work."""
    import numpy as np

    lib = _lib.load()
    x = np.zeros((4, 4))
    out = np.zeros((4, 4))
    rc = lib.b200w_adjacency(x.ctypes.data, 4, 4, 7, 6.0, 0, 4, out.ctypes.data, 0)  # bad network type or no device
    assert rc in (_lib.E_INVALID, _lib.E_NO_DEVICE)
    assert lib.b200w_adjacency(None, 4, 4, 0, 6.0, 0, 4, out.ctypes.data, 0) == _lib.E_INVALID
    assert lib.b200w_adjacency(x.ctypes.data, 4, 4, 0, 6.0, 2, 4, out.ctypes.data, 0) == _lib.E_INVALID  # rows past n
    assert lib.b200w_tom(x.ctypes.data, 0, 0, 0, 0, 0, 0, 0, out.ctypes.data, 0, None, 0) == _lib.E_INVALID
    assert lib.b200w_connectivity_sweep(x.ctypes.data, 4, 4, None, 3, 0, out.ctypes.data, None, 0) == _lib.E_INVALID
    assert b"bad args" in lib.b200w_last_error()
    stats = np.zeros(4)
    assert lib.b200w_check_adjacency(x.ctypes.data, -1, stats.ctypes.data, 0) == _lib.E_INVALID
    assert lib.b200w_dev_tom_prepare(None, 4, 0, 4, 2, None, 128, None, None, 0, None) == _lib.E_INVALID
