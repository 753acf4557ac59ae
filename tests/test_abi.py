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
