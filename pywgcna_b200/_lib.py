"""ctypes binding of ``libb200wgcna.so`` (C ABI declared in ``include/b200wgcna.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C pywgcna_b200/csrc``.  There is
no CPU fallback: a missing library raises at first use, and every compute call raises
``B200WError`` when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200wgcna.so")

NET_TYPES = ["unsigned", "signed", "signed hybrid"]  # wgcna.py:53-54
TOM_TYPES = ["unsigned", "signed"]  # wgcna.py:55
TOM_DENOMS = ["min", "mean"]  # wgcna.py:56
OUT_TOM, OUT_DISS, OUT_DISS_R8, OUT_DISS_COND, OUT_DISS_R8_COND = 0, 1, 2, 3, 4
PREC_AUTO, PREC_F64, PREC_I8, PREC_I8_FAST = 0, 1, 2, 3
I8_SLICES = 5

E_INVALID, E_NO_DEVICE, E_CUDA, E_NOMEM, E_UNSUPPORTED = -1, -2, -3, -4, -5


class B200WError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200wgcna error {code}: {msg}")
        self.code = code


_dp = C.c_void_p  # device or host pointer, passed as integer address
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double

# name -> (restype, argtypes); every symbol include/b200wgcna.h declares
SIGNATURES = {
    "b200w_version": (_int, []),
    "b200w_last_error": (C.c_char_p, []),
    "b200w_device_count": (_int, []),
    "b200w_connectivity_sweep": (_int, [_dp, _i64, _i64, _dp, _int, _int, _dp, _dp, _int]),
    "b200w_sft_exp_buckets": (_int, []),
    "b200w_soft_threshold_stats": (_int, [_dp, _i64, _i64, _dp, _int, _int, _int, _dp, _dp, _int]),
    "b200w_soft_threshold_stats_gm": (_int, [_dp, _i64, _i64, _dp, _int, _int, _int, _dp, _dp, _int]),
    "b200w_adjacency_gm": (_int, [_dp, _i64, _i64, _int, _dbl, _i64, _i64, _dp, _int]),
    "b200w_dev_standardize_gm": (_int, [_dp, _i64, _i64, _dp, _dp, _int, _dp]),
    "b200w_dev_sft_stats": (_int, [_dp, _int, _i64, _i64, _int, _dp, _dp, _int, _dp]),
    "b200w_adjacency": (_int, [_dp, _i64, _i64, _int, _dbl, _i64, _i64, _dp, _int]),
    "b200w_check_adjacency": (_int, [_dp, _i64, _dp, _int]),
    "b200w_tom": (_int, [_dp, _i64, _int, _int, _int, _int, _i64, _i64, _dp, _int, _dp, _int]),
    "b200w_network": (_int, [_dp, _i64, _i64, _int, _dbl, _int, _int, _int, _int, _dp, _dp, _int]),
    "b200w_cross_correlation": (_int, [_dp, _i64, _i64, _dp, _i64, _dp, _int]),
    "b200w_dev_cross_correlation": (_int, [_dp, _dp, _i64, _dp, _dp, _i64, _i64, _dp, _int, _dp]),
    "b200w_linkage_average": (_int, [_dp, _i64, _dp, _int]),
    "b200w_dev_linkage_average": (_int, [_dp, _i64, _dp, _int, _dp]),
    "b200w_dev_condensed_to_square": (_int, [_dp, _i64, _dp, _int, _dp]),
    "b200w_linkage_finish": (_int, [_dp, _i64]),
    "b200w_intramodular": (_int, [_dp, _i64, _dp, _int, _dp, _dp, _dp, _int]),
    "b200w_host_alloc": (C.c_void_p, [_i64]),
    "b200w_host_free": (None, [C.c_void_p]),
    "b200w_host_trim": (None, []),
    "b200w_dev_trim": (None, []),
    "b200w_set_correlation_precision": (_int, [_int]),
    "b200w_padded_k": (_i64, [_i64]),
    "b200w_dev_standardize": (_int, [_dp, _i64, _i64, _dp, _dp, _int, _dp]),
    "b200w_dev_sweep": (_int, [_dp, _dp, _i64, _i64, _dp, _int, _int, _i64, _i64, _dp, _int, _dp]),
    "b200w_dev_sweep_similarity": (_int, [_dp, _dp, _i64, _i64, _dp, _int, _int, _i64, _i64, _dp, _dp, _int, _dp]),
    "b200w_dev_sweep_dist": (_int, [_dp, _dp, _i64, _i64, _dp, _int, _int, _int, _int, _i64, C.POINTER(C.c_void_p), _dp, _int, _dp]),
    "b200w_sweep_dist_tiles": (_i64, [_i64, _int, _int, _i64, _dp, _i64]),
    "b200w_sweep_dist_regions": (_int, [_i64, _int, _int, _i64, _dp]),
    "b200w_dev_adjacency_from_similarity": (_int, [_dp, _i64, _i64, _dbl, _int, _dp]),
    "b200w_dev_adjacency_from_similarity_dp": (_int, [_dp, _i64, _i64, _dp, _int, _dp]),
    "b200w_dev_adjacency_dp": (_int, [_dp, _dp, _i64, _i64, _int, _dp, _i64, _i64, _dp, _int, _dp]),
    "b200w_dev_sft_select": (_int, [_dp, _int, _i64, _i64, _int, _dp, _dbl, _dbl, _dp, _dp, _dp, _dp, _int, _dp]),
    "b200w_dev_adjacency": (_int, [_dp, _dp, _i64, _i64, _int, _dbl, _i64, _i64, _dp, _int, _dp]),
    "b200w_dev_check_adjacency": (_int, [_dp, _i64, _dp, _int, _dp]),
    "b200w_padded_n": (_i64, [_i64]),
    "b200w_tom_operand_bytes": (_i64, [_i64, _int]),
    "b200w_dev_tom_prepare": (_int, [_dp, _i64, _i64, _i64, _int, _dp, _i64, _dp, _dp, _int, _dp]),
    "b200w_dev_tom_compute": (_int, [_dp, _dp, _i64, _dp, _dp, _i64, _i64, _i64, _int, _int, _int, _int,
                                     _dp, _int, _dp]),
    "b200w_dev_tom_compute_to_host": (_int, [_dp, _dp, _i64, _dp, _dp, _i64, _int, _int, _int, _int, _dp, _dp, _int, _dp]),
    "b200w_dev_tom_compute_dist": (_int, [_dp, _dp, _i64, _dp, _dp, _i64, _int, _int, _i64, _int, _int, _int,
                                          C.POINTER(C.c_void_p), _int, _dp]),
    "b200w_tom_shared_scale_offset": (_i64, [_i64, _i64]),
    "b200w_tom_shared_bytes": (_i64, [_i64, _i64]),
    "b200w_dev_tom_pull_planes": (_int, [_dp, C.POINTER(C.c_void_p), _int, _int, _i64, _i64, _i64, _dp, C.c_uint32,
                                         _int, _dp]),
    "b200w_dev_tom_compute_dist_gated": (_int, [_dp, _dp, _i64, _dp, _dp, _i64, _int, _int, _i64, _int, _int, _int,
                                                C.POINTER(C.c_void_p), _dp, C.c_uint32, _int, _dp]),
    "b200w_dev_malloc": (_int, [_i64, _int, C.POINTER(C.c_void_p)]),
    "b200w_dev_free": (_int, [_dp, _int]),
    "b200w_ipc_export": (_int, [_dp, C.c_char_p]),
    "b200w_ipc_open": (_int, [C.c_char_p, _int, C.POINTER(C.c_void_p)]),
    "b200w_ipc_close": (_int, [_dp, _int]),
    "b200w_dev_download": (_int, [_dp, _dp, _i64, _int, _dp]),
    "b200w_dev_upload": (_int, [_dp, _dp, _i64, _int, _dp]),
    "b200w_launch_count": (_i64, []),
    "b200w_last_kernel_ms": (_dbl, []),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pywgcna_b200/csrc`.  pywgcna_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        raise B200WError(code, load().b200w_last_error().decode("utf-8", "replace"))


def hptr(a: np.ndarray) -> int:
    """Address of a C-contiguous numpy array."""
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def as_f64_matrix(x) -> np.ndarray:
    """C-contiguous float64 2-D view/copy of ``x`` (DataFrame or array-like)."""
    if hasattr(x, "to_numpy"):
        x = x.to_numpy()
    a = np.ascontiguousarray(x, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("expected a 2-D matrix")
    return a


def as_f64_expression(x):
    """``(array, gene_major)`` for an expression matrix (samples x genes) WITHOUT copying when it is float64 and
    contiguous in either order: a C-ordered m x n array goes as it is (gene_major False); an F-ordered one -- what
    ``DataFrame.to_numpy()`` returns when pandas keeps its block gene-major -- goes as its transpose, the C-ordered
    n x m array the ``*_gm`` entry points take (gene_major True).  Anything else is converted like as_f64_matrix."""
    if hasattr(x, "to_numpy"):
        x = x.to_numpy()
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError("expected a 2-D matrix")
    if a.dtype == np.float64 and not a.flags["C_CONTIGUOUS"] and a.flags["F_CONTIGUOUS"]:
        return a.T, True
    return np.ascontiguousarray(a, dtype=np.float64), False


def result_empty(shape, dtype=np.float64) -> np.ndarray:
    """Uninitialised C-contiguous array for a (large) result.  Arrays of 32 MB and more live in a
    page-locked buffer of the library's recycling pool (DMA instead of staged copies, no first-touch
    page faults); the buffer goes back to the pool when the last view of the array dies."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes < (32 << 20) or os.environ.get("B200WGCNA_NO_PINNED_POOL"):
        return np.empty(shape, dtype=dtype)
    lib = load()
    ptr = lib.b200w_host_alloc(nbytes)
    if not ptr:
        return np.empty(shape, dtype=dtype)
    buf = (C.c_byte * nbytes).from_address(ptr)
    weakref.finalize(buf, lib.b200w_host_free, ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def device_count() -> int:
    return int(load().b200w_device_count())


def default_device() -> int:
    """Device the host entry points use: $B200WGCNA_DEVICE, else LOCAL_RANK (torchrun), else 0."""
    for var in ("B200WGCNA_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v is not None and v != "":
            return int(v)
    return 0
