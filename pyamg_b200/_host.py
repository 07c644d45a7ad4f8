"""ctypes binding of libpyamg_b200_host.so (host-side setup helpers, csrc/host_setup.cpp)."""
import ctypes

import numpy as np

from . import build as _build

_lib = None
_I = ctypes.POINTER(ctypes.c_int32)
_D = ctypes.POINTER(ctypes.c_double)
_U8 = ctypes.POINTER(ctypes.c_uint8)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build_host_library())
        i32, f64 = ctypes.c_int32, ctypes.c_double
        L.amgb_setup_classical_strength.restype = ctypes.c_int64
        L.amgb_setup_classical_strength.argtypes = [i32, _I, _I, _D, f64, _I, _I, _D, _I]
        L.amgb_setup_rs_splitting.restype = None
        L.amgb_setup_rs_splitting.argtypes = [i32, _I, _I, _I, _I, _I]
        L.amgb_setup_classical_interp_count.restype = None
        L.amgb_setup_classical_interp_count.argtypes = [i32, _I, _I, _I, _U8, _I]
        L.amgb_setup_classical_interp_fill.restype = None
        L.amgb_setup_classical_interp_fill.argtypes = [i32, _I, _I, _D, _I, _I, _D, _U8, _I, _I, _I, _D]
        L.amgb_setup_greedy_coloring.restype = i32
        L.amgb_setup_greedy_coloring.argtypes = [i32, _I, _I, _I]
        L.amgb_setup_greedy_coloring_ordered.restype = i32
        L.amgb_setup_greedy_coloring_ordered.argtypes = [i32, _I, _I, i32, _I]
        L.amgb_setup_symmetric_strength.restype = ctypes.c_int64
        L.amgb_setup_symmetric_strength.argtypes = [i32, _I, _I, _D, f64, _I, _I]
        L.amgb_setup_standard_aggregation.restype = i32
        L.amgb_setup_standard_aggregation.argtypes = [i32, _I, _I, _I, _I]
        L.amgb_setup_gauss_seidel.restype = None
        L.amgb_setup_gauss_seidel.argtypes = [i32, _I, _I, _D, _D, _D, i32, i32]
        L.amgb_setup_block_gauss_seidel.restype = None
        L.amgb_setup_block_gauss_seidel.argtypes = [i32, i32, _I, _I, _D, _D, _D, _D, i32, i32]
        L.amgb_setup_coloring_is_valid.restype = i32
        L.amgb_setup_coloring_is_valid.argtypes = [i32, _I, _I, _I]
        L.amgb_setup_arnoldi_round.restype = i32
        L.amgb_setup_arnoldi_round.argtypes = [i32, _I, _I, _D, _D, _D, i32, f64, _D]
        _lib = L
    return _lib


def ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_I)


def dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_D)


def u8p(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(_U8)
