"""ctypes binding of libbsde_b200.so (C ABI declared in include/bsde_b200.h).

The library is the product: there is no CPU path behind it.  Loading fails loudly when the shared object has not
been built (`python bsde-tensor-networks_b200/build.py`), and creating a context fails loudly without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSDE_B200_LIB selects an alternative build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("BSDE_B200_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "libbsde_b200.so"))

BASIS_PHI, BASIS_POLY, BASIS_LEGENDRE = 0, 1, 2
SOLVER_IDS = {"lstsq": 0, "lu": 1, "qr": 2, "solve": 3}
MODEL_BLACKSCHOLES, MODEL_HJB = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_i64 = C.c_int64

# name -> argtypes (restype is int for all but bsde_last_error); mirrors include/bsde_b200.h one to one
SIGNATURES = {
    "bsde_version": [],
    "bsde_ctx_create": [C.c_int, C.POINTER(_vp)],
    "bsde_ctx_destroy": [_vp],
    "bsde_ctx_set_stream": [_vp, _vp],
    "bsde_ctx_sync": [_vp],
    "bsde_ctx_launch_count": [_vp, C.POINTER(_i64)],
    "bsde_ctx_set_option": [_vp, C.c_char_p, _i64],
    "bsde_ctx_get_stat": [_vp, C.c_char_p, C.POINTER(C.c_double)],
    "bsde_comm_unique_id": [_vp],
    "bsde_comm_init": [_vp, _vp, C.c_int, C.c_int],
    "bsde_comm_p2p_export": [_vp, _vp],
    "bsde_comm_p2p_open": [_vp, _vp, C.c_int],
    "bsde_malloc": [_vp, _i64, C.POINTER(_vp)],
    "bsde_free": [_vp, _vp],
    "bsde_memcpy_h2d": [_vp, _vp, _vp, _i64],
    "bsde_memcpy_d2h": [_vp, _vp, _vp, _i64],
    "bsde_transpose": [_vp, _vp, _i64, _i64, _vp],
    "bsde_fit_als": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp],
    "bsde_fit_als_host": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp],
    "bsde_fit_als_host_assets": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp],
    "bsde_fit_salsa": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_double, C.c_double,
                       C.c_double, C.c_int, C.c_int, C.c_int, _vp, _i64, _vp],
    "bsde_tt_eval": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp],
    "bsde_tt_grad": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "bsde_tt_hessian": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bsde_paths_simulate": [_vp, C.c_int, _vp, C.c_int, _i64, C.c_int, C.c_double, _vp, C.c_uint64, _i64, C.c_int, _vp, _vp],
    "bsde_terminal": [_vp, C.c_int, _vp, C.c_int, _i64, _vp, _vp],
    "bsde_target": [_vp, C.c_int, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_int, _vp],
    "bsde_mean": [_vp, _vp, _i64, _vp],
    "bsde_normal_eq": [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp],
    "bsde_solve": [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp],
    "bsde_orthonormalize": [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int],
    "bsde_salsa_move": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp],
    "bsde_orthonormalize_shape": [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int],
    "bsde_basis_eval": [_vp, C.c_int, C.c_int, _i64, _vp, _vp, C.c_int, _vp],
}

_lib = None


class BackendError(RuntimeError):
    """A call into libbsde_b200.so failed (message from bsde_last_error)."""


def load():
    """Load the shared library once; raises if it is missing — there is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA backend is not built. Run `python bsde-tensor-networks_b200/build.py` "
            "(needs nvcc; cross-compiles for sm_100a without a GPU). This package has no CPU implementation."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.bsde_last_error.restype = C.c_char_p
    lib.bsde_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().bsde_last_error()
        raise BackendError(f"libbsde_b200 error {rc}: {msg.decode() if msg else '?'}")
