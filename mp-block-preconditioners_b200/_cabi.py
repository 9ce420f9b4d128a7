"""ctypes binding of libmpbp.so (include/mpbp.h).  No torch types cross this boundary: only raw
device/host pointers, sizes and plain structs.  Importing this module fails loudly when the CUDA
library has not been built -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpbp.so")

SUB_JACOBI, SUB_MG = 0, 1
SIDE_LEFT, SIDE_RIGHT = 0, 1


class MpbpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmpbp error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("xi", C.c_double), ("eta_n", C.c_double), ("eta_s", C.c_double),
        ("c", C.c_double), ("d_u", C.c_double), ("d_p", C.c_double), ("d_div", C.c_double),
        ("theta_host", C.c_void_p),
        ("rank", C.c_int), ("nranks", C.c_int),
        ("nccl_unique_id", C.c_void_p),
        ("F_kind", C.c_int), ("P_kind", C.c_int),
        ("F_sweeps", C.c_int), ("P_sweeps", C.c_int),
        ("F_cycles", C.c_int), ("P_cycles", C.c_int),
        ("omega", C.c_double),
        ("nu1", C.c_int), ("nu2", C.c_int),
        ("n_coarse", C.c_int),
        ("cheb", C.c_int),
        ("lmin", C.c_double), ("lmax", C.c_double),
        ("project", C.c_int),
        ("operators_only", C.c_int),
        ("dist_min_n", C.c_int),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
    ]


class GmresOpts(C.Structure):
    _fields_ = [
        ("rtol", C.c_double),
        ("restart", C.c_int),
        ("maxiter", C.c_int),
        ("side", C.c_int),
        ("use_precond", C.c_int),
        ("x0_nonzero", C.c_int),
        ("force_iters", C.c_int),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("xk_buf", C.c_void_p),
        ("iter_cb", C.c_void_p),
        ("cb_user", C.c_void_p),
    ]


ITER_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double)

_P = C.c_void_p  # plan handle / device pointer / stream
_D = C.c_void_p

# name -> (restype, argtypes); every symbol include/mpbp.h declares
SIGNATURES = {
    "mpbp_config_default": (C.c_int, [C.POINTER(Config)]),
    "mpbp_plan_workspace_bytes": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_size_t)]),
    "mpbp_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(Config)]),
    "mpbp_plan_destroy": (C.c_int, [_P]),
    "mpbp_last_error_string": (C.c_char_p, []),
    "mpbp_build_id": (C.c_char_p, []),
    "mpbp_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "mpbp_plan_rows_local": (C.c_int, [_P]),
    "mpbp_plan_num_levels": (C.c_int, [_P]),
    "mpbp_plan_launches": (C.c_longlong, [_P]),
    "mpbp_apply_A": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_apply_F": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_apply_G": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_apply_D": (C.c_int, [_P, _D, _D, _D, _P]),
    "mpbp_apply_GtG": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_apply_GtFG": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_jacobi_F": (C.c_int, [_P, _D, _D, C.c_int, C.c_double, _P]),
    "mpbp_jacobi_P": (C.c_int, [_P, _D, _D, C.c_int, C.c_double, _P]),
    "mpbp_vcycle_F": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_vcycle_P": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_solve_F": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_solve_P": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_precond_apply": (C.c_int, [_P, _D, _D, _P]),
    "mpbp_precond_apply_host": (C.c_int, [_P, C.c_void_p, C.c_void_p, _P]),
    "mpbp_precond_bytes": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "mpbp_comm_probe": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
    "mpbp_dot": (C.c_int, [_P, _D, _D, C.c_size_t, C.POINTER(C.c_double), _P]),
    "mpbp_nrm2": (C.c_int, [_P, _D, C.c_size_t, C.POINTER(C.c_double), _P]),
    "mpbp_axpy": (C.c_int, [_P, C.c_double, _D, _D, C.c_size_t, _P]),
    "mpbp_multi_dot": (C.c_int, [_P, _D, C.c_size_t, C.c_int, _D, C.c_size_t, C.POINTER(C.c_double), _P]),
    "mpbp_multi_axpy": (C.c_int, [_P, _D, C.c_size_t, C.c_int, C.POINTER(C.c_double), _D, C.c_size_t, _P]),
    "mpbp_wnorms": (C.c_int, [_P, _D, _D, C.c_size_t, C.c_double, C.POINTER(C.c_double), _P]),
    "mpbp_fill_manufactured": (C.c_int, [_P, _D, _D, C.c_double, _P]),
    "mpbp_gmres_opts_default": (C.c_int, [C.POINTER(GmresOpts)]),
    "mpbp_gmres_workspace_bytes": (C.c_int, [_P, C.POINTER(GmresOpts), C.POINTER(C.c_size_t)]),
    "mpbp_gmres": (C.c_int, [_P, _D, _D, C.POINTER(GmresOpts), C.POINTER(C.c_double), C.c_int,
                             C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "mpbp_gmres_last_hessenberg": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]),
    "mpbp_gmres_host": (C.c_int, [_P, C.c_void_p, C.c_void_p, C.POINTER(GmresOpts), C.POINTER(C.c_double), C.c_int,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
}

_lib = None


def load():
    """Load libmpbp.so (built by __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'`). This package has no CPU fallback.")
    from . import _build
    if _build.binary_id(LIB_PATH) != _build.source_id():
        # stale binary (sources edited after the last build): rebuild when a compiler is here, never run it silently
        try:
            _build.build(force=True)
        except Exception as exc:
            raise ImportError(f"{LIB_PATH} was built from different sources than the tree (build id "
                              f"{_build.binary_id(LIB_PATH)} != {_build.source_id()}) and cannot be rebuilt: {exc}")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    bid = lib.mpbp_build_id().decode()
    if bid != _build.source_id():
        raise ImportError(f"libmpbp.so build id {bid} does not match the source tree {_build.source_id()}")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().mpbp_last_error_string()
        raise MpbpError(rc, msg.decode() if msg else "")
