"""ctypes declarations of the C ABI in ``include/jfnk.h`` (one-to-one; no logic here)."""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 1

# status codes
OK, NO_CONVERGENCE, ZERO_STEP, NONFINITE, INVALID, CUDA_ERROR, NCCL_ERROR = range(7)
# problems
PROBLEM_SH, PROBLEM_SH_LINEAR, PROBLEM_PMA2, PROBLEM_DROPLET = range(4)
# Gram-Schmidt modes
GS_CGS, GS_CGS_IFNEEDED, GS_CGS2 = range(3)
GS_MODES = {"cgs": GS_CGS, "cgs-ifneeded": GS_CGS_IFNEEDED, "cgs2": GS_CGS2}


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("problem", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("row0", C.c_int32), ("nrows", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("inner_m", C.c_int32), ("outer_k", C.c_int32),
        ("gs_mode", C.c_int32), ("kernel_variant", C.c_int32),
        ("gs_tau", C.c_double),
        ("stream", C.c_void_p),
    ]


class NewtonOpts(C.Structure):
    _fields_ = [
        ("f_tol", C.c_double), ("f_rtol", C.c_double), ("x_tol", C.c_double), ("x_rtol", C.c_double),
        ("rdiff", C.c_double), ("maxiter", C.c_int64), ("iter", C.c_int32), ("line_search", C.c_int32),
    ]


class History(C.Structure):
    _fields_ = [
        ("capacity", C.c_int32), ("count", C.c_int32),
        ("nfev", C.c_int64), ("inner_iters", C.c_int64), ("reorth", C.c_int64),
        ("f0_max", C.c_double), ("f0_l2", C.c_double),
        ("f_max", C.POINTER(C.c_double)), ("f_l2", C.POINTER(C.c_double)), ("step", C.POINTER(C.c_double)),
        ("inner", C.POINTER(C.c_int32)),
    ]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("ms", C.c_double), ("bytes", C.c_double)]


CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double)
PSOLVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)

_P = C.c_void_p  # device pointers travel as plain addresses
_CTX = C.c_void_p

# name -> (restype, argtypes); every symbol include/jfnk.h declares
SIGNATURES = {
    "jfnk_abi_version": (C.c_int, []),
    "jfnk_last_error": (C.c_char_p, []),
    "jfnk_device_ok": (C.c_int, []),
    "jfnk_workspace_bytes": (C.c_size_t, [C.POINTER(Config)]),
    "jfnk_create": (C.c_int, [C.POINTER(Config), _P, C.c_size_t, C.POINTER(_CTX)]),
    "jfnk_destroy": (C.c_int, [_CTX]),
    "jfnk_set_callback": (C.c_int, [_CTX, CALLBACK, C.c_void_p]),
    "jfnk_request_stop": (C.c_int, [_CTX]),
    "jfnk_set_preconditioner": (C.c_int, [_CTX, PSOLVE, C.c_void_p]),
    "jfnk_comm_unique_id": (C.c_int, [C.c_void_p]),
    "jfnk_comm_init": (C.c_int, [_CTX, C.c_void_p]),
    "jfnk_comm_peer_memory": (C.c_int, [_CTX]),
    "jfnk_comm_bench": (C.c_int, [_CTX, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_double)]),
    "jfnk_sh_setup": (C.c_int, [_CTX, C.c_double, C.c_double, C.c_double, C.c_double]),
    "jfnk_spmv_lap": (C.c_int, [_CTX, _P, _P]),
    "jfnk_spmv_sh": (C.c_int, [_CTX, _P, _P]),
    "jfnk_set_prev": (C.c_int, [_CTX, _P]),
    "jfnk_shlin_prepare": (C.c_int, [_CTX, _P, _P, _P]),
    "jfnk_shlin_step": (C.c_int, [_CTX, _P, _P, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "jfnk_residual": (C.c_int, [_CTX, _P, _P]),
    "jfnk_linearize": (C.c_int, [_CTX, _P, C.c_double]),
    "jfnk_jvp": (C.c_int, [_CTX, _P, _P]),
    "jfnk_lgmres_reset": (C.c_int, [_CTX]),
    "jfnk_lgmres": (C.c_int, [_CTX, _P, _P, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "jfnk_newton": (C.c_int, [_CTX, _P, C.POINTER(NewtonOpts), C.POINTER(History)]),
    "jfnk_sh_step": (C.c_int, [_CTX, _P, C.c_int, C.POINTER(NewtonOpts), C.POINTER(History)]),
    "jfnk_mesh_setup": (C.c_int, [_CTX] + [C.c_double] * 6),
    "jfnk_mesh_set_potential": (C.c_int, [_CTX, _P]),
    "jfnk_mesh_laplace": (C.c_int, [_CTX, _P, _P, _P]),
    "jfnk_pma2_setup": (C.c_int, [_CTX, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]),
    "jfnk_pma2_set_prev": (C.c_int, [_CTX, _P]),
    "jfnk_droplet_setup": (C.c_int, [_CTX, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "jfnk_droplet_set_prev": (C.c_int, [_CTX, _P, C.c_double]),
    "jfnk_droplet_shape": (C.c_int, [_CTX, _P, C.c_int, C.POINTER(C.c_double), C.c_double, _P]),
    "jfnk_mesh_relax": (C.c_int, [_CTX, _P, _P, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]),
    "jfnk_launch_count": (C.c_int64, [_CTX]),
    "jfnk_profile_enable": (C.c_int, [_CTX, C.c_int]),
    "jfnk_profile_read": (C.c_int, [_CTX, C.POINTER(KernelStat), C.c_int, C.POINTER(C.c_int)]),
    "jfnk_multi_dot": (C.c_int, [_CTX, C.c_int, _P, C.c_size_t, _P, C.POINTER(C.c_double)]),
    "jfnk_multi_axpy": (C.c_int, [_CTX, C.c_int, _P, C.c_size_t, C.POINTER(C.c_double), _P, C.POINTER(C.c_double)]),
}


def bind(path: str) -> C.CDLL:
    """dlopen ``path`` and attach the prototypes; raises if any declared symbol is missing."""
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.jfnk_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.jfnk_abi_version()} != {ABI_VERSION}")
    return lib
