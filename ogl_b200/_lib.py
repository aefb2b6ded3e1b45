"""ctypes binding of libogl_b200.so (include/ogl_b200.h).

The product path has no CPU fallback: if the CUDA library is missing this
module raises at import of the symbols, and every compute entry point fails
with OGL_ERR_CUDA when no sm_100-class GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libogl_b200.so")

OGL_OK, OGL_ERR_INVALID, OGL_ERR_CUDA, OGL_ERR_NCCL, OGL_ERR_UNSUPPORTED = range(5)
OGL_NCCL_ID_BYTES = 128
OGL_VEC_B, OGL_VEC_X = 0, 1
OGL_PRECOND_NONE, OGL_PRECOND_BJ, OGL_PRECOND_ISAI, OGL_PRECOND_GISAI = 0, 1, 2, 3
OGL_PRECOND_ILU, OGL_PRECOND_IC, OGL_PRECOND_IRILU, OGL_PRECOND_MULTIGRID = 4, 5, 6, 7
OGL_SOLVER_CG, OGL_SOLVER_BICGSTAB, OGL_SOLVER_GMRES = 0, 1, 2

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
ctx_p = C.c_void_p


class SolveParams(C.Structure):
    _fields_ = [("solver", C.c_int32), ("tolerance", C.c_double), ("rel_tol", C.c_double),
                ("min_iter", C.c_int32), ("max_iter", C.c_int32), ("frequency", C.c_int32),
                ("krylov_dim", C.c_int32), ("export_res", C.c_int32)]


class SolveResult(C.Structure):
    _fields_ = [("init_residual", C.c_double), ("final_residual", C.c_double),
                ("norm_factor", C.c_double), ("criterion_calls", C.c_int32),
                ("n_iterations", C.c_int32), ("solve_us", C.c_double),
                ("resnorm_us", C.c_double), ("kernel_launches", C.c_int64),
                ("spmv_us_avg", C.c_double), ("spmv_samples", C.c_int32)]


# every symbol include/ogl_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ogl_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "ogl_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ogl_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.POINTER(ctx_p)]),
    "ogl_ctx_destroy": (C.c_int, [ctx_p]),
    "ogl_last_error": (C.c_char_p, [ctx_p]),
    "ogl_set_option": (C.c_int, [ctx_p, C.c_char_p, C.c_int64]),
    "ogl_get_option": (C.c_int, [ctx_p, C.c_char_p, i64p]),
    "ogl_pattern_from_ldu": (C.c_int, [ctx_p, C.c_int32, C.c_int32, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ogl_pattern_nnz": (C.c_int, [ctx_p, i64p, i64p]),
    "ogl_pattern_download": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ogl_partition_create": (C.c_int, [ctx_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "ogl_partition_sizes": (C.c_int, [ctx_p, i64p, i64p]),
    "ogl_partition_export": (C.c_int, [ctx_p, C.c_void_p, C.c_int64, i64p]),
    "ogl_partition_connect": (C.c_int, [ctx_p, C.c_void_p, C.c_int64]),
    "ogl_nonlocal_pattern": (C.c_int, [ctx_p, C.c_int32, C.c_void_p]),
    "ogl_nonlocal_pattern_download": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ogl_values_update": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_double]),
    "ogl_values_download": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p]),
    "ogl_vector_upload": (C.c_int, [ctx_p, C.c_int, C.c_void_p, C.c_double]),
    "ogl_vector_download": (C.c_int, [ctx_p, C.c_int, C.c_void_p]),
    "ogl_vector_fill": (C.c_int, [ctx_p, C.c_int, C.c_double]),
    "ogl_precond_setup": (C.c_int, [ctx_p, C.c_int, C.c_int32, C.c_int]),
    "ogl_precond_download": (C.c_int, [ctx_p, i32p, C.c_void_p, C.c_void_p]),
    "ogl_precond_factors_download": (C.c_int, [ctx_p, C.c_void_p]),
    "ogl_precond_apply": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p]),
    "ogl_mg_levels": (C.c_int, [ctx_p, i32p]),
    "ogl_mg_level_info": (C.c_int, [ctx_p, C.c_int32, i32p, i32p, i32p]),
    "ogl_mg_level_download": (C.c_int, [ctx_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ogl_solve": (C.c_int, [ctx_p, C.POINTER(SolveParams), C.POINTER(SolveResult)]),
    "ogl_residual_history": (C.c_int, [ctx_p, C.c_void_p, C.c_int32, i32p]),
    "ogl_spmv": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p]),
    "ogl_spmv_bench": (C.c_int, [ctx_p, C.c_int32, C.c_int, C.POINTER(C.c_float)]),
    "ogl_pcg_bench": (C.c_int, [ctx_p, C.c_int32, C.POINTER(C.c_float)]),
    "ogl_synchronize": (C.c_int, [ctx_p]),
    "ogl_trace_download": (C.c_int, [ctx_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "ogl_membench": (C.c_int, [ctx_p, C.c_int, C.c_int64, C.c_int32, C.POINTER(C.c_double)]),
    "ogl_commbench": (C.c_int, [ctx_p, C.c_int, C.c_int32, C.POINTER(C.c_double)]),
    "ogl_export_mtx": (C.c_int, [ctx_p, C.c_int, C.c_char_p]),
}

_lib = None


def _preload_bundled_nccl() -> None:
    """libogl_b200.so needs `libnccl.so.2`.  PyTorch ships its own (newer) copy
    and resolves it by SONAME: whichever copy is mapped first serves the whole
    process, and torch fails to import on top of the older system copy.  Map
    torch's copy first so that both sides share one NCCL."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    if spec and spec.submodule_search_locations:
        for base in spec.submodule_search_locations:
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return


def load() -> C.CDLL:
    """Load libogl_b200.so; raises when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        _preload_bundled_nccl()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
