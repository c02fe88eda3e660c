"""ctypes binding of the kernel-level C ABI (include/b200vec.h).

The shared library is the product; this module only declares its signatures.
There is no fallback: if the library is missing or was not built, importing the
binding raises, and every call that returns a negative code raises B200VecError
with the library's own message.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libsundials_nvecb200.so"

c_double_p = C.POINTER(C.c_double)
c_ptr_table = C.POINTER(C.c_void_p)
ctx_t = C.c_void_p

B200VEC_SUM, B200VEC_MAX, B200VEC_MIN = 0, 1, 2
UNIQUE_ID_BYTES = 128


class B200VecError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


_lib = None

# name -> (restype, argtypes); device pointers are passed as c_void_p integers
_V = C.c_void_p
_D = C.c_double
_I = C.c_int
_L = C.c_int64
_SIGS = {
    "b200vec_ctx_create": (_I, [C.POINTER(ctx_t), _I, _V]),
    "b200vec_ctx_retain": (_I, [ctx_t]),
    "b200vec_ctx_release": (_I, [ctx_t]),
    "b200vec_ctx_default": (_I, [C.POINTER(ctx_t)]),
    "b200vec_ctx_set_stream": (_I, [ctx_t, _V]),
    "b200vec_ctx_get_stream": (_V, [ctx_t]),
    "b200vec_ctx_device": (_I, [ctx_t]),
    "b200vec_ctx_sync": (_I, [ctx_t]),
    "b200vec_ctx_set_tuning": (_I, [ctx_t, C.c_char_p, _L]),
    "b200vec_ctx_get_tuning": (_L, [ctx_t, C.c_char_p]),
    "b200vec_ctx_launch_count": (_L, [ctx_t]),
    "b200vec_last_error": (C.c_char_p, []),
    "b200vec_version": (C.c_char_p, []),
    "b200vec_malloc_device": (_I, [ctx_t, C.c_size_t, C.POINTER(_V)]),
    "b200vec_free_device": (_I, [ctx_t, _V, C.c_size_t]),
    "b200vec_malloc_host": (_I, [ctx_t, C.c_size_t, C.POINTER(_V)]),
    "b200vec_free_host": (_I, [ctx_t, _V]),
    "b200vec_malloc_managed": (_I, [ctx_t, C.c_size_t, C.POINTER(_V)]),
    "b200vec_free_managed": (_I, [ctx_t, _V]),
    "b200vec_copy_h2d": (_I, [ctx_t, _V, _V, C.c_size_t, _I]),
    "b200vec_copy_d2h": (_I, [ctx_t, _V, _V, C.c_size_t, _I]),
    "b200vec_copy_d2d": (_I, [ctx_t, _V, _V, C.c_size_t]),
    "b200vec_copy_h2d_async": (_I, [ctx_t, _V, _V, C.c_size_t]),
    "b200vec_copy_join": (_I, [ctx_t]),
    "b200vec_linear_sum": (_I, [ctx_t, _D, _V, _D, _V, _V, _L]),
    "b200vec_const": (_I, [ctx_t, _D, _V, _L]),
    "b200vec_prod": (_I, [ctx_t, _V, _V, _V, _L]),
    "b200vec_div": (_I, [ctx_t, _V, _V, _V, _L]),
    "b200vec_scale": (_I, [ctx_t, _D, _V, _V, _L]),
    "b200vec_abs": (_I, [ctx_t, _V, _V, _L]),
    "b200vec_inv": (_I, [ctx_t, _V, _V, _L]),
    "b200vec_add_const": (_I, [ctx_t, _V, _D, _V, _L]),
    "b200vec_compare": (_I, [ctx_t, _D, _V, _V, _L]),
    "b200vec_dot_prod": (_I, [ctx_t, _V, _V, _L, c_double_p]),
    "b200vec_max_norm": (_I, [ctx_t, _V, _L, c_double_p]),
    "b200vec_min": (_I, [ctx_t, _V, _L, c_double_p]),
    "b200vec_l1_norm": (_I, [ctx_t, _V, _L, c_double_p]),
    "b200vec_wsqr_sum": (_I, [ctx_t, _V, _V, _L, c_double_p]),
    "b200vec_wsqr_sum_mask": (_I, [ctx_t, _V, _V, _V, _L, c_double_p]),
    "b200vec_inv_test": (_I, [ctx_t, _V, _V, _L, c_double_p]),
    "b200vec_constr_mask": (_I, [ctx_t, _V, _V, _V, _L, c_double_p]),
    "b200vec_min_quotient": (_I, [ctx_t, _V, _V, _L, c_double_p]),
    "b200vec_ewt_set": (_I, [ctx_t, _D, _D, _V, _V, _V, _L, c_double_p]),
    "b200vec_axpy_dot": (_I, [ctx_t, _D, _V, _V, _V, _L, c_double_p]),
    "b200vec_cv_ewt": (_I, [ctx_t, _D, _D, _V, _V, _V, _V, _L]),
    "b200vec_cv_constraints": (_I, [ctx_t, _V, _V, _V, _V, _V, _L]),
    "b200vec_cv_nls_resid": (_I, [ctx_t, _D, _D, _V, _V, _V, _V, _L]),
    "b200vec_cv_diag_form_y": (_I, [ctx_t, _D, _D, _V, _V, _V, _V, _V, _L]),
    "b200vec_cv_diag_build_m": (_I, [ctx_t, _D, _D, _V, _V, _V, _V, _V, _V, _V, _L]),
    "b200vec_cv_diag_update_m": (_I, [ctx_t, _D, _V, _L]),
    "b200vec_mgs_sweep": (_I, [ctx_t, _I, _V, c_ptr_table, _L, c_double_p, c_double_p]),
    "b200vec_cgs_step": (_I, [ctx_t, _I, _V, c_ptr_table, c_ptr_table, _V, _L, c_double_p, c_double_p]),
    "b200vec_result_device": (_V, [ctx_t]),
    "b200vec_result_fetch": (_I, [ctx_t, _I, c_double_p]),
    "b200vec_linear_combination": (_I, [ctx_t, _I, c_double_p, c_ptr_table, _V, _L]),
    "b200vec_scale_add_multi": (_I, [ctx_t, _I, c_double_p, _V, c_ptr_table, c_ptr_table, _L]),
    "b200vec_dot_prod_multi": (_I, [ctx_t, _I, _V, c_ptr_table, _L, c_double_p]),
    "b200vec_linear_combination_sqnorm": (_I, [ctx_t, _I, c_double_p, c_ptr_table, _V, _L, c_double_p]),
    "b200vec_linear_sum_vector_array": (_I, [ctx_t, _I, _D, c_ptr_table, _D, c_ptr_table, c_ptr_table, _I, _I, _L]),
    "b200vec_scale_vector_array": (_I, [ctx_t, _I, c_double_p, c_ptr_table, c_ptr_table, _L]),
    "b200vec_const_vector_array": (_I, [ctx_t, _I, _D, c_ptr_table, _L]),
    "b200vec_wsqr_sum_vector_array": (_I, [ctx_t, _I, c_ptr_table, c_ptr_table, _V, _L, c_double_p]),
    "b200vec_scale_add_multi_vector_array": (_I, [ctx_t, _I, _I, c_double_p, c_ptr_table, c_ptr_table, c_ptr_table,
                                                  _I, _L]),
    "b200vec_linear_combination_vector_array": (_I, [ctx_t, _I, _I, c_double_p, c_ptr_table, c_ptr_table, _I, _L]),
    "b200vec_comm_get_unique_id": (_I, [C.POINTER(C.c_ubyte)]),
    "b200vec_comm_init": (_I, [ctx_t, C.POINTER(C.c_ubyte), _I, _I]),
    "b200vec_comm_finalize": (_I, [ctx_t]),
    "b200vec_comm_peer_alloc": (_I, [ctx_t, C.c_size_t, C.POINTER(_V)]),
    "b200vec_comm_peer_free": (_I, [ctx_t, C.POINTER(_V)]),
    "b200vec_comm_rank": (_I, [ctx_t]),
    "b200vec_comm_size": (_I, [ctx_t]),
    "b200vec_allreduce": (_I, [ctx_t, _I, _I]),
    "b200vec_ctx_set_scope": (_I, [ctx_t, _I]),
    "b200vec_comm_transport": (C.c_char_p, [ctx_t]),
    "b200vec_allreduce_buffer": (_I, [ctx_t, _V, _I, _I]),
    "b200vec_allreduce_i64_host": (_I, [ctx_t, C.POINTER(C.c_int64), _I]),
}

# the symbols include/b200vec.h declares (checked by the CPU test-suite)
EXPORTED_SYMBOLS = tuple(_SIGS)


def load() -> C.CDLL:
    """Load libsundials_nvecb200.so (building it first if this is a source tree
    with nvcc and the library is stale).  Raises if it cannot be provided."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        from . import build as _build

        _build.build()
    if not _LIB_PATH.exists():
        raise B200VecError(f"{_LIB_PATH} is missing: run `python -m sundials_b200.build`")
    lib = C.CDLL(str(_LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200vec_last_error().decode(errors="replace")
        raise B200VecError(f"{what or 'b200vec call'} failed with code {rc}: {msg}")
