"""ctypes binding of libcyrk_b200.so (include/cyrk_b200.h).

The product path has NO CPU fallback: if the shared library is missing, or no CUDA device is
present, importing works (so the ABI can be inspected on a CPU box) but every solve raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CYRK_B200_LIB", os.path.join(_HERE, "libcyrk_b200.so"))

c_size_p = C.POINTER(C.c_size_t)
c_dbl_p = C.POINTER(C.c_double)

# Every symbol include/cyrk_b200.h declares: name -> (restype, argtypes)
_SOLVE_ARGS = [
    C.c_int, C.c_int, C.c_size_t,            # rhs_id, method, n_traj
    C.c_void_p, C.c_size_t,                  # t_span, n_t_span
    C.c_void_p, C.c_size_t,                  # y0, n_y
    C.c_size_t,                              # num_extra
    C.c_void_p, C.c_size_t, C.c_size_t,      # args, args_len, n_args
    C.c_size_t, C.c_size_t,                  # max_num_steps, max_ram_MB
    C.c_int,                                 # dense_output
    C.c_void_p, C.c_size_t,                  # t_eval, n_t_eval
    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,   # event ids/dirs/max, n_events
    C.c_void_p, C.c_size_t,                  # rtols
    C.c_void_p, C.c_size_t,                  # atols
    C.c_double, C.c_double,                  # max_step, first_step
    C.c_size_t,                              # expected_size
    C.c_int, C.c_uint, C.c_int,              # store_mode, flags, device
]

SIGNATURES = {
    "cyrk_b200_version": (C.c_int, []),
    "cyrk_b200_device_count": (C.c_int, []),
    "cyrk_b200_last_error": (C.c_char_p, []),
    "cyrk_b200_status_message": (None, [C.c_int, C.c_char_p, C.c_size_t]),
    "cyrk_b200_rhs_builtin": (C.c_int, [C.c_char_p]),
    "cyrk_b200_rhs_info": (C.c_int, [C.c_int, c_size_p, c_size_p, c_size_p]),
    "cyrk_b200_rhs_compile": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                        C.c_char_p, C.c_size_t]),
    "cyrk_b200_solve_batch": (C.c_void_p, _SOLVE_ARGS),
    "cyrk_b200_solve_batch_noreturn": (C.c_int, [C.c_void_p] + _SOLVE_ARGS),
    "cyrk_b200_result_create": (C.c_void_p, []),
    "cyrk_b200_free": (None, [C.c_void_p]),
    "cyrk_b200_get_n_traj": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_num_y": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_num_dy": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_status_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_size_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_steps_taken_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_n_rejected_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_nfev_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_t_last_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_y_last_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_row_capacity": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_t_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_y_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_num_events": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_event_capacity": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_event_offsets_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_event_total": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_event_count_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_event_t_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_event_y_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_event_terminate_index_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_get_num_interpolates_capacity": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_num_interpolates_ptr": (C.c_void_p, [C.c_void_p]),
    "cyrk_b200_call_vectorize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "cyrk_b200_ensemble_sums": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cyrk_b200_get_kernel_ms": (C.c_double, [C.c_void_p]),
    "cyrk_b200_get_total_ms": (C.c_double, [C.c_void_p]),
    "cyrk_b200_get_launch_count": (C.c_uint, [C.c_void_p]),
    "cyrk_b200_get_h2d_bytes": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_d2h_bytes": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_get_lazy_d2h_bytes": (C.c_size_t, [C.c_void_p]),
    "cyrk_b200_jit_compile": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint]),
    "cyrk_b200_jit_cache_stats": (None, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "cyrk_b200_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "cyrk_b200_measure_fp64_tflops": (C.c_double, [C.c_int]),
    "cyrk_b200_solve_batch_sharded": (C.c_int, [C.POINTER(C.c_int), C.c_size_t, C.POINTER(C.c_void_p), c_size_p] + _SOLVE_ARGS[:-1]),
    "cyrk_b200_debug_sincos": (C.c_int, [c_dbl_p, c_dbl_p, c_dbl_p, C.c_size_t, C.c_int]),
}

_lib = None


class CyrkB200Error(RuntimeError):
    """Call-level failure of the native library (CUDA, NVRTC or argument error)."""


def lib():
    """Load libcyrk_b200.so (built by ``__graft_entry__.build()`` / ``make -C cyrk_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CyrkB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)      # raises AttributeError when a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def last_error() -> str:
    return lib().cyrk_b200_last_error().decode("utf-8", "replace")


def status_message(code: int) -> str:
    buf = C.create_string_buffer(256)
    lib().cyrk_b200_status_message(int(code), buf, 256)
    return buf.value.decode()
