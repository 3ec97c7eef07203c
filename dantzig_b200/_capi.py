"""ctypes binding of include/dantzig_b200.h (libdantzig_b200.so).

Loading fails loudly when the CUDA library has not been built: there is no
Python or CPU implementation to fall back to.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libdantzig_b200.so")
# Test-only override: the CPU suite points the binding at the SIMT-emulator build of the same
# sources (tests/emu).  Honoured only together with DZ_LIB_TEST_ONLY=1, which nothing but the
# test harness sets, so a stray DZ_LIB cannot turn the product into a CPU path.
if os.environ.get("DZ_LIB") and os.environ.get("DZ_LIB_TEST_ONLY") == "1":
    LIB_PATH = os.environ["DZ_LIB"]

OK, ERR_ARG, ERR_CUDA, ERR_LIMIT, ERR_ALLOC = 0, -1, -2, -3, -4
OPTIMAL, UNBOUNDED, INFEASIBLE, BREAKDOWN, PIVOT_CAP = range(5)
NUMERICS_EXACT, NUMERICS_FAST = 0, 1
STATUS_NAMES = ["optimal", "unbounded", "infeasible", "breakdown", "pivot_cap"]

# every symbol include/dantzig_b200.h declares (tests check the export list)
SYMBOLS = [
    "dz_template_create", "dz_template_destroy", "dz_template_get_info",
    "dz_template_get_arrays", "dz_template_pack_theta", "dz_options_default",
    "dz_solve_batch", "dz_solve_batch_multi", "dz_batch_create", "dz_batch_destroy", "dz_batch_upload",
    "dz_batch_pack_dense", "dz_batch_solve", "dz_batch_download", "dz_batch_sync", "dz_batch_last_timing",
    "dz_batch_launch_info", "dz_batch_io_bytes", "dz_solve_model", "dz_last_error",
    "dz_device_count", "dz_device_info", "dz_version", "dz_measure_fp64_peak",
]


class DzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dantzig_b200 error {code}: {msg}")
        self.code = code


class Model(C.Structure):
    _fields_ = [
        ("n_vars", C.c_int32),
        ("has_lb", C.c_void_p), ("has_ub", C.c_void_p),
        ("lb", C.c_void_p), ("ub", C.c_void_p),
        ("n_obj", C.c_int32),
        ("obj_var", C.c_void_p), ("obj_coef", C.c_void_p),
        ("obj_const", C.c_double),
        ("n_rows", C.c_int32),
        ("row_ptr", C.c_void_p), ("row_var", C.c_void_p),
        ("row_coef", C.c_void_p), ("rhs", C.c_void_p),
    ]


class TemplateInfo(C.Structure):
    _fields_ = [
        ("m", C.c_int32), ("n_int", C.c_int32), ("n_orig", C.c_int32),
        ("nnz", C.c_int64), ("n_theta", C.c_int64),
    ]


class Options(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("max_pivots", C.c_int64), ("trace_cap", C.c_int32),
        ("worker_warps", C.c_int32), ("ctas_per_sm", C.c_int32), ("stream", C.c_void_p),
        ("profile", C.c_int32), ("basis_home", C.c_int32), ("numerics", C.c_int32),
    ]


class BatchResult(C.Structure):
    _fields_ = [
        ("status", C.c_void_p), ("pivots", C.c_void_p), ("n_primal", C.c_void_p),
        ("trace_hash", C.c_void_p), ("objective", C.c_void_p), ("values", C.c_void_p),
        ("x_basic", C.c_void_p), ("basis", C.c_void_p), ("trace", C.c_void_p),
        ("work", C.c_void_p), ("prof", C.c_void_p),
    ]


class Solution(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("pivots", C.c_int32), ("n_primal", C.c_int32),
        ("trace_hash", C.c_uint64), ("objective", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dantzig_b200.build` "
            "(nvcc, sm_100a).  dantzig_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.dz_template_create.argtypes = [C.POINTER(Model), C.POINTER(vp)]
    L.dz_template_destroy.argtypes = [vp]
    L.dz_template_destroy.restype = None
    L.dz_template_get_info.argtypes = [vp, C.POINTER(TemplateInfo)]
    L.dz_template_get_arrays.argtypes = [vp] * 11
    L.dz_template_pack_theta.argtypes = [vp, C.POINTER(Model), vp]
    L.dz_options_default.argtypes = [C.POINTER(Options)]
    L.dz_options_default.restype = None
    L.dz_solve_batch.argtypes = [vp, i64, vp, C.POINTER(Options), C.POINTER(BatchResult)]
    L.dz_batch_pack_dense.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32]
    L.dz_solve_batch_multi.argtypes = [vp, i64, vp, i32, C.POINTER(Options), C.POINTER(BatchResult)]
    L.dz_batch_create.argtypes = [vp, i64, C.POINTER(Options), C.POINTER(vp)]
    L.dz_batch_destroy.argtypes = [vp]
    L.dz_batch_destroy.restype = None
    L.dz_batch_upload.argtypes = [vp, vp]
    L.dz_batch_solve.argtypes = [vp]
    L.dz_batch_download.argtypes = [vp, C.POINTER(BatchResult)]
    L.dz_batch_sync.argtypes = [vp]
    L.dz_batch_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    L.dz_batch_launch_info.argtypes = [vp] + [C.POINTER(i32)] * 5
    L.dz_batch_io_bytes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.dz_solve_model.argtypes = [C.POINTER(Model), C.POINTER(Options), C.POINTER(Solution), vp]
    L.dz_last_error.restype = C.c_char_p
    L.dz_device_count.restype = C.c_int
    L.dz_device_info.argtypes = [
        C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
        C.POINTER(C.c_int), C.POINTER(i64),
    ]
    L.dz_version.restype = C.c_int
    L.dz_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        raise DzError(rc, lib().dz_last_error().decode("utf-8", "replace"))
