"""ctypes binding of the CPU oracle (oracle/libdzo.so).

TEST INFRASTRUCTURE ONLY -- see oracle/dzo.h.  Imported by tests/, by
``__graft_entry__.smoke()`` and by ``bench.py``'s cpu_baseline / reference
legs; never by the product package ``dantzig_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdzo.so")

OPTIMAL, UNBOUNDED, INFEASIBLE, PANIC, PIVOT_CAP = range(5)
LITERAL, SKIP, SPARSE = 0, 1, 2
STATUS_NAMES = ["optimal", "unbounded", "infeasible", "panic", "pivot_cap"]


def build(force: bool = False) -> str:
    """Compile libdzo.so with oracle/Makefile if it is missing or stale."""
    src = [os.path.join(_HERE, f) for f in ("dzo.cpp", "dzo.h", "Makefile")]
    stale = not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libdzo.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Model(C.Structure):
    _fields_ = [
        ("n_vars", C.c_int32),
        ("has_lb", C.c_void_p),
        ("has_ub", C.c_void_p),
        ("lb", C.c_void_p),
        ("ub", C.c_void_p),
        ("n_obj", C.c_int32),
        ("obj_var", C.c_void_p),
        ("obj_coef", C.c_void_p),
        ("obj_const", C.c_double),
        ("n_rows", C.c_int32),
        ("row_ptr", C.c_void_p),
        ("row_var", C.c_void_p),
        ("row_coef", C.c_void_p),
        ("rhs", C.c_void_p),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("pivots", C.c_int64),
        ("n_primal", C.c_int64),
        ("n_dual", C.c_int64),
        ("trace_hash", C.c_uint64),
        ("objective", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.dzo_lower.restype = C.c_void_p
        _lib.dzo_lower.argtypes = [C.POINTER(_Model)]
        _lib.dzo_lowered_free.argtypes = [C.c_void_p]
        _lib.dzo_lowered_from_arrays.restype = C.c_void_p
        _lib.dzo_lowered_from_arrays.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 4 + [
            C.c_double
        ] + [C.c_void_p] * 3
        _lib.dzo_lowered_dims.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.dzo_lowered_get.argtypes = [C.c_void_p] * 12
        _lib.dzo_solve.restype = C.c_int
        _lib.dzo_solve.argtypes = [
            C.c_void_p, C.c_int, C.c_int64, C.POINTER(_Result),
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
        ]
        _lib.dzo_lu_factorize.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _lib.dzo_lu_solve.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int]
        _lib.dzo_neg_t_dot.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5
        _lib.dzo_dense_to_csc.restype = C.c_int64
        _lib.dzo_dense_to_csc.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 3
        _lib.dzo_last_flops.argtypes = [C.c_void_p] * 3
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


@dataclass
class Solved:
    status: int
    pivots: int
    n_primal: int
    n_dual: int
    trace_hash: int
    objective: float
    x_basic: np.ndarray
    basis: np.ndarray
    values: np.ndarray          # per original variable, first-appearance order
    trace: np.ndarray           # [pivots, 3] (kind, leaving, entering), possibly truncated
    flops: tuple[float, float, float]

    @property
    def status_name(self) -> str:
        return STATUS_NAMES[self.status]


class Lowered:
    """Owns a dzo_lowered handle; exposes the arrays Simplex::new produces."""

    def __init__(self, handle: int):
        if not handle:
            raise ValueError("oracle lowering failed (malformed model)")
        self._h = handle
        m, n, nnz, no = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int32()
        lib().dzo_lowered_dims(handle, C.byref(m), C.byref(n), C.byref(nnz), C.byref(no))
        self.m, self.n_int, self.nnz, self.n_orig = m.value, n.value, nnz.value, no.value
        self.col_ptr = np.zeros(self.n_int + 1, np.int64)
        self.row_idx = np.zeros(self.nnz, np.int32)
        self.val = np.zeros(self.nnz, np.float64)
        self.c = np.zeros(self.n_int, np.float64)
        self.b = np.zeros(self.m, np.float64)
        self.basis0 = np.zeros(self.m, np.int32)
        self.nonbasis0 = np.zeros(self.n_int - self.m, np.int32)
        self.orig_var = np.zeros(self.n_orig, np.int32)
        self.pos_index = np.zeros(self.n_orig, np.int32)
        self.neg_index = np.zeros(self.n_orig, np.int32)
        c0 = C.c_double()
        lib().dzo_lowered_get(
            handle, _p(self.col_ptr), _p(self.row_idx), _p(self.val), _p(self.c),
            C.cast(C.byref(c0), C.c_void_p), _p(self.b), _p(self.basis0), _p(self.nonbasis0),
            _p(self.orig_var), _p(self.pos_index), _p(self.neg_index),
        )
        self.c0 = c0.value

    def __del__(self):
        try:
            if self._h:
                lib().dzo_lowered_free(self._h)
                self._h = 0
        except Exception:
            pass

    def solve(self, variant: int = LITERAL, max_pivots: int = 0, trace_cap: int = 0) -> Solved:
        res = _Result()
        x = np.zeros(self.m, np.float64)
        basis = np.zeros(self.m, np.int32)
        values = np.zeros(max(self.n_orig, 1), np.float64)
        trace = np.zeros((max(trace_cap, 1), 3), np.int32)
        lib().dzo_solve(
            self._h, variant, max_pivots, C.byref(res), _p(x), _p(basis), _p(values),
            _p(trace) if trace_cap > 0 else None, trace_cap,
        )
        f = (C.c_double * 3)()
        lib().dzo_last_flops(
            C.cast(C.byref(f, 0), C.c_void_p), C.cast(C.byref(f, 8), C.c_void_p),
            C.cast(C.byref(f, 16), C.c_void_p),
        )
        return Solved(
            res.status, res.pivots, res.n_primal, res.n_dual, res.trace_hash, res.objective,
            x, basis, values[: self.n_orig], trace[: min(res.pivots, trace_cap)],
            (f[0], f[1], f[2]),
        )


def lower(model) -> Lowered:
    """Lower a model given as an object with the dzo_model fields as arrays
    (n_vars, has_lb, has_ub, lb, ub, obj_var, obj_coef, obj_const, row_ptr,
    row_var, row_coef, rhs)."""
    keep = dict(
        has_lb=_arr(model.has_lb, np.uint8), has_ub=_arr(model.has_ub, np.uint8),
        lb=_arr(model.lb, np.float64), ub=_arr(model.ub, np.float64),
        obj_var=_arr(model.obj_var, np.int32), obj_coef=_arr(model.obj_coef, np.float64),
        row_ptr=_arr(model.row_ptr, np.int64), row_var=_arr(model.row_var, np.int32),
        row_coef=_arr(model.row_coef, np.float64), rhs=_arr(model.rhs, np.float64),
    )
    m = _Model()
    m.n_vars = int(model.n_vars)
    m.n_obj = len(keep["obj_var"])
    m.obj_const = float(model.obj_const)
    m.n_rows = len(keep["rhs"])
    for k, v in keep.items():
        setattr(m, k, v.ctypes.data)
    return Lowered(lib().dzo_lower(C.byref(m)))


def lowered_from_arrays(m, n_int, col_ptr, row_idx, val, c, c0, b, basis, nonbasis) -> Lowered:
    a = [
        _arr(col_ptr, np.int64), _arr(row_idx, np.int32), _arr(val, np.float64),
        _arr(c, np.float64),
    ]
    t = [_arr(b, np.float64), _arr(basis, np.int32), _arr(nonbasis, np.int32)]
    h = lib().dzo_lowered_from_arrays(
        int(m), int(n_int), *[_p(x) for x in a], float(c0), *[_p(x) for x in t]
    )
    return Lowered(h)


def lu_factorize(a: np.ndarray):
    """Matrix::factorize: returns (packed LU row-major, p)."""
    a = np.array(a, dtype=np.float64, order="C")
    n = a.shape[0]
    p = np.zeros(max(n - 1, 1), np.int32)
    lib().dzo_lu_factorize(_p(a), n, _p(p))
    return a, p[: n - 1]


def lu_solve(a: np.ndarray, b: np.ndarray, variant: int = LITERAL) -> np.ndarray:
    a = np.array(a, dtype=np.float64, order="C")
    b = np.array(b, dtype=np.float64)
    lib().dzo_lu_solve(_p(a), a.shape[0], _p(b), variant)
    return b


def dense_to_csc(dense: np.ndarray):
    dense = np.array(dense, dtype=np.float64, order="C")
    nr, nc = dense.shape
    col_ptr = np.zeros(nc + 1, np.int64)
    row_idx = np.zeros(nr * nc, np.int32)
    val = np.zeros(nr * nc, np.float64)
    nnz = lib().dzo_dense_to_csc(_p(dense), nr, nc, _p(col_ptr), _p(row_idx), _p(val))
    return col_ptr, row_idx[:nnz], val[:nnz]


def neg_t_dot(nrows, ncols, col_ptr, row_idx, val, v) -> np.ndarray:
    out = np.zeros(ncols, np.float64)
    a = [_arr(col_ptr, np.int64), _arr(row_idx, np.int32), _arr(val, np.float64),
         _arr(v, np.float64)]
    lib().dzo_neg_t_dot(nrows, ncols, *[_p(x) for x in a], _p(out))
    return out
