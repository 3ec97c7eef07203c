"""Array form of the model that crosses the C ABI (``dz_model`` in
include/dantzig_b200.h) and helpers that build it.

``ModelArrays`` is exactly what ``rust.solve(objective, constraints)`` receives
in the reference (/root/reference/src/lib.rs:16-27): a MAXIMISATION objective
and a list of ``linexpr <= b`` rows over variables that carry optional bounds.
The helpers restate how the reference's Python frontend turns user-level
``==``/``<=``/``>=`` rows and ``Minimize`` into that form
(/root/reference/python-source/dantzig/model.py:323-375, optimize.py:114-117).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi


@dataclass
class ModelArrays:
    n_vars: int
    has_lb: np.ndarray
    has_ub: np.ndarray
    lb: np.ndarray
    ub: np.ndarray
    obj_var: np.ndarray
    obj_coef: np.ndarray
    obj_const: float
    row_ptr: np.ndarray
    row_var: np.ndarray
    row_coef: np.ndarray
    rhs: np.ndarray
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self) -> None:
        self.has_lb = np.ascontiguousarray(self.has_lb, dtype=np.uint8)
        self.has_ub = np.ascontiguousarray(self.has_ub, dtype=np.uint8)
        self.lb = np.ascontiguousarray(self.lb, dtype=np.float64)
        self.ub = np.ascontiguousarray(self.ub, dtype=np.float64)
        self.obj_var = np.ascontiguousarray(self.obj_var, dtype=np.int32)
        self.obj_coef = np.ascontiguousarray(self.obj_coef, dtype=np.float64)
        self.row_ptr = np.ascontiguousarray(self.row_ptr, dtype=np.int64)
        self.row_var = np.ascontiguousarray(self.row_var, dtype=np.int32)
        self.row_coef = np.ascontiguousarray(self.row_coef, dtype=np.float64)
        self.rhs = np.ascontiguousarray(self.rhs, dtype=np.float64)
        if len(self.row_ptr) != len(self.rhs) + 1:
            raise ValueError("row_ptr must have n_rows+1 entries")

    @property
    def n_rows(self) -> int:
        return len(self.rhs)

    def as_c(self) -> _capi.Model:
        m = _capi.Model()
        m.n_vars = int(self.n_vars)
        m.n_obj = len(self.obj_var)
        m.obj_const = float(self.obj_const)
        m.n_rows = self.n_rows
        for name in ("has_lb", "has_ub", "lb", "ub", "obj_var", "obj_coef", "row_ptr",
                     "row_var", "row_coef", "rhs"):
            setattr(m, name, getattr(self, name).ctypes.data)
        return m


class ModelBuilder:
    """Term-by-term construction, the shape of the reference's unit tests
    (AffExpr::new / Inequality::new, src/simplex.rs:485-796)."""

    def __init__(self) -> None:
        self._lb: list[float | None] = []
        self._ub: list[float | None] = []
        self._obj: list[tuple[float, int]] = []
        self._const = 0.0
        self._rows: list[tuple[list[tuple[float, int]], float]] = []

    def var(self, lb: float | None = None, ub: float | None = None) -> int:
        self._lb.append(lb)
        self._ub.append(ub)
        return len(self._lb) - 1

    def nonneg(self) -> int:
        return self.var(lb=0.0)

    def free(self) -> int:
        return self.var()

    def maximize(self, terms: list[tuple[float, int]], constant: float = 0.0) -> "ModelBuilder":
        self._obj = [(float(c), int(v)) for c, v in terms]
        self._const = float(constant)
        return self

    def minimize(self, terms: list[tuple[float, int]], constant: float = 0.0) -> "ModelBuilder":
        # Minimize.solve negates the objective before the call (optimize.py:115)
        return self.maximize([(-float(c), v) for c, v in terms], -float(constant))

    def leq(self, terms: list[tuple[float, int]], b: float) -> "ModelBuilder":
        self._rows.append(([(float(c), int(v)) for c, v in terms], float(b)))
        return self

    def geq(self, terms: list[tuple[float, int]], b: float) -> "ModelBuilder":
        # Constraint.greater_than_eq, model.py:367-375
        return self.leq([(-float(c), v) for c, v in terms], -float(b))

    def eq(self, terms: list[tuple[float, int]], b: float) -> "ModelBuilder":
        # Constraint.equality, model.py:351-359: the <= row then the negated row
        self.leq(terms, b)
        return self.geq(terms, b)

    def build(self) -> ModelArrays:
        n = len(self._lb)
        row_ptr = [0]
        rv: list[int] = []
        rc: list[float] = []
        for terms, _ in self._rows:
            for c, v in terms:
                rv.append(v)
                rc.append(c)
            row_ptr.append(len(rv))
        return ModelArrays(
            n_vars=n,
            has_lb=[x is not None for x in self._lb],
            has_ub=[x is not None for x in self._ub],
            lb=[0.0 if x is None else x for x in self._lb],
            ub=[0.0 if x is None else x for x in self._ub],
            obj_var=[v for _, v in self._obj],
            obj_coef=[c for c, _ in self._obj],
            obj_const=self._const,
            row_ptr=row_ptr, row_var=rv, row_coef=rc,
            rhs=[b for _, b in self._rows],
        )


LE, GE, EQ = 0, 1, 2


def dense_structure(m: int, n: int, senses, has_lb, has_ub) -> ModelArrays:
    """Structure-only model of a dense user-level LP (every row mentions every
    variable in index order; the objective mentions every variable)."""
    senses = np.asarray(senses, dtype=np.int32)
    n_low = int(np.sum(senses != EQ) + 2 * np.sum(senses == EQ))
    row_ptr = np.arange(n_low + 1, dtype=np.int64) * n
    row_var = np.tile(np.arange(n, dtype=np.int32), n_low)
    return ModelArrays(
        n_vars=n, has_lb=has_lb, has_ub=has_ub, lb=np.zeros(n), ub=np.zeros(n),
        obj_var=np.arange(n, dtype=np.int32), obj_coef=np.zeros(n), obj_const=0.0,
        row_ptr=row_ptr, row_var=row_var, row_coef=np.zeros(n_low * n), rhs=np.zeros(n_low),
    )


def dense_theta(A: np.ndarray, b: np.ndarray, c: np.ndarray, senses, lb, ub, has_lb, has_ub,
                minimize: bool = True) -> np.ndarray:
    """Parameter vectors theta[B, P] (layout in include/dantzig_b200.h) for a
    batch of dense user-level LPs  min/max c.x  s.t.  A x (senses) b, lb<=x<=ub,
    lowered the way the reference frontend does it."""
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 2:
        A, b, c = A[None], np.asarray(b)[None], np.asarray(c)[None]
    B, m, n = A.shape
    b = np.asarray(b, dtype=np.float64).reshape(B, m)
    c = np.asarray(c, dtype=np.float64).reshape(B, n)
    senses = np.asarray(senses, dtype=np.int32)
    rows_coef, rows_rhs = [], []
    for i in range(m):
        s = senses[i]
        if s == LE or s == EQ:
            rows_coef.append(A[:, i, :])
            rows_rhs.append(b[:, i])
        if s == GE or s == EQ:
            rows_coef.append(-A[:, i, :])
            rows_rhs.append(-b[:, i])
    n_low = len(rows_rhs)
    P = 2 + n + n_low * n + n_low + 2 * n
    theta = np.empty((B, P), dtype=np.float64)
    theta[:, 0] = 1.0
    theta[:, 1] = -0.0 if minimize else 0.0
    theta[:, 2:2 + n] = -c if minimize else c
    o = 2 + n
    if n_low:
        theta[:, o:o + n_low * n] = np.stack(rows_coef, axis=1).reshape(B, n_low * n)
        o += n_low * n
        theta[:, o:o + n_low] = np.stack(rows_rhs, axis=1)
        o += n_low
    lb = np.where(np.asarray(has_lb, bool), np.asarray(lb, np.float64), 0.0)
    ub = np.where(np.asarray(has_ub, bool), np.asarray(ub, np.float64), 0.0)
    theta[:, o:o + n] = lb
    theta[:, o + n:o + 2 * n] = ub
    return theta


def model_from_theta(structure: ModelArrays, theta: np.ndarray) -> ModelArrays:
    """Inverse of packing: the numeric model one row of theta describes."""
    n_obj, T, nr, nv = len(structure.obj_var), len(structure.row_var), structure.n_rows, structure.n_vars
    o = 2
    obj = theta[o:o + n_obj]; o += n_obj
    rc = theta[o:o + T]; o += T
    rhs = theta[o:o + nr]; o += nr
    lb = theta[o:o + nv]; o += nv
    ub = theta[o:o + nv]
    return ModelArrays(
        n_vars=nv, has_lb=structure.has_lb, has_ub=structure.has_ub, lb=lb.copy(), ub=ub.copy(),
        obj_var=structure.obj_var, obj_coef=obj.copy(), obj_const=float(theta[1]),
        row_ptr=structure.row_ptr, row_var=structure.row_var, row_coef=rc.copy(), rhs=rhs.copy(),
    )
