"""Python host API over the C ABI: templates, device-resident batches, solves.

Everything numeric happens in libdantzig_b200.so (CUDA); this module only
marshals numpy buffers.  It mirrors the reference's solver interface:

    reference                                   here
    ------------------------------------------  ---------------------------------
    Simplex::new(objective, constraints)        Template(structure)  (+ theta)
    Simplex::solve / objective_value/solution   solve_model(model) -> Solution
    (none: one LP per call)                     solve_batch(template, theta)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import BREAKDOWN, INFEASIBLE, OPTIMAL, PIVOT_CAP, UNBOUNDED  # noqa: F401
from .model import ModelArrays


def _vp(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Template:
    """Host-side lowering of a model STRUCTURE (replaces Simplex::new)."""

    def __init__(self, structure: ModelArrays):
        self._h = C.c_void_p()
        self._structure = structure
        cm = structure.as_c()
        _capi.check(_capi.lib().dz_template_create(C.byref(cm), C.byref(self._h)))
        info = _capi.TemplateInfo()
        _capi.check(_capi.lib().dz_template_get_info(self._h, C.byref(info)))
        self.m, self.n_int, self.n_orig = info.m, info.n_int, info.n_orig
        self.nnz, self.n_theta = info.nnz, info.n_theta

    def __del__(self):
        try:
            if self._h:
                _capi.lib().dz_template_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def arrays(self) -> dict[str, np.ndarray]:
        out = dict(
            col_ptr=np.zeros(self.n_int + 1, np.int64), row_idx=np.zeros(self.nnz, np.int32),
            val_ref=np.zeros(self.nnz, np.int32), c_ref=np.zeros(self.n_int, np.int32),
            b_ref=np.zeros(self.m, np.int32), basis0=np.zeros(self.m, np.int32),
            nonbasis0=np.zeros(self.n_int - self.m, np.int32),
            orig_var=np.zeros(self.n_orig, np.int32), pos_index=np.zeros(self.n_orig, np.int32),
            neg_index=np.zeros(self.n_orig, np.int32),
        )
        _capi.check(_capi.lib().dz_template_get_arrays(self._h, *[_vp(a) for a in out.values()]))
        return out

    def pack_theta(self, model: ModelArrays) -> np.ndarray:
        theta = np.zeros(self.n_theta, np.float64)
        cm = model.as_c()
        _capi.check(_capi.lib().dz_template_pack_theta(self._h, C.byref(cm), _vp(theta)))
        return theta

    def lowered_values(self, theta: np.ndarray) -> dict[str, np.ndarray]:
        """Numeric lowered arrays (val, c, b, c0) one theta row stands for."""
        a = self.arrays()

        def deref(ref):
            ref = np.asarray(ref)
            v = np.where(ref >= 0, theta[np.maximum(ref, 0) >> 1], 0.0)
            return np.where((ref >= 0) & ((ref & 1) == 1), -v, v)

        return dict(val=deref(a["val_ref"]), c=deref(a["c_ref"]), b=deref(a["b_ref"]),
                    c0=float(theta[1]))


@dataclass
class BatchResult:
    status: np.ndarray
    pivots: np.ndarray
    n_primal: np.ndarray
    trace_hash: np.ndarray
    objective: np.ndarray
    values: np.ndarray      # [B, n_orig], first-appearance order (Template.arrays()["orig_var"])
    x_basic: np.ndarray | None
    basis: np.ndarray | None
    trace: np.ndarray | None
    work: np.ndarray        # [B, 8] executed flops (LU, solves, pricing, updates) + kernel statistics
    prof: np.ndarray | None = None  # [B, 16] phase cycles (profile=True)


def _options(device=0, max_pivots=0, trace_cap=0, worker_warps=0, ctas_per_sm=0, stream=None,
             profile=False, basis_home=0, numerics="exact"):
    o = _capi.Options()
    _capi.lib().dz_options_default(C.byref(o))
    o.device, o.max_pivots, o.trace_cap = int(device), int(max_pivots), int(trace_cap)
    o.worker_warps, o.ctas_per_sm = int(worker_warps), int(ctas_per_sm)
    o.basis_home = int(basis_home)
    # "exact" (default): the reference's floating-point order, bit-identical pivot sequences;
    # "fast": OPT-IN fast numerics (dz_fast.cu), agrees to rounding only
    o.numerics = {"exact": _capi.NUMERICS_EXACT, "fast": _capi.NUMERICS_FAST}[numerics]
    o.stream = stream
    o.profile = 1 if profile else 0
    return o


class Batch:
    """B LPs sharing one template, resident on one GPU."""

    def __init__(self, template: Template, B: int, *, device: int = 0, max_pivots: int = 0,
                 trace_cap: int = 0, worker_warps: int = 0, ctas_per_sm: int = 0,
                 stream: int | None = None, want_basis: bool = False, profile: bool = False,
                 basis_home: int = 0, numerics: str = "exact"):
        self.template, self.B = template, int(B)
        self.trace_cap, self.want_basis, self.profile = int(trace_cap), want_basis, profile
        self._h = C.c_void_p()
        o = _options(device, max_pivots, trace_cap, worker_warps, ctas_per_sm,
                     stream, profile, basis_home, numerics)
        _capi.check(_capi.lib().dz_batch_create(template.handle, self.B, C.byref(o),
                                                C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            _capi.lib().dz_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, theta: np.ndarray) -> None:
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.size != self.B * self.template.n_theta:
            raise ValueError("theta must hold B * n_theta doubles")
        self._theta_keep = theta
        _capi.check(_capi.lib().dz_batch_upload(self._h, _vp(theta)))

    def upload_ptr(self, ptr: int) -> None:
        """Upload from a raw host pointer (e.g. pinned memory owned by the caller)."""
        _capi.check(_capi.lib().dz_batch_upload(self._h, C.c_void_p(ptr)))

    def pack_dense(self, A, b, c, senses, lb=None, ub=None, *, minimize: bool = True, on_device: bool = False) -> None:
        """dz_batch_pack_dense: the batch's parameter vectors written on the device from the plain
        user-level arrays (host numpy arrays, or raw device pointers when on_device)."""
        senses = np.ascontiguousarray(senses, dtype=np.int32)
        m = len(senses)
        if on_device:
            pa, pb, pc = (C.c_void_p(int(x)) for x in (A, b, c))
            n = self.template.n_orig
        else:
            A = np.ascontiguousarray(A, dtype=np.float64).reshape(self.B, m, -1)
            n = A.shape[2]
            b = np.ascontiguousarray(b, dtype=np.float64).reshape(self.B, m)
            c = np.ascontiguousarray(c, dtype=np.float64).reshape(self.B, n)
            self._dense_keep = (A, b, c)
            pa, pb, pc = _vp(A), _vp(b), _vp(c)
        lbv = None if lb is None else np.ascontiguousarray(lb, dtype=np.float64)
        ubv = None if ub is None else np.ascontiguousarray(ub, dtype=np.float64)
        _capi.check(_capi.lib().dz_batch_pack_dense(
            self._h, pa, pb, pc, _vp(senses), None if lbv is None else _vp(lbv),
            None if ubv is None else _vp(ubv), m, n, 1 if minimize else 0, 1 if on_device else 0))

    def solve(self) -> None:
        _capi.check(_capi.lib().dz_batch_solve(self._h))

    def sync(self) -> None:
        _capi.check(_capi.lib().dz_batch_sync(self._h))

    def kernel_ms(self) -> float:
        ms, n = C.c_float(), C.c_int32()
        _capi.check(_capi.lib().dz_batch_last_timing(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value)

    def launches(self) -> int:
        """Kernel launches of the last solve (2 when a hand-over launch follows the main one)."""
        ms, n = C.c_float(), C.c_int32()
        _capi.check(_capi.lib().dz_batch_last_timing(self._h, C.byref(ms), C.byref(n)))
        return int(n.value)

    def launch_info(self) -> dict[str, int]:
        v = [C.c_int32() for _ in range(5)]
        _capi.check(_capi.lib().dz_batch_launch_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(["grid", "block", "smem_bytes", "ctas_per_sm", "w_in_smem"],
                        [x.value for x in v]))

    def io_bytes(self) -> tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _capi.check(_capi.lib().dz_batch_io_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def download(self, light: bool = False) -> BatchResult:
        """light=True skips x_basic/basis/trace (what a caller needing only
        status, objective and primal values reads)."""
        B, t = self.B, self.template
        res = BatchResult(
            status=np.zeros(B, np.int32), pivots=np.zeros(B, np.int32),
            n_primal=np.zeros(B, np.int32), trace_hash=np.zeros(B, np.uint64),
            objective=np.zeros(B, np.float64), values=np.zeros((B, max(t.n_orig, 1)), np.float64),
            x_basic=None if light else np.zeros((B, t.m), np.float64),
            basis=None if light else np.zeros((B, t.m), np.int32),
            trace=None if (light or not self.trace_cap) else np.zeros((B, self.trace_cap, 3), np.int32),
            work=np.zeros((B, 8), np.float64),
            prof=np.zeros((B, 16), np.int64) if self.profile else None,
        )
        r = _capi.BatchResult()
        for name in ("status", "pivots", "n_primal", "trace_hash", "objective", "values",
                     "x_basic", "basis", "trace", "work", "prof"):
            a = getattr(res, name)
            setattr(r, name, None if a is None else a.ctypes.data)
        _capi.check(_capi.lib().dz_batch_download(self._h, C.byref(r)))
        res.values = res.values[:, : t.n_orig]
        return res


def solve_batch_multi(template: Template, theta: np.ndarray, n_gpus: int = 0, *, max_pivots: int = 0,
                      trace_cap: int = 0, light: bool = False, **kw) -> BatchResult:
    """dz_solve_batch_multi: the batch sharded over the first ``n_gpus`` devices (0 = all),
    contiguous LP ranges, one host thread and stream per device, no collective."""
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, template.n_theta)
    B, t = theta.shape[0], template
    res = BatchResult(
        status=np.zeros(B, np.int32), pivots=np.zeros(B, np.int32),
        n_primal=np.zeros(B, np.int32), trace_hash=np.zeros(B, np.uint64),
        objective=np.zeros(B, np.float64), values=np.zeros((B, max(t.n_orig, 1)), np.float64),
        x_basic=None if light else np.zeros((B, t.m), np.float64),
        basis=None if light else np.zeros((B, t.m), np.int32),
        trace=None if (light or not trace_cap) else np.zeros((B, trace_cap, 3), np.int32),
        work=np.zeros((B, 8), np.float64), prof=None,
    )
    r = _capi.BatchResult()
    for name in ("status", "pivots", "n_primal", "trace_hash", "objective", "values",
                 "x_basic", "basis", "trace", "work", "prof"):
        a = getattr(res, name)
        setattr(r, name, None if a is None else a.ctypes.data)
    o = _options(0, max_pivots, trace_cap, kw.get("worker_warps", 0), kw.get("ctas_per_sm", 0), None, False,
                 kw.get("basis_home", 0), kw.get("numerics", "exact"))
    _capi.check(_capi.lib().dz_solve_batch_multi(t.handle, B, _vp(theta), int(n_gpus), C.byref(o), C.byref(r)))
    res.values = res.values[:, : t.n_orig]
    return res


def solve_batch(template: Template, theta: np.ndarray, **kw) -> BatchResult:
    """The new batched entry point (host buffers in, host buffers out)."""
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, template.n_theta)
    b = Batch(template, theta.shape[0], **kw)
    try:
        b.upload(theta)
        b.solve()
        return b.download()
    finally:
        b.close()


@dataclass
class Solution:
    status: int
    pivots: int
    n_primal: int
    trace_hash: int
    objective: float
    values: np.ndarray  # one per variable of the model (0.0 when never mentioned)

    @property
    def status_name(self) -> str:
        return _capi.STATUS_NAMES[self.status]


def solve_model(model: ModelArrays, *, device: int = 0, max_pivots: int = 0, numerics: str = "exact") -> Solution:
    """Replacement for ``Simplex::new(objective, constraints).solve()``.  ``numerics="fast"`` opts this one
    solve into the fast-numerics kernel (lowered LPs of up to 512 rows; never the default)."""
    cm = model.as_c()
    o = _options(device, max_pivots, numerics=numerics)
    sol = _capi.Solution()
    values = np.zeros(max(model.n_vars, 1), np.float64)
    _capi.check(_capi.lib().dz_solve_model(C.byref(cm), C.byref(o), C.byref(sol), _vp(values)))
    return Solution(sol.status, sol.pivots, sol.n_primal, sol.trace_hash, sol.objective,
                    values[: model.n_vars])


def device_count() -> int:
    return int(_capi.lib().dz_device_count())


def device_info(device: int = 0) -> dict:
    name = C.create_string_buffer(256)
    sm, maj, mnr, smem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    _capi.check(_capi.lib().dz_device_info(device, name, 256, C.byref(sm), C.byref(maj),
                                           C.byref(mnr), C.byref(smem)))
    return dict(name=name.value.decode(), sm_count=sm.value, cc=(maj.value, mnr.value),
                smem_per_sm=smem.value)


def measure_fp64_peak(device: int = 0) -> tuple[float, float]:
    a, b = C.c_double(), C.c_double()
    _capi.check(_capi.lib().dz_measure_fp64_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def solve_dense_batch(A, b, c, senses, lb=None, ub=None, *, minimize: bool = True, **kw):
    """Array-native batched front door: B dense LPs of one shape,

        min / max  c[i] . x   s.t.  A[i] x  (senses)  b[i],   lb <= x <= ub,

    with `senses[r]` in {LE, GE, EQ} and `lb`/`ub` entries of +-inf or None
    meaning "no bound" (shared by the batch).  Lowered exactly the way the
    reference frontend lowers the same model (see model.py), solved on the GPU.
    Returns (status[B], objective[B] in the caller's sense, x[B, n], BatchResult).
    """
    from .model import dense_structure

    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 2:
        A = A[None]
    B, m, n = A.shape
    lb = np.zeros(n) if lb is None else np.asarray([-np.inf if v is None else v for v in lb], float)
    ub = np.full(n, np.inf) if ub is None else np.asarray([np.inf if v is None else v for v in ub], float)
    has_lb, has_ub = np.isfinite(lb), np.isfinite(ub)
    structure = dense_structure(m, n, senses, has_lb, has_ub)
    # the numbers are lowered on the device (dz_batch_pack_dense): no host-side theta
    batch = Batch(Template(structure), B, **kw)
    try:
        batch.pack_dense(A, b, c, senses, np.where(has_lb, lb, 0.0), np.where(has_ub, ub, 0.0), minimize=minimize)
        batch.solve()
        res = batch.download()
    finally:
        batch.close()
    objective = -res.objective if minimize else res.objective
    return res.status, objective, res.values, res
