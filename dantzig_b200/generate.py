"""Deterministic synthetic LPs for the BASELINE.json configs.

Counter-based: every number is splitmix64(key(seed, family, lp, stream) + i)
mapped to [0,1) with 53 bits, so any LP of a batch can be regenerated on its
own, on any machine, bit for bit (the derived quantities use only elementwise
products and numpy's fixed-order reductions, never BLAS).

Families
  general   A~U(-1,1), feasible and bounded by construction (configs 1, 2, 5)
  packing   max c.x, A~U(0,1) x <= b (config 3; numerically benign)
"""
from __future__ import annotations

import numpy as np

from .model import EQ, GE, LE, ModelArrays, dense_structure, dense_theta

BASE_SEED = 1234
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _key(seed: int, family: int, lp: np.ndarray, stream: int) -> np.ndarray:
    lp = np.asarray(lp, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = _mix(np.uint64(seed) ^ (np.uint64(family) * np.uint64(0xD1B54A32D192ED03)))
        k = _mix(k ^ (lp * np.uint64(0x8CB92BA72F3D8DD7)))
        k = _mix(k ^ (np.uint64(stream) * np.uint64(0xAEF17502108EF2D9)))
    return k


def uniform(seed: int, family: int, lp, stream: int, count: int) -> np.ndarray:
    """[len(lp), count] doubles in [0,1)."""
    lp = np.atleast_1d(np.asarray(lp, dtype=np.uint64))
    key = _key(seed, family, lp, stream)
    with np.errstate(over="ignore"):
        ctr = key[:, None] + np.arange(count, dtype=np.uint64)[None, :] * np.uint64(0x9E3779B97F4A7C15)
    bits = _mix(ctr)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def general_lps(lp_ids, m: int, n: int, senses, free_mask, seed: int = BASE_SEED, family: int = 1):
    """Feasible+bounded LPs in MIN form: returns A[B,m,n], b[B,m], c[B,n]."""
    lp_ids = np.atleast_1d(np.asarray(lp_ids, dtype=np.int64))
    B = len(lp_ids)
    senses = np.asarray(senses, dtype=np.int32)
    free_mask = np.asarray(free_mask, dtype=bool)
    A = uniform(seed, family, lp_ids, 0, m * n).reshape(B, m, n) * 2.0 - 1.0
    x0u = uniform(seed, family, lp_ids, 1, n)
    x0 = np.where(free_mask[None, :], x0u * 2.0 - 1.0, x0u)
    s = 0.1 + 0.9 * uniform(seed, family, lp_ids, 2, m)
    yu = uniform(seed, family, lp_ids, 3, m)
    y = np.where(senses[None, :] == LE, -yu, np.where(senses[None, :] == GE, yu, yu * 2.0 - 1.0))
    r = 0.1 + 0.9 * uniform(seed, family, lp_ids, 4, n)
    r = np.where(free_mask[None, :], 0.0, r)
    ax = (A * x0[:, None, :]).sum(axis=2)
    b = np.where(senses[None, :] == LE, ax + s, np.where(senses[None, :] == GE, ax - s, ax))
    c = (A * y[:, :, None]).sum(axis=1) + r
    return A, b, c


def packing_lps(lp_ids, m: int, n: int, seed: int = BASE_SEED, family: int = 3):
    """Dense packing LPs: maximise c.x s.t. A x <= b, x >= 0."""
    lp_ids = np.atleast_1d(np.asarray(lp_ids, dtype=np.int64))
    B = len(lp_ids)
    A = uniform(seed, family, lp_ids, 0, m * n).reshape(B, m, n)
    b = n * (0.2 + 0.1 * uniform(seed, family, lp_ids, 1, m))
    c = 0.5 + uniform(seed, family, lp_ids, 2, n)
    return A, b, c


class Workload:
    """A shape-uniform batch: structure + per-LP parameter vectors."""

    def __init__(self, name: str, structure: ModelArrays, theta: np.ndarray, m: int, n: int,
                 minimize: bool, lp_ids: np.ndarray):
        self.name, self.structure, self.theta = name, structure, theta
        self.m, self.n, self.minimize, self.lp_ids = m, n, minimize, lp_ids

    @property
    def B(self) -> int:
        return self.theta.shape[0]


def _general_workload(name, lp_ids, m, n, senses, free_mask, family) -> Workload:
    A, b, c = general_lps(lp_ids, m, n, senses, free_mask, family=family)
    has_lb = ~np.asarray(free_mask, bool)
    has_ub = np.zeros(n, bool)
    st = dense_structure(m, n, senses, has_lb, has_ub)
    th = dense_theta(A, b, c, senses, np.zeros(n), np.zeros(n), has_lb, has_ub, minimize=True)
    return Workload(name, st, th, m, n, True, np.asarray(lp_ids))


def config1(seeds=(0,)) -> Workload:
    """m=100 x n=200 dense, rows i%3 -> <=,>=,==, variable j free iff j%4==3."""
    m, n = 100, 200
    senses = np.array([[LE, GE, EQ][i % 3] for i in range(m)], np.int32)
    free = np.array([j % 4 == 3 for j in range(n)])
    return _general_workload("c1_100x200_mixed", np.asarray(seeds), m, n, senses, free, family=1)


def small_batch(B: int, m: int, n: int, first: int = 0, family: int = 2) -> Workload:
    """All-<= rows, nonneg variables (the unit of configs 2 and 5)."""
    senses = np.full(m, LE, np.int32)
    free = np.zeros(n, bool)
    ids = np.arange(first, first + B)
    return _general_workload(f"batch_{m}x{n}", ids, m, n, senses, free, family=family)


def config2(B: int = 4096, first: int = 0) -> Workload:
    w = small_batch(B, 32, 64, first, family=2)
    w.name = "c2_batch_32x64"
    return w


def config5(B: int = 262144, first: int = 0) -> Workload:
    w = small_batch(B, 64, 128, first, family=5)
    w.name = "c5_batch_64x128"
    return w


def mixed_batch(B: int, m: int, n: int, first: int = 0, family: int = 7) -> Workload:
    """Mixed ==/<=/>= rows and 25% free variables at small size (parity cases
    that exercise two-row equalities and split free variables)."""
    senses = np.array([[LE, GE, EQ][i % 3] for i in range(m)], np.int32)
    free = np.array([j % 4 == 3 for j in range(n)])
    ids = np.arange(first, first + B)
    return _general_workload(f"mixed_{m}x{n}", ids, m, n, senses, free, family=family)


def packing(B: int, m: int, n: int, first: int = 0) -> Workload:
    ids = np.arange(first, first + B)
    A, b, c = packing_lps(ids, m, n)
    senses = np.full(m, LE, np.int32)
    has_lb, has_ub = np.ones(n, bool), np.zeros(n, bool)
    st = dense_structure(m, n, senses, has_lb, has_ub)
    th = dense_theta(A, b, c, senses, np.zeros(n), np.zeros(n), has_lb, has_ub, minimize=False)
    return Workload(f"packing_{m}x{n}", st, th, m, n, False, ids)


def transportation_model(seed_id: int, n_supply: int, n_demand: int, n_arcs: int, k: int = 1,
                         seed: int = BASE_SEED, family: int = 4) -> ModelArrays:
    """Transportation-style sparse LP of BASELINE configs[3]:
    `n_supply` supply rows (<=), `n_demand` demand rows (>=), one non-negative
    variable per arc; an arc touches `k` supply rows and `k` demand rows with
    coefficient +1 (k=1 is the pure transportation form).  Integer supplies and
    demands built from an integer feasible flow, continuous costs U(1,2).
    Returned in the MAXIMISE/<= form of dz_model (Minimize negates the costs,
    >= rows are negated), term order = arc order within each row."""
    ids = np.array([seed_id], dtype=np.int64)
    us = uniform(seed, family, ids, 0, n_arcs * k)[0]
    ud = uniform(seed, family, ids, 1, n_arcs * k)[0]
    sup = np.minimum((us * n_supply).astype(np.int64), n_supply - 1).reshape(n_arcs, k)
    dem = np.minimum((ud * n_demand).astype(np.int64), n_demand - 1).reshape(n_arcs, k)
    x0 = np.minimum((uniform(seed, family, ids, 2, n_arcs)[0] * 5).astype(np.int64), 4)
    slack = np.minimum((uniform(seed, family, ids, 3, n_supply)[0] * 3).astype(np.int64), 2)
    cost = 1.0 + uniform(seed, family, ids, 4, n_arcs)[0]
    rows: list[list[int]] = [[] for _ in range(n_supply + n_demand)]
    for a in range(n_arcs):
        for s_ in sorted(set(sup[a].tolist())):
            rows[s_].append(a)
        for d_ in sorted(set(dem[a].tolist())):
            rows[n_supply + d_].append(a)
    supply = np.zeros(n_supply)
    demand = np.zeros(n_demand)
    for a in range(n_arcs):
        for s_ in set(sup[a].tolist()):
            supply[s_] += x0[a]
        for d_ in set(dem[a].tolist()):
            demand[d_] += x0[a]
    supply += slack
    row_ptr = [0]
    row_var: list[int] = []
    row_coef: list[float] = []
    rhs: list[float] = []
    for r in range(n_supply + n_demand):
        is_demand = r >= n_supply
        for a in rows[r]:
            row_var.append(a)
            row_coef.append(-1.0 if is_demand else 1.0)     # >= rows are negated
        row_ptr.append(len(row_var))
        rhs.append(-float(demand[r - n_supply]) if is_demand else float(supply[r]))
    return ModelArrays(
        n_vars=n_arcs, has_lb=np.ones(n_arcs, bool), has_ub=np.zeros(n_arcs, bool),
        lb=np.zeros(n_arcs), ub=np.zeros(n_arcs),
        obj_var=np.arange(n_arcs, dtype=np.int32), obj_coef=-cost, obj_const=-0.0,
        row_ptr=row_ptr, row_var=row_var, row_coef=row_coef, rhs=rhs)
