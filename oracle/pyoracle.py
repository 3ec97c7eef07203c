"""Pure-Python twin of the CPU oracle (TEST INFRASTRUCTURE ONLY).

A second, independent restatement of /root/reference/src/simplex.rs and
src/linalg.rs in plain Python floats (IEEE binary64, never fused), written
against the reference text rather than against dzo.cpp.  It exists to
cross-check the C++ oracle bit for bit on small cases
(tests/test_oracle_variants.py); it is far too slow for anything else.
"""
from __future__ import annotations

import math


def lu_factorize(a: list[list[float]]) -> list[int]:
    """Matrix::factorize, linalg.rs:88-128 (in place; returns p)."""
    n = len(a)
    p = []
    for k in range(n - 1):
        mu, mag = k, abs(a[k][k])
        for i in range(k + 1, n):
            if abs(a[i][k]) > mag:
                mu, mag = i, abs(a[i][k])
        for j in range(k, n):
            a[mu][j], a[k][j] = a[k][j], a[mu][j]
        p.append(mu)
        pivot = a[k][k]
        if pivot != 0.0:
            for i in range(k + 1, n):
                a[i][k] = a[i][k] / pivot
                for j in range(k + 1, n):
                    a[i][j] = a[i][j] - a[i][k] * a[k][j]
    return p


def _div(x: float, y: float) -> float:
    """IEEE division including the cases Python raises on."""
    try:
        return x / y
    except ZeroDivisionError:
        if x != x or x == 0.0:
            return math.nan
        neg = (math.copysign(1.0, x) < 0) != (math.copysign(1.0, y) < 0)
        return -math.inf if neg else math.inf


def lu_solve(a: list[list[float]], b: list[float]) -> list[float]:
    """lu_solve, linalg.rs:8-10 + LU::solve :282-299."""
    n = len(b)
    p = lu_factorize(a)
    for k in range(n - 1):
        b[k], b[p[k]] = b[p[k]], b[k]
        for i in range(k + 1, n):
            b[i] = b[i] - b[k] * a[i][k]
    for i in range(n - 1, -1, -1):
        for j in range(i + 1, n):
            b[i] = b[i] - a[i][j] * b[j]
        b[i] = _div(b[i], a[i][i])
    return b


class Lowered:
    pass


def lower(model) -> Lowered:
    """Simplex::new, simplex.rs:123-224, with explicit id bookkeeping."""
    next_id = [int(model.n_vars)]

    def fresh():
        next_id[0] += 1
        return next_id[0] - 1

    key: dict[int, tuple[int, int]] = {}
    order: list[int] = []
    extra = []
    obj_terms = list(zip([float(c) for c in model.obj_coef], [int(v) for v in model.obj_var]))
    rows = []
    rp = [int(x) for x in model.row_ptr]
    for r in range(len(model.rhs)):
        rows.append(([(float(model.row_coef[t]), int(model.row_var[t])) for t in range(rp[r], rp[r + 1])],
                     float(model.rhs[r])))
    for _, v in obj_terms + [t for row, _ in rows for t in row]:
        if v not in key:
            pos, neg = fresh(), fresh()
            if model.has_ub[v]:
                extra.append(([(1.0, pos), (-1.0, neg)], float(model.ub[v])))
            if model.has_lb[v]:
                extra.append(([(-1.0, pos), (1.0, neg)], -float(model.lb[v])))
            key[v] = (pos, neg)
            order.append(v)

    def split(terms):
        out = []
        for c, v in terms:
            out.append((c, key[v][0]))
            out.append((-c, key[v][1]))
        return out

    objective = split(obj_terms)
    eqs = [(split(t), b) for t, b in rows] + extra
    slack_b = {}
    full = []
    for terms, b in eqs:
        s = fresh()
        full.append((terms + [(1.0, s)], b))
        slack_b[s] = b
    index_of: dict[int, int] = {}
    ids = []
    for _, i in objective + [t for terms, _ in full for t in terms]:
        if i not in index_of:
            index_of[i] = len(ids)
            ids.append(i)
    n_int, m = len(ids), len(full)
    c = [0.0] * n_int
    for coef, i in objective:
        c[index_of[i]] = coef
    lo = Lowered()
    lo.m, lo.n_int, lo.c, lo.c0 = m, n_int, c, float(model.obj_const)
    lo.basis, lo.nonbasis, lo.b = [], [], []
    for i in ids:
        if i in slack_b:
            lo.basis.append(index_of[i])
            lo.b.append(slack_b[i])
        else:
            lo.nonbasis.append(index_of[i])
    dense = {}
    for r, (terms, _) in enumerate(full):
        for coef, i in terms:
            dense[(r, index_of[i])] = coef
    lo.cols = [[] for _ in range(n_int)]
    for (r, j) in sorted(dense, key=lambda rj: (rj[1], rj[0])):
        if dense[(r, j)] != 0.0:
            lo.cols[j].append((r, dense[(r, j)]))
    lo.orig = order
    lo.pos = [index_of[key[v][0]] for v in order]
    lo.neg = [index_of[key[v][1]] for v in order]
    return lo


def _first(y, yb):
    best, ratio = -1, 0.0
    for k in range(len(y)):
        if yb[k] > 0.0:
            r = _div(-y[k], yb[k])
            if best < 0 or r > ratio:
                best, ratio = k, r
    return best


def _second(mu, y, yb, dy):
    best, ratio = -1, 0.0
    for k in range(len(y)):
        r = _div(dy[k], y[k] + mu * yb[k])
        if r > 0.0 and (best < 0 or r > ratio):
            best, ratio = k, r
    return best


def _safe_divide(x, y):
    d = 0.0 if (x == 0.0 and y == 0.0) else _div(x, y)
    if math.isinf(d) or math.isnan(d):
        raise ArithmeticError("safe divide")
    return d


def solve(lo: Lowered, max_pivots: int = 0):
    """Simplex::solve, simplex.rs:274-343.  Returns (status, pivots, trace, objective, x, basis)."""
    M = lo.m
    b, n = list(lo.basis), list(lo.nonbasis)
    x, z = list(lo.b), [-lo.c[j] for j in lo.nonbasis]
    xb, zb = [1.0] * M, [1.0] * len(n)
    trace = []
    status = 0

    def basis_dense(transposed):
        a = [[0.0] * M for _ in range(M)]
        for p, col in enumerate(b):
            for r, v in lo.cols[col]:
                if transposed:
                    a[p][r] = v
                else:
                    a[r][p] = v
        return a

    def dx_for(j):
        rhs = [0.0] * M
        for r, v in lo.cols[j]:
            rhs[r] = v
        return lu_solve(basis_dense(False), rhs)

    def dz_for(p):
        e = [0.0] * M
        e[p] = 1.0
        v = lu_solve(basis_dense(True), e)
        out = []
        for col in n:
            s = 0.0
            for r, a in lo.cols[col]:
                s += a * -v[r]
            out.append(s)
        return out

    while True:
        if M == 0:
            status = 3
            break
        q0, p0 = _first(z, zb), _first(x, xb)
        if q0 >= 0 and p0 >= 0:
            primal, dual = _div(-x[p0], xb[p0]), _div(-z[q0], zb[q0])
            if primal <= 1e-12 and dual <= 1e-12:
                break
            primal_step, mu = (True, dual) if primal < dual else (False, primal)
        elif q0 >= 0:
            primal_step, mu = True, _div(-z[q0], zb[q0])
        elif p0 >= 0:
            primal_step, mu = False, _div(-x[p0], xb[p0])
        else:
            status = 3
            break
        if max_pivots and len(trace) >= max_pivots:
            status = 4
            break
        if primal_step:
            q = q0
            dx = dx_for(n[q])
            p = _second(mu, x, xb, dx)
            if p < 0:
                status = 1
                break
            dz = dz_for(p)
        else:
            p = p0
            dz = dz_for(p)
            q = _second(mu, z, zb, dz)
            if q < 0:
                status = 2
                break
            dx = dx_for(n[q])
        try:
            t, s = _safe_divide(x[p], dx[p]), _safe_divide(z[q], dz[q])
            tb, sb = _safe_divide(xb[p], dx[p]), _safe_divide(zb[q], dz[q])
        except ArithmeticError:
            status = 3
            break
        for vec, d, idx, step in ((x, dx, p, t), (xb, dx, p, tb), (z, dz, q, s), (zb, dz, q, sb)):
            for k in range(len(vec)):
                vec[k] = step if k == idx else vec[k] - step * d[k]
        trace.append((0 if primal_step else 1, b[p], n[q]))
        b[p], n[q] = n[q], b[p]
    obj = 0.0
    for p in range(M):
        obj += lo.c[b[p]] * x[p]
    return status, len(trace), trace, lo.c0 + obj, x, b
