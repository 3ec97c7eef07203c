"""Parity of the CUDA path with the oracle, called through the C ABI.
Bit-exact: status, pivot count, pivot trace (hash and entries), objective bits,
primal values.  Run on the B200 box:  pytest -m gpu"""
import json
import os
import struct

import numpy as np
import pytest

from dantzig_b200 import Batch, Template, generate, solve_batch, solve_model
from dantzig_b200.model import model_from_theta
from tests import cases, kat

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATUS = {"optimal": 0, "unbounded": 1, "infeasible": 2}


def bits(x):
    return struct.pack("<d", float(x)).hex()


def sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ---- the reference's own known-answer tests through dz_solve_model --------------
@pytest.mark.parametrize("name,model,expect", kat.rust_kats(), ids=[k[0] for k in kat.rust_kats()])
def test_rust_kats_on_device(oracle, name, model, expect):
    s = solve_model(model)
    assert s.status == STATUS[expect[0]]
    o = oracle.lower(model).solve(oracle.LITERAL)
    assert (s.pivots, s.trace_hash) == (o.pivots, o.trace_hash)
    if expect[0] == "optimal":
        assert abs(s.objective - expect[1]) <= 1e-12          # src/simplex.rs:477-482
        assert bits(s.objective) == bits(o.objective)
        for v, val in expect[2].items():
            assert abs(s.values[v] - val) <= 1e-12


@pytest.mark.parametrize("name,model,minimize,expect", kat.python_kats(),
                         ids=[k[0] for k in kat.python_kats()])
def test_python_kats_on_device(name, model, minimize, expect):
    s = solve_model(model)
    assert s.status == STATUS[expect[0]]
    if expect[0] == "optimal":                                   # exact ==, tests/test_optimize.py
        assert (-s.objective if minimize else s.objective) == expect[1]
        for v, val in expect[2].items():
            assert s.values[v] == val


@pytest.mark.parametrize("name,model", cases.ragged_models(), ids=[c[0] for c in cases.ragged_models()])
def test_ragged_models_on_device(oracle, name, model):
    s = solve_model(model)
    lo = oracle.lower(model)
    o = lo.solve(oracle.LITERAL)
    assert (s.status, s.pivots, s.trace_hash, bits(s.objective)) == (
        o.status, o.pivots, o.trace_hash, bits(o.objective))
    for k, v in enumerate(lo.orig_var):
        assert bits(s.values[v]) == bits(o.values[k])


@pytest.mark.parametrize("shape", [(10, 10, 40, 1), (20, 20, 120, 1), (20, 20, 120, 3)])
def test_transportation_family_on_device(oracle, shape):
    """Sparse transportation-style LPs (BASELINE configs[3] family at test size):
    CSC pricing over 2..6-nonzero columns, almost every elimination step trivial."""
    from scipy.optimize import linprog

    for seed in range(2):
        model = generate.transportation_model(seed, *shape)
        s = solve_model(model)
        lo = oracle.lower(model)
        o = lo.solve(oracle.LITERAL)
        assert (s.status, s.pivots, s.trace_hash, bits(s.objective)) == (
            o.status, o.pivots, o.trace_hash, bits(o.objective))
        for k, v in enumerate(lo.orig_var):
            assert bits(s.values[v]) == bits(o.values[k])
        # independent optimum (HiGHS) on the same sparse model, max form -> min form
        n_rows, n = model.n_rows, model.n_vars
        A = np.zeros((n_rows, n))
        for r in range(n_rows):
            for t in range(model.row_ptr[r], model.row_ptr[r + 1]):
                A[r, model.row_var[t]] = model.row_coef[t]
        ref = linprog(-model.obj_coef, A_ub=A, b_ub=model.rhs, bounds=[(0, None)] * n, method="highs")
        assert s.status == 0 and ref.status == 0
        assert abs(-s.objective - ref.fun) <= 1e-9 * max(1.0, abs(ref.fun))


def test_random_ragged_models_on_device(oracle):
    """Fuzz: random small models with duplicate terms, explicit zeros, empty rows,
    boxed / free / non-positive variables, integer data (ties everywhere).  Every
    outcome of the reference -- optimal, unbounded, infeasible, breakdown -- must be
    reproduced with the same pivot trace."""
    from dantzig_b200.model import ModelBuilder

    rng = np.random.default_rng(20261018)
    seen = np.zeros(5, int)
    for _ in range(150):
        mb = ModelBuilder()
        nv = int(rng.integers(1, 8))
        vs = [mb.var(lb=[None, 0.0, -1.5][rng.integers(3)], ub=[None, 2.5, 4.0][rng.integers(3)])
              for _ in range(nv)]
        pick = lambda k: [(float(rng.integers(-3, 4)), vs[rng.integers(nv)]) for _ in range(k)]
        mb.maximize(pick(int(rng.integers(0, 5))), float(rng.integers(-2, 3)))
        for _ in range(int(rng.integers(0, 7))):
            [mb.leq, mb.geq, mb.eq][rng.integers(3)](pick(int(rng.integers(0, 5))), float(rng.integers(-4, 5)))
        model = mb.build()
        if len(model.obj_var) + len(model.row_var) == 0:
            continue
        lo = oracle.lower(model)
        o = lo.solve(oracle.LITERAL, max_pivots=2000)
        s = solve_model(model, max_pivots=2000)
        seen[o.status] += 1
        assert (s.status, s.pivots, s.trace_hash) == (o.status, o.pivots, o.trace_hash), model
        assert bits(s.objective) == bits(o.objective) or (s.objective != s.objective and o.objective != o.objective)
        for k, v in enumerate(lo.orig_var):
            assert bits(s.values[v]) == bits(o.values[k]) or (s.values[v] != s.values[v] and o.values[k] != o.values[k])
    assert seen[0] > 20 and seen[1] > 5 and seen[2] > 5      # the fuzz reaches all outcomes


@pytest.mark.parametrize("shape", [(-1, 0), (2, 1), (2, 2), (0, 4), (2, 5, 3)],
                         ids=["warp", "cta-smem", "cta-hbm", "core-on-chip", "grid-3ctas"])
def test_integer_data_batches(oracle, shape):
    """Degenerate batches: small-integer coefficients make ties in every pivot
    search and ratio test, exact zeros in the data, infeasible and unbounded LPs."""
    from dantzig_b200.model import dense_structure, dense_theta

    rng = np.random.default_rng(7)
    B, m, n = 96, 7, 10
    A = rng.integers(-2, 3, size=(B, m, n)).astype(np.float64)
    b = rng.integers(-1, 6, size=(B, m)).astype(np.float64)
    c = rng.integers(-3, 4, size=(B, n)).astype(np.float64)
    senses = np.array([0, 1, 2, 0, 0, 1, 0], np.int32)
    has_lb = np.array([j % 3 != 2 for j in range(n)])
    has_ub = np.array([j % 4 == 1 for j in range(n)])
    lb, ub = np.zeros(n), np.full(n, 3.0)
    st = dense_structure(m, n, senses, has_lb, has_ub)
    th = dense_theta(A, b, c, senses, lb, ub, has_lb, has_ub, minimize=True)
    t = Template(st)
    res = solve_batch(t, th, trace_cap=64, worker_warps=shape[0], basis_home=shape[1], max_pivots=500,
                      ctas_per_sm=shape[2] if len(shape) > 2 else 0)
    hist = np.zeros(5, int)
    for i in range(B):
        o = oracle.lower(model_from_theta(st, th[i])).solve(oracle.LITERAL, max_pivots=500, trace_cap=64)
        hist[o.status] += 1
        assert (res.status[i], res.pivots[i], int(res.trace_hash[i])) == (o.status, o.pivots, o.trace_hash), i
        assert np.array_equal(res.trace[i, : min(o.pivots, 64)], o.trace), i
        assert bits(res.objective[i]) == bits(o.objective) or o.objective != o.objective, i
    assert hist[0] > 0 and hist[1] + hist[2] > 0


def test_array_native_front_door():
    """solve_dense_batch: mixed senses, boxed/free variables, against HiGHS."""
    from scipy.optimize import linprog

    from dantzig_b200 import EQ, GE, LE, solve_dense_batch

    rng = np.random.default_rng(3)
    B, m, n = 16, 6, 9
    A = rng.uniform(-1, 1, (B, m, n))
    x0 = rng.uniform(0.2, 0.8, (B, n))
    senses = np.array([LE, GE, EQ, LE, LE, GE], np.int32)
    slack = np.where(senses == LE, 0.3, np.where(senses == GE, -0.3, 0.0))
    b = np.einsum("bij,bj->bi", A, x0) + slack
    c = rng.uniform(-1, 1, (B, n))
    lb = [0.0] * n
    ub = [1.0] * n
    ub[4] = None                         # x4 unbounded above
    status, obj, x, _ = solve_dense_batch(A, b, c, senses, lb, ub, minimize=True)
    for i in range(B):
        a_ub = np.vstack([A[i][senses == LE], -A[i][senses == GE]])
        b_ub = np.concatenate([b[i][senses == LE], -b[i][senses == GE]])
        ref = linprog(c[i], A_ub=a_ub, b_ub=b_ub, A_eq=A[i][senses == EQ], b_eq=b[i][senses == EQ],
                      bounds=[(l, u) for l, u in zip(lb, ub)], method="highs")
        if ref.status == 0 and status[i] == 0:
            assert abs(obj[i] - ref.fun) <= 1e-9 * max(1.0, abs(ref.fun))
        else:                # HiGHS: 2 infeasible, 3 unbounded.  This family is feasible and bounded by construction.
            assert {0: 0, 2: 2, 3: 1}.get(ref.status, -1) == status[i]


def test_pack_dense_on_device_equals_host_theta():
    """dz_batch_pack_dense (the numbers lowered by a kernel from plain A, b, c) gives the very
    parameter vectors the host-side packing gives: same results bit for bit, mixed row senses,
    boxed / free variables, both objective senses."""
    from dantzig_b200 import EQ, GE, LE, dense_structure, dense_theta

    rng = np.random.default_rng(11)
    B, m, n = 24, 7, 10
    A = rng.uniform(-1, 1, (B, m, n))
    x0 = rng.uniform(0.2, 0.8, (B, n))
    senses = np.array([LE, GE, EQ, LE, GE, LE, EQ], np.int32)
    b = np.einsum("bij,bj->bi", A, x0) + np.where(senses == LE, 0.3, np.where(senses == GE, -0.3, 0.0))
    c = rng.uniform(-1, 1, (B, n))
    has_lb = np.array([j % 4 != 3 for j in range(n)])
    has_ub = np.array([j % 3 == 0 for j in range(n)])
    lb, ub = np.where(has_lb, 0.0, 0.0), np.where(has_ub, 1.0, 0.0)
    st = dense_structure(m, n, senses, has_lb, has_ub)
    t = Template(st)
    for minimize in (True, False):
        ref = solve_batch(t, dense_theta(A, b, c, senses, lb, ub, has_lb, has_ub, minimize=minimize), trace_cap=32)
        bt = Batch(t, B, trace_cap=32)
        bt.pack_dense(A, b, c, senses, lb, ub, minimize=minimize)
        bt.solve()
        got = bt.download()
        bt.close()
        for name in ("status", "pivots", "trace_hash", "objective", "values", "x_basic", "basis", "trace"):
            a1, a2 = getattr(ref, name), getattr(got, name)
            assert np.array_equal(a1, a2, equal_nan=a1.dtype.kind == "f"), (minimize, name)
    with pytest.raises(Exception):                       # a template of another shape is refused
        bt = Batch(Template(dense_structure(m, n + 1, senses, np.ones(n + 1, bool), np.zeros(n + 1, bool))), B)
        try:
            bt.pack_dense(A, b, c, senses, lb, ub)
        finally:
            bt.close()


def test_empty_basis_is_breakdown():
    from dantzig_b200.model import ModelBuilder

    mb = ModelBuilder()
    x = mb.free()
    mb.maximize([(1.0, x)])
    assert solve_model(mb.build()).status == 3                   # the reference panics


# ---- batched kernel against the live oracle and the committed fixtures ----------
def _check_batch(oracle, w, res, n_oracle, variant):
    for i in range(min(n_oracle, w.B)):
        o = oracle.lower(model_from_theta(w.structure, w.theta[i])).solve(variant, trace_cap=res.trace.shape[1])
        assert (res.status[i], res.pivots[i], res.n_primal[i], int(res.trace_hash[i])) == (
            o.status, o.pivots, o.n_primal, o.trace_hash), i
        assert np.array_equal(res.trace[i, : min(o.pivots, res.trace.shape[1])], o.trace), i
        assert bits(res.objective[i]) == bits(o.objective), i
        assert np.array_equal(res.values[i].view(np.uint64), o.values.view(np.uint64)), i
        assert np.array_equal(res.x_basic[i].view(np.uint64), o.x_basic.view(np.uint64)), i
        assert np.array_equal(res.basis[i], o.basis), i


@pytest.mark.parametrize("wl", sorted(cases.GOLDEN_WORKLOADS))
@pytest.mark.parametrize("shape", [(0, 0), (-1, 0), (1, 2), (3, 1), (2, 3), (0, 4)],
                         ids=["auto", "warp-per-lp", "cta-1warp-hbm", "cta-3warps-smem", "cta-all-hbm", "core-on-chip"])
def test_batch_parity(oracle, wl, shape):
    """Every launch shape (worker warps per LP, home of the working basis) must
    give the same bits: they only change which thread does which operation."""
    w = cases.GOLDEN_WORKLOADS[wl]()
    if w.m >= 100 and shape not in ((0, 0), (2, 3)):
        pytest.skip("config-1 size: auto and all-in-HBM shapes only (minutes on one warp)")
    if w.m >= 64 and shape == (1, 2):
        pytest.skip("config-5 size on one worker warp per CTA: minutes")
    t = Template(w.structure)
    res = solve_batch(t, w.theta, trace_cap=256, worker_warps=shape[0], basis_home=shape[1])
    g = json.load(open(os.path.join(GOLD, wl + ".json")))
    assert sha(w.theta) == g["theta_sha"]
    for i, e in enumerate(g["lps"]):                             # every LP against the fixture
        assert (res.status[i], res.pivots[i], res.n_primal[i], int(res.trace_hash[i])) == (
            e["status"], e["pivots"], e["n_primal"], e["trace_hash"]), i
        assert bits(res.objective[i]) == e["objective_bits"], i
        assert sha(res.values[i]) == e["values_sha"], i
    big = w.m >= 60
    _check_batch(oracle, w, res, (1 if w.m >= 100 else 2) if big else 12,
                 oracle.SKIP if big else oracle.LITERAL)


def test_batch_order_and_resolve_invariance():
    """Solving is a pure function of each LP: permuting the batch permutes the
    results, re-solving a resident batch reproduces them bit for bit."""
    w = generate.config2(96)
    t = Template(w.structure)
    b = Batch(t, w.B)
    b.upload(w.theta)
    b.solve()
    r1 = b.download()
    b.solve()
    r2 = b.download()
    perm = np.random.default_rng(0).permutation(w.B)
    r3 = solve_batch(t, w.theta[perm])
    for name in ("status", "pivots", "trace_hash", "objective", "values", "x_basic", "basis"):
        a1, a2, a3 = getattr(r1, name), getattr(r2, name), getattr(r3, name)
        assert np.array_equal(a1, a2, equal_nan=a1.dtype.kind == "f"), name
        assert np.array_equal(a1[perm], a3, equal_nan=a1.dtype.kind == "f"), name
    b.close()


def test_full_config2_properties(oracle):
    """BASELINE.json configs[1] at full size: 4096 LPs.  Size-independent checks:
    feasibility and complementary objective agreement via HiGHS on a sample,
    a checksum of trace hashes stable across launch shapes, oracle parity on a
    strided sample, and the three known false-unbounded LPs."""
    from scipy.optimize import linprog

    w = generate.config2(4096)
    t = Template(w.structure)
    r = solve_batch(t, w.theta)
    r2 = solve_batch(t, w.theta, worker_warps=4, ctas_per_sm=1, basis_home=1)
    assert np.array_equal(r.trace_hash, r2.trace_hash) and np.array_equal(r.objective, r2.objective)
    assert sorted(np.flatnonzero(r.status != 0).tolist()) == [287, 2142, 3300]
    assert (r.status[[287, 2142, 3300]] == 1).all()
    m, n = w.m, w.n
    for i in range(0, 4096, 512):
        th = w.theta[i]
        o = oracle.lower(model_from_theta(w.structure, th)).solve(oracle.SKIP)
        assert (r.status[i], r.pivots[i], int(r.trace_hash[i]), bits(r.objective[i])) == (
            o.status, o.pivots, o.trace_hash, bits(o.objective))
        c = -th[2:2 + n]
        A = th[2 + n:2 + n + m * n].reshape(m, n)
        b = th[2 + n + m * n:2 + n + m * n + m]
        x = r.values[i]
        assert (x >= -1e-9).all() and (A @ x <= b + 1e-7).all()          # primal feasible
        ref = linprog(c, A_ub=A, b_ub=b, bounds=[(0, None)] * n, method="highs")
        assert abs(-r.objective[i] - ref.fun) <= 1e-9 * max(1.0, abs(ref.fun))
        assert np.abs(x - ref.x).max() <= 1e-7


def test_solve_batch_multi(oracle):
    """dz_solve_batch_multi: contiguous shards over the devices, no collective; the results
    are those of the single-device entry, LP for LP.  With one visible device the sharding
    logic still runs (n_gpus=1); with two or more the batch really splits."""
    from dantzig_b200 import device_count, solve_batch_multi

    w = generate.config2(203)                       # not a multiple of anything
    t = Template(w.structure)
    one = solve_batch(t, w.theta, trace_cap=16)
    for n in sorted({1, min(2, device_count()), device_count()}):
        r = solve_batch_multi(t, w.theta, n_gpus=n, trace_cap=16)
        for name in ("status", "pivots", "n_primal", "trace_hash", "objective", "values", "x_basic", "basis", "trace"):
            a, b = getattr(one, name), getattr(r, name)
            assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), (n, name)
    with pytest.raises(Exception):
        solve_batch_multi(t, w.theta, n_gpus=device_count() + 1)


def test_pivot_cap_is_reported():
    w = generate.config2(4)
    t = Template(w.structure)
    r = solve_batch(t, w.theta, max_pivots=10)
    assert (r.status == 4).all() and (r.pivots == 10).all()


# ---- the drop-in extension module ----------------------------------------------------
def test_rust_module_end_to_end():
    import sys
    import types

    import dantzig_b200.rust as rs

    exc = types.ModuleType("dantzig.exceptions")
    exc.UnboundedError = type("UnboundedError", (Exception,), {})
    exc.InfeasibleError = type("InfeasibleError", (Exception,), {})
    had = {k: sys.modules.get(k) for k in ("dantzig", "dantzig.exceptions", "dantzig.rust")}
    try:
        if "dantzig.exceptions" not in sys.modules:
            pkg = types.ModuleType("dantzig")
            pkg.exceptions = exc
            sys.modules["dantzig"], sys.modules["dantzig.exceptions"] = pkg, exc
        sys.modules["dantzig.rust"] = rs
        exc_mod = sys.modules["dantzig.exceptions"]
        x, y = rs.Variable(lb=0.0, ub=None), rs.Variable(lb=0.0, ub=None)
        # tests/test_optimize.py:4-11  min 2x-2y st y == 3  (negated for the max-form solver)
        obj = rs.PyAffExpr(linexpr=rs.PyLinExpr([-2.0, 2.0], [x, y]), constant=-0.0)
        rows = [rs.PyInequality(linexpr=rs.PyLinExpr([1.0], [y]), b=3.0),
                rs.PyInequality(linexpr=rs.PyLinExpr([-1.0], [y]), b=-3.0)]
        s = rs.solve(obj, rows)
        assert -s.objective_value == -6.0 and s[x] == 0.0 and s[y] == 3.0
        assert s[rs.Variable(lb=None, ub=None)] == 0.0               # unknown variable
        with pytest.raises(exc_mod.UnboundedError, match="The objective is unbounded"):
            rs.solve(rs.PyAffExpr(linexpr=rs.PyLinExpr([1.0], [x]), constant=0.0), [])
        with pytest.raises(exc_mod.InfeasibleError, match="The model is infeasible"):
            le = rs.PyLinExpr([1.0, 1.0], [x, y])
            rs.solve(rs.PyAffExpr(linexpr=le, constant=0.0),
                     [rs.PyInequality(linexpr=le, b=1.0), rs.PyInequality(linexpr=-le, b=-1.0),
                      rs.PyInequality(linexpr=le, b=2.0), rs.PyInequality(linexpr=-le, b=-2.0)])
    finally:
        for k, v in had.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _frontend(rs):
    import sys
    import types

    if "dantzig.exceptions" not in sys.modules:
        exc = types.ModuleType("dantzig.exceptions")
        exc.UnboundedError = type("UnboundedError", (Exception,), {})
        exc.InfeasibleError = type("InfeasibleError", (Exception,), {})
        pkg = types.ModuleType("dantzig")
        pkg.exceptions = exc
        sys.modules["dantzig"], sys.modules["dantzig.exceptions"] = pkg, exc
    sys.modules.setdefault("dantzig.rust", rs)
    from tests.frontend_shim import bind

    return bind(rs) + (sys.modules["dantzig.exceptions"],)


def test_rust_module_solve_batch():
    """dantzig.rust.solve_batch: mixed structures and outcomes in one call; every
    entry equals what solve() returns (or raises) for the same LP, bit for bit."""
    import dantzig_b200.rust as rs

    _, _, _, exc_mod = _frontend(rs)
    rng = np.random.default_rng(7)
    objs, rows = [], []
    for i in range(24):                                  # three structures, 8 LPs each
        m, n = [(3, 4), (5, 6), (4, 8)][i % 3]
        xs = [rs.Variable(lb=0.0, ub=None) for _ in range(n)]
        A = rng.uniform(0.1, 1.0, (m, n))
        objs.append(rs.PyAffExpr(linexpr=rs.PyLinExpr(rng.uniform(0.5, 1.5, n).tolist(), xs), constant=float(i)))
        rows.append([rs.PyInequality(linexpr=rs.PyLinExpr(A[r].tolist(), xs), b=float(n) * 0.25) for r in range(m)])
    x, y = rs.Variable(lb=0.0, ub=None), rs.Variable(lb=0.0, ub=None)
    le = rs.PyLinExpr([1.0, 1.0], [x, y])
    objs.append(rs.PyAffExpr(linexpr=rs.PyLinExpr([1.0], [x]), constant=0.0))       # unbounded
    rows.append([])
    objs.append(rs.PyAffExpr(linexpr=le, constant=0.0))                            # infeasible
    rows.append([rs.PyInequality(linexpr=le, b=1.0), rs.PyInequality(linexpr=-le, b=-1.0),
                 rs.PyInequality(linexpr=le, b=2.0), rs.PyInequality(linexpr=-le, b=-2.0)])
    out = rs.solve_batch(objs, rows)
    assert len(out) == len(objs)
    assert isinstance(out[-2], exc_mod.UnboundedError) and str(out[-2]) == "The objective is unbounded"
    assert isinstance(out[-1], exc_mod.InfeasibleError) and str(out[-1]) == "The model is infeasible"
    for i in range(24):
        one = rs.solve(objs[i], rows[i])
        assert isinstance(out[i], rs.PySolution)
        assert (out[i].pivots, out[i].trace_hash) == (one.pivots, one.trace_hash)
        assert bits(out[i].objective_value) == bits(one.objective_value)
    assert rs.solve_batch([], []) == []


def _real_frontend(rs):
    """The reference's UNMODIFIED Python frontend (installed by __graft_entry__.build() under
    baseline/_ref, or the checkout) on top of the drop-in module, wrapped to the same four
    names the stand-in exposes."""
    import importlib
    import sys

    from tests.conftest import reference_frontend_dir

    ref = reference_frontend_dir()
    if ref is None:
        pytest.skip("reference frontend not present")
    for k in [k for k in sys.modules if k == "dantzig" or k.startswith("dantzig.")]:
        sys.modules.pop(k)
    sys.modules["dantzig.rust"] = rs
    sys.path.insert(0, ref)
    try:
        dz = importlib.import_module("dantzig")
    finally:
        sys.path.remove(ref)
    assert os.path.dirname(dz.__file__).startswith(ref)
    minimize = lambda obj, cs=(): dz.Minimize(obj).subject_to(list(cs)).solve()
    maximize = lambda obj, cs=(): dz.Maximize(obj).subject_to(list(cs)).solve()
    return dz.Variable, minimize, maximize, dz.exceptions


@pytest.mark.parametrize("frontend", ["stand-in", "reference"])
def test_reference_python_tests_through_the_module(frontend):
    """The reference's tests/test_optimize.py and tests/test_exceptions.py, solved on the GPU
    through the drop-in module, asserted with exact ==: once with the reference's own
    unmodified frontend (dz.Minimize(...).subject_to(...).solve()), once with the stand-in
    that makes the same call sequence (tests/frontend_shim.py)."""
    import sys

    import dantzig_b200.rust as rs

    saved = {k: v for k, v in sys.modules.items() if k == "dantzig" or k.startswith("dantzig.")}
    try:
        _run_reference_cases(*(_real_frontend(rs) if frontend == "reference" else _frontend(rs)))
    finally:
        for k in [k for k in sys.modules if k == "dantzig" or k.startswith("dantzig.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)


def _run_reference_cases(Var, minimize, maximize, exc):
    x, y = Var.nonneg(), Var.nonneg()                                    # test_problem_1
    s = minimize(2 * x - 2 * y, [y == 3])
    assert (s.objective_value, s[x], s[y]) == (-6.0, 0.0, 3.0)
    x, y = Var.nonneg(), Var.nonneg()                                    # test_problem_2
    s = minimize(2 * x - 2 * y, [y <= 5, x >= y + 1, y == 5.0])
    assert (s.objective_value, s[x], s[y]) == (2.0, 6.0, 5.0)
    x, y, z = Var.nonneg(), Var.nonneg(), Var.nonneg()                   # test_problem_3
    s = minimize(x + y - z, [x + y + z <= 1])
    assert (s.objective_value, s[x], s[y], s[z]) == (-1.0, 0.0, 0.0, 1.0)
    x, y, z = Var.nonneg(), Var.nonneg(), Var.nonneg()                   # test_problem_4
    s = minimize(x + y + z, [x - y == -2])
    assert (s.objective_value, s[x], s[y], s[z]) == (2.0, 0.0, 2.0, 0.0)
    x, y = Var.nonneg(), Var.nonneg()                                    # min/max equivalence
    a, b = minimize(-x, [x + y <= 1]), maximize(x, [x + y <= 1])
    assert a.objective_value == -1.0 == -b.objective_value and a[x] == 1.0 == b[x] and a[y] == 0.0 == b[y]
    x, y, z = Var(lb=-2.0, ub=2.0), Var.free(), Var.nonpos()             # non standard variables
    s = minimize(x + y + z, [y == 4, x <= 3.0, z >= -1])
    assert (s.objective_value, s[x], s[y], s[z]) == (1.0, -2.0, 4.0, -1.0)
    p, h, d = [0.5, 3.5, 5.0], [1.0, 5.5, 1.5], [50, 75, 100]           # inventory balance
    xs, zs = [Var.nonneg() for _ in range(3)], [Var.nonneg() for _ in range(3)]
    cost = sum(pt * xt for pt, xt in zip(p, xs)) + sum(ht * zt for ht, zt in zip(h, zs))
    s = minimize(cost, [xs[0] >= d[0], xs[1] + zs[0] >= d[1], xs[2] + zs[1] >= d[2],
                        zs[0] == xs[0] - d[0], zs[1] == xs[1] + zs[0] - d[1],
                        zs[2] == xs[2] + zs[1] - d[2]])
    assert (s.objective_value, s[xs[0]], s[xs[1]], s[xs[2]]) == (637.5, 125.0, 0.0, 100.0)
    x = Var.nonneg()                                                     # test_exceptions.py
    with pytest.raises(exc.UnboundedError):
        minimize(-1.0 * x)
    x, y = Var.nonneg(), Var.nonneg()
    with pytest.raises(exc.InfeasibleError):
        minimize(x + y, [x + y == 1, x + y == 2])
