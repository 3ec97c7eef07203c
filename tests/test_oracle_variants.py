"""Oracle self-consistency: LITERAL vs SKIP variants bit-identical, C++ vs the
pure-Python twin bit-identical, and the committed golden fixtures.  CPU only."""
import json
import os
import struct

import numpy as np
import pytest

from dantzig_b200.model import model_from_theta
from tests import cases, kat

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(x):
    return struct.pack("<d", float(x)).hex()


@pytest.mark.parametrize("wl", ["tiny_4x6", "small_8x16", "mixed_9x12", "mixed_20x40", "c2_32x64",
                                "c2_false_unbounded"])
def test_literal_equals_skip(oracle, wl):
    w = cases.GOLDEN_WORKLOADS[wl]()
    for i in range(min(w.B, 12)):
        lo = oracle.lower(model_from_theta(w.structure, w.theta[i]))
        a, b = lo.solve(oracle.LITERAL, trace_cap=4096), lo.solve(oracle.SKIP, trace_cap=4096)
        assert (a.status, a.pivots, a.trace_hash) == (b.status, b.pivots, b.trace_hash)
        assert np.array_equal(a.trace, b.trace)
        assert bits(a.objective) == bits(b.objective)
        assert np.array_equal(a.x_basic, b.x_basic, equal_nan=True)
        assert b.flops[0] <= a.flops[0]


@pytest.mark.parametrize("wl", ["tiny_4x6", "small_8x16", "mixed_9x12", "mixed_20x40", "c2_32x64",
                                "packing_24x48", "c2_false_unbounded"])
def test_sparse_equals_skip(oracle, wl):
    """The sparse-row variant (the only one that can follow config 4 at full size) does the
    SKIP variant's operations in the SKIP variant's order."""
    w = cases.GOLDEN_WORKLOADS[wl]()
    for i in range(min(w.B, 6)):
        lo = oracle.lower(model_from_theta(w.structure, w.theta[i]))
        a, b = lo.solve(oracle.SKIP, trace_cap=4096), lo.solve(oracle.SPARSE, trace_cap=4096)
        assert (a.status, a.pivots, a.trace_hash) == (b.status, b.pivots, b.trace_hash)
        assert bits(a.objective) == bits(b.objective)
        assert np.array_equal(a.x_basic, b.x_basic, equal_nan=True)
        assert np.array_equal(a.values, b.values, equal_nan=True)


def test_sparse_equals_skip_on_breakdown_and_sparse_families(oracle):
    from dantzig_b200 import generate

    from dantzig_b200.model import ModelBuilder

    rng = np.random.default_rng(20261018)      # small integer models: ties, zero pivots, every outcome
    seen = np.zeros(5, int)
    for _ in range(300):
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
        a, b = lo.solve(oracle.SKIP, max_pivots=2000), lo.solve(oracle.SPARSE, max_pivots=2000)
        seen[a.status] += 1
        assert (a.status, a.pivots, a.trace_hash) == (b.status, b.pivots, b.trace_hash)
        assert np.array_equal(a.x_basic, b.x_basic, equal_nan=True)
    assert seen[:4].min() > 0
    for seed, shape in ((0, (10, 10, 40, 1)), (1, (20, 20, 120, 3)), (2, (60, 60, 150, 5))):
        lo = oracle.lower(generate.transportation_model(seed, *shape))
        a, b = lo.solve(oracle.SKIP), lo.solve(oracle.SPARSE)
        assert (a.status, a.pivots, a.trace_hash, bits(a.objective)) == (
            b.status, b.pivots, b.trace_hash, bits(b.objective))
        assert np.array_equal(a.x_basic, b.x_basic, equal_nan=True)
    # config 4 at 1/8 scale, 40-pivot prefix: the dense SKIP variant needs ~20 s, the sparse one 0.1 s
    lo = oracle.lower(generate.transportation_model(0, 1000, 1000, 2500, 10))
    a, b = lo.solve(oracle.SKIP, max_pivots=12), lo.solve(oracle.SPARSE, max_pivots=12)
    assert (a.status, a.pivots, a.trace_hash, bits(a.objective)) == (4, 12, b.trace_hash, bits(b.objective))


def test_skip_variant_on_breakdown_cases(oracle):
    """Statuses other than optimal must agree too (non-finite values disable skipping)."""
    w = cases.GOLDEN_WORKLOADS["mixed_80x160_breakdown"]()
    g = json.load(open(os.path.join(GOLD, "mixed_80x160_breakdown.json")))
    for i in range(w.B):
        r = oracle.lower(model_from_theta(w.structure, w.theta[i])).solve(oracle.SKIP)
        e = g["lps"][i]
        assert (r.status, r.pivots, r.trace_hash) == (e["status"], e["pivots"], e["trace_hash"])


@pytest.mark.parametrize("name,model,expect", kat.rust_kats(), ids=[k[0] for k in kat.rust_kats()])
def test_python_twin_on_kats(oracle, name, model, expect):
    from oracle import pyoracle

    st, piv, trace, obj, x, basis = pyoracle.solve(pyoracle.lower(model))
    r = oracle.lower(model).solve(oracle.LITERAL, trace_cap=64)
    assert (st, piv) == (r.status, r.pivots)
    assert [tuple(t) for t in r.trace] == trace
    assert bits(obj) == bits(r.objective)
    assert list(basis) == list(r.basis) and [bits(v) for v in x] == [bits(v) for v in r.x_basic]


@pytest.mark.parametrize("wl", ["tiny_4x6", "mixed_9x12"])
def test_python_twin_on_random(oracle, wl):
    from oracle import pyoracle

    w = cases.GOLDEN_WORKLOADS[wl]()
    for i in range(6):
        model = model_from_theta(w.structure, w.theta[i])
        st, piv, trace, obj, x, basis = pyoracle.solve(pyoracle.lower(model))
        r = oracle.lower(model).solve(oracle.LITERAL, trace_cap=256)
        assert (st, piv, bits(obj)) == (r.status, r.pivots, bits(r.objective))
        assert [tuple(t) for t in r.trace] == trace


def test_python_twin_linalg_kats():
    from oracle import pyoracle

    a = [[3.0, 17.0, 10.0], [2.0, 4.0, -2.0], [6.0, 18.0, -12.0]]  # linalg.rs:323-345
    assert pyoracle.lu_factorize(a) == [2, 2]
    assert a == [[6.0, 18.0, -12.0], [1.0 / 3.0, 8.0, 16.0], [0.5, -0.25, 6.0]]
    assert pyoracle.lu_solve([[6.0, 18.0, 3.0], [2.0, 12.0, 1.0], [4.0, 15.0, 3.0]],
                             [3.0, 19.0, 0.0]) == [-3.0, 3.0, -11.0]      # linalg.rs:361-369


@pytest.mark.parametrize("wl", sorted(cases.GOLDEN_WORKLOADS))
def test_golden_fixtures(oracle, wl):
    """Generator and oracle reproduce the committed fixtures (guards against
    drift between this container and the GPU box)."""
    import hashlib

    g = json.load(open(os.path.join(GOLD, wl + ".json")))
    w = cases.GOLDEN_WORKLOADS[wl]()
    assert hashlib.sha256(np.ascontiguousarray(w.theta).tobytes()).hexdigest()[:16] == g["theta_sha"]
    big = w.m >= 60
    for i in range(w.B if not big else min(w.B, 3)):
        r = oracle.lower(model_from_theta(w.structure, w.theta[i])).solve(oracle.SKIP)
        e = g["lps"][i]
        assert (r.status, r.pivots, r.n_primal, r.trace_hash) == (
            e["status"], e["pivots"], e["n_primal"], e["trace_hash"])
        assert bits(r.objective) == e["objective_bits"]


def test_highs_cross_check(oracle):
    """Independent check of the optimum where the oracle says Optimal."""
    from scipy.optimize import linprog

    w = cases.GOLDEN_WORKLOADS["c2_32x64"]()
    for i in range(4):
        th = w.theta[i]
        m, n = w.m, w.n
        c = -th[2:2 + n]
        A = th[2 + n:2 + n + m * n].reshape(m, n)
        b = th[2 + n + m * n:2 + n + m * n + m]
        ref = linprog(c, A_ub=A, b_ub=b, bounds=[(0, None)] * n, method="highs")
        r = oracle.lower(model_from_theta(w.structure, th)).solve(oracle.SKIP)
        assert ref.status == 0 and r.status == 0
        assert abs(-r.objective - ref.fun) <= 1e-9 * max(1.0, abs(ref.fun))
