"""Pins the oracle against every known-answer test the reference holds for the
path.  CPU only."""
import numpy as np
import pytest

from tests import kat

STATUS = {"optimal": 0, "unbounded": 1, "infeasible": 2}
EPS = 1e-12  # src/simplex.rs:9,477-482


# ---- src/linalg.rs known answers -------------------------------------------
def test_lu_factorization_golden(oracle):
    # src/linalg.rs:323-345 (exact ==)
    lu, p = oracle.lu_factorize(np.array([[3.0, 17.0, 10.0], [2.0, 4.0, -2.0], [6.0, 18.0, -12.0]]))
    assert list(p) == [2, 2]
    assert list(lu.ravel()) == [6.0, 18.0, -12.0, 1.0 / 3.0, 8.0, 16.0, 1.0 / 2.0, -1.0 / 4.0, 6.0]


@pytest.mark.parametrize("variant", [0, 1])
def test_lu_solve_golden(oracle, variant):
    # src/linalg.rs:361-380 (exact ==)
    a = np.array([[6.0, 18.0, 3.0], [2.0, 12.0, 1.0], [4.0, 15.0, 3.0]])
    assert list(oracle.lu_solve(a, np.array([3.0, 19.0, 0.0]), variant)) == [-3.0, 3.0, -11.0]
    a = np.array([[2.0, 0.0, 0.0], [4.0, 1.0, 0.0], [3.0, 0.0, 1.0]])
    assert list(oracle.lu_solve(a, np.array([1.0, 2.0, 2.0]), variant)) == [0.5, 0.0, 0.5]


def test_csc_from_dense_golden(oracle):
    # src/linalg.rs:383-393
    cp, ri, val = oracle.dense_to_csc(np.array([[1.0, 0.0, 2.0], [0.0, 0.0, 3.0], [4.0, 5.0, 6.0]]))
    assert list(ri) == [0, 2, 2, 0, 1, 2]
    assert list(cp) == [0, 2, 3, 6]
    assert list(val) == [1.0, 4.0, 5.0, 2.0, 3.0, 6.0]


def test_matrix_roundtrip_drops_zeros(oracle):
    # src/linalg.rs:348-358
    cp, ri, val = oracle.dense_to_csc(np.array([[0.0, 1.0], [0.0, 2.0]]))
    assert list(cp) == [0, 0, 2] and list(ri) == [0, 1] and list(val) == [1.0, 2.0]


def test_neg_t_dot_golden(oracle):
    # src/linalg.rs:436-446
    dense = np.arange(12, dtype=np.float64).reshape(3, 4)
    cp, ri, val = oracle.dense_to_csc(dense)
    out = oracle.neg_t_dot(3, 4, cp, ri, val, np.array([1.0, 2.0, 3.0]))
    assert list(out) == [-32.0, -38.0, -44.0, -50.0]


# ---- src/simplex.rs known answers --------------------------------------------
@pytest.mark.parametrize("name,model,expect", kat.rust_kats(), ids=[k[0] for k in kat.rust_kats()])
@pytest.mark.parametrize("variant", [0, 1])
def test_rust_simplex_kats(oracle, name, model, expect, variant):
    r = oracle.lower(model).solve(variant, trace_cap=64)
    assert r.status == STATUS[expect[0]]
    if expect[0] == "optimal":
        assert abs(r.objective - expect[1]) <= EPS
        for v, val in expect[2].items():
            assert abs(r.values[list(oracle.lower(model).orig_var).index(v)] - val) <= EPS


# survey-derived internal dimensions and pivot counts
DIMS = {
    "nonneg_1": (5, 9, 2), "nonneg_2": (6, 12, 6), "nonneg_3": (11, 19, 11),
    "nonneg_4": (6, 12, 3), "nonneg_5": (5, 9, 3), "nonneg_6": (6, 12, 3),
    "nonneg_8": (5, 9, 3), "nonneg_9": (14, 26, 9), "nonneg_no_constraints": (1, 3, 1),
    "variable_constraints": (4, 8, 4),
}


def test_rust_kat_dims_and_traces(oracle):
    cases = {k[0]: k[1] for k in kat.rust_kats()}
    for name, (m, n, piv) in DIMS.items():
        lo = oracle.lower(cases[name])
        assert (lo.m, lo.n_int) == (m, n), name
        assert lo.solve(0).pivots == piv, name
    t = oracle.lower(cases["nonneg_2"]).solve(0, trace_cap=16).trace
    assert [tuple(x) for x in t] == [(0, 7, 0), (0, 8, 2), (1, 6, 8), (1, 2, 3), (0, 8, 4), (1, 10, 7)]
    t = oracle.lower(cases["nonneg_1"]).solve(0, trace_cap=16).trace
    assert [tuple(x) for x in t] == [(0, 5, 0), (0, 6, 2)]
    # not exactly 750: the sum order of objective_value is unspecified in the reference
    assert abs(oracle.lower(cases["nonneg_3"]).solve(0).objective - 750.0) <= EPS


# ---- tests/test_optimize.py, tests/test_exceptions.py (exact float ==) -------
@pytest.mark.parametrize("name,model,minimize,expect", kat.python_kats(),
                         ids=[k[0] for k in kat.python_kats()])
@pytest.mark.parametrize("variant", [0, 1])
def test_python_kats(oracle, name, model, minimize, expect, variant):
    lo = oracle.lower(model)
    r = lo.solve(variant)
    assert r.status == STATUS[expect[0]]
    if expect[0] == "optimal":
        obj = -r.objective if minimize else r.objective
        assert obj == expect[1]
        order = list(lo.orig_var)
        for v, val in expect[2].items():
            assert r.values[order.index(v)] == val


def test_empty_model_panics(oracle):
    # 0x0 basis: the reference panics
    from dantzig_b200.model import ModelBuilder

    mb = ModelBuilder()
    x = mb.free()
    mb.maximize([(1.0, x)])
    assert oracle.lower(mb.build()).solve(0).status == oracle.PANIC
