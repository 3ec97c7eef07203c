"""Host lowering (dz_template_create, product code) against the oracle's
restatement of Simplex::new -- the boundary's structural half.  CPU only."""
import numpy as np
import pytest

from dantzig_b200 import Template
from dantzig_b200.model import model_from_theta
from tests import cases, kat


def _check(oracle, model):
    lo = oracle.lower(model)
    t = Template(model)
    assert (t.m, t.n_int, t.n_orig) == (lo.m, lo.n_int, lo.n_orig)
    a = t.arrays()
    lv = t.lowered_values(t.pack_theta(model))
    # the template keeps structural entries whose value is an exact zero; the
    # reference drops them (linalg.rs:254-270).  Compare after dropping.
    keep = lv["val"] != 0.0
    col_of = np.repeat(np.arange(t.n_int), np.diff(a["col_ptr"]))
    cp = np.concatenate([[0], np.cumsum(np.bincount(col_of[keep], minlength=t.n_int))])
    assert np.array_equal(cp, lo.col_ptr)
    assert np.array_equal(a["row_idx"][keep], lo.row_idx)
    assert np.array_equal(lv["val"][keep], lo.val)
    assert np.array_equal(lv["c"], lo.c) and lv["c0"] == lo.c0
    assert np.array_equal(lv["b"].view(np.uint64), lo.b.view(np.uint64))  # keeps -0.0 of -lb
    for k in ("basis0", "nonbasis0", "orig_var", "pos_index", "neg_index"):
        assert np.array_equal(a[k], getattr(lo, k)), k


@pytest.mark.parametrize("name,model,expect", kat.rust_kats(), ids=[k[0] for k in kat.rust_kats()])
def test_lowering_rust_kats(oracle, name, model, expect):
    _check(oracle, model)


@pytest.mark.parametrize("name,model,minimize,expect", kat.python_kats(),
                         ids=[k[0] for k in kat.python_kats()])
def test_lowering_python_kats(oracle, name, model, minimize, expect):
    _check(oracle, model)


@pytest.mark.parametrize("name,model", cases.ragged_models(), ids=[c[0] for c in cases.ragged_models()])
def test_lowering_ragged(oracle, name, model):
    _check(oracle, model)


@pytest.mark.parametrize("wl", ["mixed_9x12", "c2_32x64", "packing_24x48"])
def test_lowering_dense_templates(oracle, wl):
    w = cases.GOLDEN_WORKLOADS[wl]()
    for i in (0, w.B - 1):
        _check(oracle, model_from_theta(w.structure, w.theta[i]))


def test_lowering_random_sparse_structures(oracle):
    rng = np.random.default_rng(7)
    from dantzig_b200.model import ModelBuilder

    for _ in range(40):
        mb = ModelBuilder()
        nv = int(rng.integers(1, 7))
        vs = [mb.var(lb=[None, 0.0, -1.5][rng.integers(3)], ub=[None, 2.5][rng.integers(2)])
              for _ in range(nv)]
        pick = lambda k: [(float(rng.integers(-3, 4)), vs[rng.integers(nv)]) for _ in range(k)]
        mb.maximize(pick(int(rng.integers(0, 4))), float(rng.integers(-2, 3)))
        for _ in range(int(rng.integers(0, 5))):
            [mb.leq, mb.geq, mb.eq][rng.integers(3)](pick(int(rng.integers(0, 4))), float(rng.integers(-4, 5)))
        model = mb.build()
        if len(model.obj_var) + len(model.row_var) == 0:
            continue
        _check(oracle, model)


def test_template_rejects_bad_indices():
    from dantzig_b200._capi import DzError
    from dantzig_b200.model import ModelBuilder

    mb = ModelBuilder()
    x = mb.nonneg()
    mb.maximize([(1.0, x)])
    model = mb.build()
    model.obj_var[0] = 5
    with pytest.raises(DzError):
        Template(model)


@pytest.mark.parametrize("shape", [(10, 10, 40, 1), (20, 20, 120, 3)])
def test_lowering_transportation(oracle, shape):
    """Sparse structure of BASELINE configs[3]: ragged rows, >= rows negated."""
    from dantzig_b200 import generate

    _check(oracle, generate.transportation_model(0, *shape))
