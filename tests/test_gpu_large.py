"""Parity of the single-large-LP path (BASELINE.json configs[2] and configs[3]) with the
oracle: pivot prefixes at the full sizes, a full solve at a reduced size.  Bit-exact:
status, pivot count, pivot trace, objective bits, basic values, basis.

The oracle follows these sizes with its sparse-row variant (DZO_SPARSE: the SKIP
variant's operations in the same order on rows stored as ordered maps; checked against
SKIP in tests/test_oracle_variants.py).  Run on the B200 box:  pytest -m gpu"""
import numpy as np
import pytest

from dantzig_b200 import Batch, Template, generate
from dantzig_b200.model import model_from_theta

pytestmark = pytest.mark.gpu


def _gpu_prefix(template, theta, cap, **kw):
    b = Batch(template, 1, max_pivots=cap, trace_cap=max(cap, 1), **kw)
    b.upload(np.ascontiguousarray(theta).reshape(1, -1))
    b.solve()
    r = b.download()
    b.close()
    return r


def _same(r, o, cap):
    assert (int(r.status[0]), int(r.pivots[0]), int(r.n_primal[0]), int(r.trace_hash[0])) == (
        o.status, o.pivots, o.n_primal, o.trace_hash)
    assert np.array_equal(r.trace[0, : min(o.pivots, cap)], o.trace)
    assert np.array_equal(np.array([r.objective[0]]).view(np.uint64), np.array([o.objective]).view(np.uint64))
    assert np.array_equal(r.x_basic[0].view(np.uint64), o.x_basic.view(np.uint64))
    assert np.array_equal(r.basis[0], o.basis)


def test_config3_family_full_solve(oracle):
    """Dense packing LP (the config-3 family) at m=150 x n=300, lowered 450x1050, solved
    to optimality: 606 pivots, every one of them compared."""
    w = generate.packing(1, 150, 300)
    r = _gpu_prefix(Template(w.structure), w.theta[0], 0)
    lo = oracle.lower(model_from_theta(w.structure, w.theta[0]))
    o = lo.solve(oracle.SKIP, trace_cap=r.trace.shape[1])
    assert o.status == 0 and o.pivots > 400
    _same(r, o, r.trace.shape[1])
    # independent optimum (HiGHS): maximise c.x, A x <= b, x >= 0
    from scipy.optimize import linprog

    th, m, n = w.theta[0], w.m, w.n
    c = th[2:2 + n]
    A = th[2 + n:2 + n + m * n].reshape(m, n)
    b = th[2 + n + m * n:2 + n + m * n + m]
    ref = linprog(-c, A_ub=A, b_ub=b, bounds=[(0, None)] * n, method="highs")
    assert ref.status == 0 and abs(r.objective[0] + ref.fun) <= 1e-9 * max(1.0, abs(ref.fun))


def test_config3_full_size_prefix(oracle):
    """BASELINE configs[2] at full size (m=2000 x n=4000 dense, lowered 6000x14000): the
    first 40 pivots."""
    cap = 40
    w = generate.packing(1, 2000, 4000)
    r = _gpu_prefix(Template(w.structure), w.theta[0], cap)
    o = oracle.lower(model_from_theta(w.structure, w.theta[0])).solve(oracle.SPARSE, max_pivots=cap, trace_cap=cap)
    assert o.status == 4 and o.pivots == cap
    _same(r, o, cap)


@pytest.mark.parametrize("scale,arcs,cap", [(1000, 2500, 40), (10000, 50000, 60)], ids=["eighth-scale", "full-size"])
def test_config4_prefix(oracle, scale, arcs, cap):
    """BASELINE configs[3]: transportation-style sparse LP, `scale` supply + `scale` demand
    rows, `arcs` columns with 10+10 nonzeros each.  Full size (m=20k x n=50k) is lowered
    70000x170000 with 2.17 M nonzeros; the oracle follows it with sparse rows."""
    model = generate.transportation_model(0, scale, scale, arcs, 10)
    t = Template(model)
    r = _gpu_prefix(t, t.pack_theta(model), cap)
    o = oracle.lower(model).solve(oracle.SPARSE, max_pivots=cap, trace_cap=cap)
    assert o.status == 4 and o.pivots == cap
    _same(r, o, cap)
