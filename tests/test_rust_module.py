"""Host logic of the drop-in `dantzig.rust` module (no solve: CPU only), and --
where the reference checkout is present (this container, not the GPU box) --
the reference's own frontend and expression tests running on top of it."""
import os
import sys

import pytest

from tests.conftest import reference_frontend_dir

REF = reference_frontend_dir()


@pytest.fixture()
def rs():
    import dantzig_b200.rust as rs

    return rs


def test_surface(rs):
    assert rs.__name__ == "dantzig.rust"
    for name in ("Variable", "PyLinExpr", "PyAffExpr", "PyInequality", "PySolution", "solve", "solve_batch"):
        assert hasattr(rs, name)
    assert rs.Variable.__module__ == "dantzig.rust"
    with pytest.raises(TypeError):
        rs.Variable(0.0, None)                       # keyword-only, pyobjs.rs:25
    with pytest.raises(TypeError):
        rs.Variable(lb=0.0)                          # both required
    a, b = rs.Variable(lb=0.0, ub=None), rs.Variable(lb=None, ub=2.5)
    assert b.id == a.id + 1 and a.lb == 0.0 and a.ub is None and b.ub == 2.5
    with pytest.raises(AttributeError):
        a.id = 3


def test_linexpr_algebra(rs):
    x, y = rs.Variable(lb=0.0, ub=None), rs.Variable(lb=0.0, ub=None)
    ex, ey = rs.PyLinExpr([1.0], [x]), rs.PyLinExpr([1.0], [y])
    assert (ex + ey + ex).map_ids_to_coefs() == {x.id: 2.0, y.id: 1.0}      # pyobjs.rs:78-104
    assert (-(ex + ey)).map_ids_to_coefs() == {x.id: -1.0, y.id: -1.0}
    assert ((ex + ey) * 2.0).map_ids_to_coefs() == {x.id: 2.0, y.id: 2.0}
    aff = rs.PyAffExpr(linexpr=ex + ey, constant=5.0)
    assert aff.constant == 5.0 and aff.pylinexpr.map_ids_to_coefs() == {x.id: 1.0, y.id: 1.0}
    rs.PyInequality(linexpr=ex, b=1.0)


def test_solve_batch_host_side(rs):
    """Argument checking, and no CPU fallback: without a GPU the call fails loudly."""
    from dantzig_b200 import device_count

    x = rs.Variable(lb=0.0, ub=None)
    obj = rs.PyAffExpr(linexpr=rs.PyLinExpr([1.0], [x]), constant=0.0)
    row = rs.PyInequality(linexpr=rs.PyLinExpr([1.0], [x]), b=1.0)
    with pytest.raises(ValueError):
        rs.solve_batch([obj, obj], [[row]])
    if device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            rs.solve_batch([obj], [[row]])


@pytest.mark.skipif(REF is None, reason="reference frontend not present (build() installs it under baseline/_ref)")
def test_reference_frontend_expression_tests(rs):
    """tests/test_model.py of the reference, unmodified frontend on our module."""
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "dantzig" or k.startswith("dantzig.")}
    sys.modules["dantzig.rust"] = rs
    sys.path.insert(0, REF)
    try:
        import dantzig as dz

        x, y = dz.Variable.nonneg(), dz.Variable.nonneg()
        eq = lambda a, b: a.map_ids_to_coefs() == b.map_ids_to_coefs()
        assert eq(-x, -1.0 * x) and eq(x + x, 2 * x) and eq(x - y, -y + x)
        assert eq(x + y + x, y + 2 * x) and eq(2 * x + 2 * y, (x + y) * 2)
        assert eq(-(x + y), -y - x) and eq(2 * x - x, x.to_linexpr())
        f, g = dz.Variable.free(), dz.Variable.free()
        a1, a2 = (f + g + 5) + (f + g + 5), 10.0 + 2 * f + 2 * g
        assert eq(a1.linexpr, a2.linexpr) and a1.constant == a2.constant
        # the solve path must refuse to run without a GPU instead of falling back
        from dantzig_b200 import device_count

        if device_count() == 0:
            with pytest.raises(RuntimeError, match="no CUDA device"):
                dz.Minimize(2 * x - 2 * y).subject_to(y == 3).solve()
        else:
            s = dz.Minimize(2 * x - 2 * y).subject_to(y == 3).solve()
            assert s.objective_value == -6.0 and s[x] == 0.0 and s[y] == 3.0
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "dantzig" or k.startswith("dantzig.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
