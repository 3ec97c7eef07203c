"""The reference's known-answer LPs, restated as ModelBuilder programs.

Each case cites the reference test it restates.  `expect` is what the
reference asserts: ("optimal", objective, {var: value}) in MAXIMISE form as
seen by Simplex (Rust tests) or in user form (Python tests, `sense`), or
("unbounded",) / ("infeasible",).
"""
from __future__ import annotations

from dantzig_b200.model import ModelBuilder


def rust_kats():
    out = []

    def case(name, build):
        mb = ModelBuilder()
        expect = build(mb)
        out.append((name, mb.build(), expect))

    # src/simplex.rs:485-501
    def nonneg_1(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(4.0, x), (3.0, y)])
        m.leq([(1.0, x), (-1.0, y)], 1.0).leq([(2.0, x), (-1.0, y)], 3.0).leq([(1.0, y)], 5.0)
        return ("optimal", 31.0, {x: 4.0, y: 5.0})
    case("nonneg_1", nonneg_1)

    # src/simplex.rs:504-522
    def nonneg_2(m):
        a, b, c = m.nonneg(), m.nonneg(), m.nonneg()
        m.maximize([(5.0, a), (4.0, b), (3.0, c)])
        m.leq([(2.0, a), (3.0, b), (1.0, c)], 5.0)
        m.leq([(4.0, a), (1.0, b), (2.0, c)], 11.0)
        m.leq([(3.0, a), (4.0, b), (2.0, c)], 8.0)
        return ("optimal", 13.0, {a: 2.0, b: 0.0, c: 1.0})
    case("nonneg_2", nonneg_2)

    # src/simplex.rs:525-562
    def nonneg_3(m):
        x1, x2, x3, x4 = (m.nonneg() for _ in range(4))
        m.maximize([(300.0, x1), (90.0, x2), (400.0, x3), (150.0, x4)])
        m.leq([(35000.0, x1), (10000.0, x2), (25000.0, x3), (90000.0, x4)], 120000.0)
        m.leq([(4.0, x1), (2.0, x2), (7.0, x3), (3.0, x4)], 12.0)
        m.leq([(1.0, x1), (1.0, x2)], 1.0)
        for v in (x1, x2, x3, x4):
            m.leq([(1.0, v)], 1.0)
        return ("optimal", 750.0, {x1: 1.0, x2: 0.0, x3: 1.0, x4: 1.0 / 3.0})
    case("nonneg_3", nonneg_3)

    # src/simplex.rs:565-583
    def nonneg_4(m):
        a, b, c = m.nonneg(), m.nonneg(), m.nonneg()
        m.maximize([(10.0, a), (12.0, b), (12.0, c)])
        m.leq([(1.0, a), (2.0, b), (2.0, c)], 20.0)
        m.leq([(2.0, a), (1.0, b), (2.0, c)], 20.0)
        m.leq([(2.0, a), (2.0, b), (1.0, c)], 20.0)
        return ("optimal", 136.0, {a: 4.0, b: 4.0, c: 4.0})
    case("nonneg_4", nonneg_4)

    # src/simplex.rs:586-602
    def nonneg_5(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(-1.0, x), (-1.0, y)])
        m.leq([(-2.0, x), (-1.0, y)], 4.0).leq([(-2.0, x), (4.0, y)], -8.0)
        m.leq([(-1.0, x), (3.0, y)], -7.0)
        return ("optimal", -7.0, {x: 7.0, y: 0.0})
    case("nonneg_5", nonneg_5)

    # src/simplex.rs:605-623
    def nonneg_6(m):
        a, b, c = m.nonneg(), m.nonneg(), m.nonneg()
        m.maximize([(-10.0, a), (-12.0, b), (-12.0, c)])
        m.leq([(-1.0, a), (-2.0, b), (-2.0, c)], -20.0)
        m.leq([(-2.0, a), (-1.0, b), (-2.0, c)], -20.0)
        m.leq([(-2.0, a), (-2.0, b), (-1.0, c)], -20.0)
        return ("optimal", -136.0, {a: 4.0, b: 4.0, c: 4.0})
    case("nonneg_6", nonneg_6)

    # src/simplex.rs:626-642
    def nonneg_8(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(-2.0, x), (3.0, y)])
        m.leq([(-1.0, x), (1.0, y)], -1.0).leq([(-1.0, x), (-2.0, y)], -2.0).leq([(1.0, y)], 1.0)
        return ("optimal", -1.0, {x: 2.0, y: 1.0})
    case("nonneg_8", nonneg_8)

    # src/simplex.rs:645-674
    def nonneg_9(m):
        x1, x2, x3, x4, x5, x6 = (m.nonneg() for _ in range(6))
        m.maximize([(2.0, x2), (3.0, x5)], 10.0)
        m.leq([(1.0, x1), (-1.0, x2), (1.0, x4)], 4.0)
        m.leq([(-1.0, x1), (1.0, x2), (-1.0, x4)], -4.0)
        m.leq([(3.0, x2), (1.0, x3), (-1.0, x5)], 12.0)
        m.leq([(-3.0, x2), (-1.0, x3), (1.0, x5)], -12.0)
        m.leq([(1.0, x2), (1.0, x4), (2.0, x5)], 14.0)
        m.leq([(-1.0, x2), (-1.0, x4), (-2.0, x5)], -14.0)
        m.leq([(2.0, x2), (1.0, x5), (1.0, x6)], 13.0)
        m.leq([(-2.0, x2), (-1.0, x5), (-1.0, x6)], -13.0)
        return ("optimal", 33.0, {x1: 8.0, x2: 4.0, x3: 5.0, x4: 0.0, x5: 5.0, x6: 0.0})
    case("nonneg_9", nonneg_9)

    # src/simplex.rs:677-687
    def no_constraints(m):
        x = m.nonneg()
        m.maximize([(-3.0, x)], 2.0)
        return ("optimal", 2.0, {x: 0.0})
    case("nonneg_no_constraints", no_constraints)

    # src/simplex.rs:690-703
    def variable_constraints(m):
        x, y = m.var(1.0, 1.0), m.var(-3.0, -1.0)
        m.maximize([(1.0, x), (-1.0, y)], 5.0)
        return ("optimal", 9.0, {x: 1.0, y: -3.0})
    case("variable_constraints", variable_constraints)

    # src/simplex.rs:706-720
    def unbounded_1(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(-1.0, x), (4.0, y)])
        m.leq([(-2.0, x), (-1.0, y)], 4.0).leq([(-2.0, x), (4.0, y)], -8.0)
        m.leq([(-1.0, x), (3.0, y)], -7.0)
        return ("unbounded",)
    case("unbounded_1", unbounded_1)

    # src/simplex.rs:723-734
    def unbounded_2(m):
        x = m.nonneg()
        m.maximize([(1.0, x)])
        m.leq([(-2.0, x)], -4.0)
        return ("unbounded",)
    case("unbounded_2", unbounded_2)

    # src/simplex.rs:737-747
    def unbounded_nc(m):
        x = m.nonneg()
        m.maximize([(1.0, x)], 10.0)
        return ("unbounded",)
    case("unbounded_no_constraints", unbounded_nc)

    # src/simplex.rs:750-763
    def infeasible_1(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(1.0, x), (1.0, y)])
        m.leq([(1.0, x)], -1.0).leq([(5.0, y)], 0.5)
        return ("infeasible",)
    case("infeasible_1", infeasible_1)

    # src/simplex.rs:766-778
    def infeasible_2(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(1.0, x), (-1.0, y)])
        m.leq([(1.0, x), (1.0, y)], -1.0)
        return ("infeasible",)
    case("infeasible_2", infeasible_2)

    # src/simplex.rs:781-796
    def infeasible_3(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(1.0, x), (1.0, y)])
        m.leq([(1.0, x), (1.0, y)], 1.0).leq([(-1.0, x), (-1.0, y)], -1.0)
        m.leq([(1.0, x), (1.0, y)], 2.0).leq([(-1.0, x), (-1.0, y)], -2.0)
        return ("infeasible",)
    case("infeasible_3", infeasible_3)
    return out


def python_kats():
    """tests/test_optimize.py, tests/test_exceptions.py, README.md:60-73 of the
    reference, lowered term-by-term exactly as the reference frontend lowers
    them (merge-on-add keeps first-appearance order, pyobjs.rs:78-104).
    expect: (status, user-sense objective, {var: value}); `minimize` tells how
    to map the solver's objective back (optimize.py:23-24)."""
    out = []

    def case(name, build):
        mb = ModelBuilder()
        minimize, expect = build(mb)
        out.append((name, mb.build(), minimize, expect))

    # tests/test_optimize.py:4-11   min 2x-2y st y == 3
    def p1(m):
        x, y = m.nonneg(), m.nonneg()
        m.minimize([(2.0, x), (-2.0, y)])
        m.eq([(1.0, y)], 3.0)
        return True, ("optimal", -6.0, {x: 0.0, y: 3.0})
    case("problem_1", p1)

    # tests/test_optimize.py:14-23  min 2x-2y st y<=5, x>=y+1, y==5
    def p2(m):
        x, y = m.nonneg(), m.nonneg()
        m.minimize([(2.0, x), (-2.0, y)])
        m.leq([(1.0, y)], 5.0)
        # x >= y + 1: affexpr = x - (y+1) -> linexpr [x, -y], constant -1 -> b = 1
        m.geq([(1.0, x), (-1.0, y)], 1.0)
        m.eq([(1.0, y)], 5.0)
        return True, ("optimal", 2.0, {x: 6.0, y: 5.0})
    case("problem_2", p2)

    # tests/test_optimize.py:26-35
    def p3(m):
        x, y, z = m.nonneg(), m.nonneg(), m.nonneg()
        m.minimize([(1.0, x), (1.0, y), (-1.0, z)])
        m.leq([(1.0, x), (1.0, y), (1.0, z)], 1.0)
        return True, ("optimal", -1.0, {x: 0.0, y: 0.0, z: 1.0})
    case("problem_3", p3)

    # tests/test_optimize.py:38-47
    def p4(m):
        x, y, z = m.nonneg(), m.nonneg(), m.nonneg()
        m.minimize([(1.0, x), (1.0, y), (1.0, z)])
        m.eq([(1.0, x), (-1.0, y)], -2.0)
        return True, ("optimal", 2.0, {x: 0.0, y: 2.0, z: 0.0})
    case("problem_4", p4)

    # tests/test_optimize.py:50-59
    def minmax_min(m):
        x, y = m.nonneg(), m.nonneg()
        m.minimize([(-1.0, x)])
        m.leq([(1.0, x), (1.0, y)], 1.0)
        return True, ("optimal", -1.0, {x: 1.0, y: 0.0})
    case("minmax_min", minmax_min)

    def minmax_max(m):
        x, y = m.nonneg(), m.nonneg()
        m.maximize([(1.0, x)])
        m.leq([(1.0, x), (1.0, y)], 1.0)
        return False, ("optimal", 1.0, {x: 1.0, y: 0.0})
    case("minmax_max", minmax_max)

    # tests/test_optimize.py:62-71  (-3.0 <= x <= 3.0 collapses to x <= 3.0)
    def nonstandard(m):
        x, y, z = m.var(-2.0, 2.0), m.free(), m.var(None, 0.0)
        m.minimize([(1.0, x), (1.0, y), (1.0, z)])
        m.eq([(1.0, y)], 4.0)
        m.leq([(1.0, x)], 3.0)
        m.geq([(1.0, z)], -1.0)
        return True, ("optimal", 1.0, {x: -2.0, y: 4.0, z: -1.0})
    case("non_standard_variables", nonstandard)

    # tests/test_optimize.py:74-114
    def inventory(m):
        x1, x2, x3 = m.nonneg(), m.nonneg(), m.nonneg()
        z1, z2, z3 = m.nonneg(), m.nonneg(), m.nonneg()
        m.minimize([(0.5, x1), (3.5, x2), (5.0, x3), (1.0, z1), (5.5, z2), (1.5, z3)])
        m.geq([(1.0, x1)], 50.0)
        m.geq([(1.0, x2), (1.0, z1)], 75.0)
        m.geq([(1.0, x3), (1.0, z2)], 100.0)
        # z_1 == x_1 - d_1  ->  z1 - x1 == -50
        m.eq([(1.0, z1), (-1.0, x1)], -50.0)
        # z_2 == x_2 + z_1 - d_2
        m.eq([(1.0, z2), (-1.0, x2), (-1.0, z1)], -75.0)
        m.eq([(1.0, z3), (-1.0, x3), (-1.0, z2)], -100.0)
        return True, ("optimal", 637.5, {x1: 125.0, x2: 0.0, x3: 100.0})
    case("inventory", inventory)

    # tests/test_exceptions.py:6-9
    def unbounded(m):
        x = m.nonneg()
        m.minimize([(-1.0, x)])
        return True, ("unbounded",)
    case("unbounded_error", unbounded)

    # tests/test_exceptions.py:12-16
    def infeasible(m):
        x, y = m.nonneg(), m.nonneg()
        m.minimize([(1.0, x), (1.0, y)])
        m.eq([(1.0, x), (1.0, y)], 1.0)
        m.eq([(1.0, x), (1.0, y)], 2.0)
        return True, ("infeasible",)
    case("infeasible_error", infeasible)

    # optimize.py:94-101 docstring example
    def doc_min(m):
        x, y = m.var(1.0, None), m.var(None, 2.0)
        m.minimize([(1.0, x), (-5.0, y)])
        return True, ("optimal", -9.0, {x: 1.0, y: 2.0})
    case("doc_minimize", doc_min)
    return out
