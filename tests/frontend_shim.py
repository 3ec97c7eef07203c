"""A minimal stand-in for the reference's Python modelling frontend, used only
by GPU tests on boxes where /root/reference is not mounted.  It is NOT the
reference code: it is the smallest set of classes that makes the same sequence
of calls into the `dantzig.rust` module as the reference frontend does
(python-source/dantzig/model.py:124-375, optimize.py:104-154), so that the
drop-in module sees exactly the objects, term orders and negations it would see
in production.
"""
from __future__ import annotations


def bind(rs):
    class Lin:
        def __init__(self, le):
            self.le = le

        @staticmethod
        def of(var):                                      # model.py:191-193
            return Lin(rs.PyLinExpr(coefs=[1.0], vars=[var.v]))

        def __add__(self, o):                             # model.py:202-211
            if isinstance(o, (int, float, Aff)):
                return Aff(self, 0.0) + o
            if isinstance(o, Var):
                o = Lin.of(o)
            return Lin(self.le + o.le)

        __radd__ = __add__

        def __neg__(self):
            return Lin(-self.le)

        def __sub__(self, o):                             # model.py:225-234
            if isinstance(o, (int, float, Aff)):
                return Aff(self, 0.0) - o
            if isinstance(o, Var):
                o = Lin.of(o)
            return self + (-o)

        def __mul__(self, k):                             # model.py:239-242
            return Lin(self.le * float(k))

        __rmul__ = __mul__

        def aff(self):
            return Aff(self, 0.0)

        def __le__(self, o):
            return self.aff() <= o

        def __ge__(self, o):
            return self.aff() >= o

        def __eq__(self, o):
            return self.aff() == o

    class Aff:
        def __init__(self, lin, const):
            self.lin, self.const = lin, float(const)

        def aff(self):
            return self

        def __add__(self, o):                             # model.py:286-296
            if isinstance(o, (int, float)):
                return Aff(self.lin, self.const + o)
            o = o.aff() if not isinstance(o, Var) else Lin.of(o).aff()
            return Aff(self.lin + o.lin, self.const + o.const)

        __radd__ = __add__

        def __sub__(self, o):                             # model.py:301-311
            if isinstance(o, (int, float)):
                return Aff(self.lin, self.const - o)
            o = o.aff() if not isinstance(o, Var) else Lin.of(o).aff()
            return Aff(self.lin - o.lin, self.const - o.const)

        def __neg__(self):                                # model.py:343-344
            return Aff(-self.lin, -self.const)

        def __mul__(self, k):
            return Aff(k * self.lin, k * self.const)

        __rmul__ = __mul__

        def _rows(self, o, kind):                         # model.py:323-341,351-375
            d = self - o
            le = lambda lin, b: rs.PyInequality(linexpr=lin.le, b=b)
            b = -d.const
            if kind == "le":
                return [le(d.lin, b)]
            if kind == "ge":
                return [le(-d.lin, -b)]
            return [le(d.lin, b), le(-d.lin, -b)]

        def __le__(self, o):
            return self._rows(o, "le")

        def __ge__(self, o):
            return self._rows(o, "ge")

        def __eq__(self, o):
            return self._rows(o, "eq")

        def rust(self):
            return rs.PyAffExpr(linexpr=self.lin.le, constant=self.const)

    class Var:
        def __init__(self, lb=None, ub=None):
            self.v = rs.Variable(lb=lb, ub=ub)

        @classmethod
        def nonneg(cls):
            return cls(lb=0.0, ub=None)

        @classmethod
        def free(cls):
            return cls()

        @classmethod
        def nonpos(cls):
            return cls(lb=None, ub=0.0)

        def __add__(self, o):
            return Lin.of(self) + o

        __radd__ = __add__

        def __sub__(self, o):
            return Lin.of(self) - o

        def __rsub__(self, o):
            return -Lin.of(self) + o

        def __neg__(self):
            return -Lin.of(self)

        def __mul__(self, k):
            return Lin.of(self) * k

        __rmul__ = __mul__

        def __le__(self, o):
            return Lin.of(self) <= o

        def __ge__(self, o):
            return Lin.of(self) >= o

        def __eq__(self, o):
            return Lin.of(self) == o

        __hash__ = object.__hash__

    class Solved:
        def __init__(self, sol, sign):
            self.sol, self.sign = sol, sign

        @property
        def objective_value(self):                        # optimize.py:20-27
            return self.sign * self.sol.objective_value

        def __getitem__(self, var):
            return self.sol[var.v]

    def _solve(obj, constraints, minimize):
        obj = (Lin.of(obj) if isinstance(obj, Var) else obj).aff()
        rows = [r for c in constraints for r in c]
        target = (-obj) if minimize else obj              # optimize.py:115
        return Solved(rs.solve(target.rust(), rows), -1.0 if minimize else 1.0)

    def minimize(obj, constraints=()):
        return _solve(obj, list(constraints), True)

    def maximize(obj, constraints=()):
        return _solve(obj, list(constraints), False)

    return Var, minimize, maximize
