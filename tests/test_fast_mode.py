"""Opt-in fast numerics (dz_options.numerics = DZ_NUMERICS_FAST, dantzig_b200/csrc/dz_fast.cu).

FAST is NOT the parity path and never a default: one factorisation per pivot reused for BTRAN
(the reference factorises B and B^T separately, /root/reference/src/simplex.rs:226-236), fused
multiply-add, identity blocks of the basis eliminated symbolically.  The bar here is therefore
the one BASELINE.json's north_star states for solutions, with the tolerances written out:

    same termination status, objective within 1e-9 relative, primal values within 1e-7,

against the exact path / the golden fixtures, on LPs where the reference's own arithmetic is
well-posed; on the fixtures where the reference breaks down (false unbounded / safe_divide panic)
the fast path is checked against HiGHS instead.  The CPU tests run the kernel source on the SIMT
emulator (test infrastructure, tests/emu); the `-m gpu` tests run it on the B200 through the C ABI.
"""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")

REL_OBJ, ABS_VAL = 1e-9, 1e-7


def _emu_lib():
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(EMU, "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_fast_numerics_on_the_emulator():
    """Every size class of the tiled Gauss-Jordan inversion (k <= 32, <= 64 on 128 threads; <= 128
    on 512 threads), the generic fallback (small_40x80: k up to 80 on a 128-thread CTA), and the
    blocked tensor-core variant of each class (the emulator restates mma.m8n8k4.f64's fragment
    layout; the GPU test checks the real instruction)."""
    env = dict(os.environ, DZ_LIB=_emu_lib(), DZ_LIB_TEST_ONLY="1")
    r = subprocess.run([sys.executable, os.path.join(EMU, "run_child.py"), "fast", "tiny_4x6:8", "small_8x16:6",
                        "mixed_9x12:6", "mixed_20x40:3", "packing_24x48:2", "c2_32x64:2", "small_40x80:1",
                        "c5_64x128:1"],
                       env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("EMU ")][-1][4:])
    bad = {k: v for k, v in res.items() if v[0] != 0}
    assert not bad, bad
    assert len(res) == 8 and all(v[1] > 0 for v in res.values())


def test_exact_is_the_default():
    """dz_options_default leaves numerics at DZ_NUMERICS_EXACT (0); FAST has to be asked for."""
    import ctypes as C

    from dantzig_b200 import _capi

    o = _capi.Options()
    o.numerics = 7
    _capi.lib().dz_options_default(C.byref(o))
    assert o.numerics == _capi.NUMERICS_EXACT == 0
    assert _capi.NUMERICS_FAST == 1


# --------------------------------------------------------------------------- GPU
def _agree(ex, fa):
    """(both optimal mask, max objective rel err, max primal abs err)."""
    opt = (ex.status == 0) & (fa.status == 0)
    rel = np.abs(ex.objective[opt] - fa.objective[opt]) / np.maximum(1.0, np.abs(ex.objective[opt]))
    dv = np.abs(ex.values[opt] - fa.values[opt]).max() if opt.any() else 0.0
    return opt, (rel.max() if rel.size else 0.0), dv


def _highs_max(model):
    """HiGHS optimum of a ModelArrays LP (maximise obj_const + c.x s.t. rows <= rhs, bounds)."""
    from scipy.optimize import linprog

    n = int(model.n_vars)
    c = np.zeros(n)
    np.add.at(c, np.asarray(model.obj_var), np.asarray(model.obj_coef))
    rp = np.asarray(model.row_ptr)
    A = np.zeros((len(rp) - 1, n))
    for r in range(len(rp) - 1):
        np.add.at(A[r], np.asarray(model.row_var)[rp[r]:rp[r + 1]], np.asarray(model.row_coef)[rp[r]:rp[r + 1]])
    bounds = [(float(model.lb[j]) if model.has_lb[j] else None, float(model.ub[j]) if model.has_ub[j] else None)
              for j in range(n)]
    ref = linprog(-c, A_ub=A, b_ub=np.asarray(model.rhs), bounds=bounds, method="highs")
    assert ref.status == 0, ref.message
    return float(model.obj_const) - ref.fun


WELL_POSED = ["tiny_4x6", "small_8x16", "mixed_9x12", "mixed_20x40", "c2_32x64", "packing_24x48", "small_40x80",
              "mixed_60x120"]


@pytest.mark.gpu
@pytest.mark.parametrize("wl", WELL_POSED)
def test_fast_agrees_with_exact_on_well_posed_workloads(wl):
    from dantzig_b200 import Template, solve_batch
    from tests import cases

    w = cases.GOLDEN_WORKLOADS[wl]()
    t = Template(w.structure)
    ex = solve_batch(t, w.theta)
    fa = solve_batch(t, w.theta, numerics="fast")
    assert (ex.status == 0).all()                      # these fixtures are optimal under the reference's arithmetic
    assert (fa.status == ex.status).all()
    opt, rel, dv = _agree(ex, fa)
    assert rel <= REL_OBJ and dv <= ABS_VAL
    assert (fa.pivots == ex.pivots).all()              # same pivot count (north_star) -- here even the same trace
    assert (fa.trace_hash == ex.trace_hash).all()


@pytest.mark.gpu
@pytest.mark.parametrize("wl,count", [("c2", 1024), ("c5", 148)])
def test_fast_on_baseline_batches(wl, count):
    """Configs 2 and 5: wherever the exact path ends optimal the fast path does too, same objective
    to 1e-9 and same primal values to 1e-7; pivot-count deltas are confined to a few LPs."""
    from dantzig_b200 import Template, generate, solve_batch

    w = generate.config2(count) if wl == "c2" else generate.config5(count)
    t = Template(w.structure)
    ex = solve_batch(t, w.theta)
    fa = solve_batch(t, w.theta, numerics="fast")
    ex_opt = ex.status == 0
    assert (fa.status[ex_opt] == 0).all()
    opt, rel, dv = _agree(ex, fa)
    assert rel <= REL_OBJ and dv <= ABS_VAL
    assert (fa.pivots[opt] != ex.pivots[opt]).mean() <= 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("wl,count", [("c2", 512), ("c5", 74)])
def test_fast_blocked_tensor_core_elimination(wl, count):
    """worker_warps=2 with numerics="fast": four elimination steps per pass, the rank-4 update on the
    FP64 tensor cores (mma.m8n8k4.f64).  Same pivots as the step-by-step loop, results to rounding."""
    from dantzig_b200 import Template, generate, solve_batch

    w = generate.config2(count) if wl == "c2" else generate.config5(count)
    t = Template(w.structure)
    fa = solve_batch(t, w.theta, numerics="fast")
    bl = solve_batch(t, w.theta, numerics="fast", worker_warps=2)
    assert (bl.status == fa.status).all() and (fa.status == 0).all()
    assert (bl.pivots == fa.pivots).mean() >= 0.98
    rel = np.abs(bl.objective - fa.objective) / np.maximum(1.0, np.abs(fa.objective))
    assert rel.max() <= REL_OBJ and np.abs(bl.values - fa.values).max() <= ABS_VAL


@pytest.mark.gpu
def test_fast_is_deterministic_and_launch_shape_invariant():
    from dantzig_b200 import Template, generate, solve_batch

    w = generate.config2(256)
    t = Template(w.structure)
    a = solve_batch(t, w.theta, numerics="fast")
    b = solve_batch(t, w.theta, numerics="fast")
    c = solve_batch(t, w.theta[::-1].copy(), numerics="fast", ctas_per_sm=2)
    for x, y in ((a, b), (a, None)):
        if y is None:
            assert (a.objective.view(np.uint64) == c.objective[::-1].view(np.uint64)).all()
            assert (a.trace_hash == c.trace_hash[::-1]).all()
        else:
            assert (x.objective.view(np.uint64) == y.objective.view(np.uint64)).all()
            assert (x.values.view(np.uint64) == y.values.view(np.uint64)).all()
            assert (x.trace_hash == y.trace_hash).all()


@pytest.mark.gpu
@pytest.mark.parametrize("wl", ["c2_false_unbounded", "c5_64x128", "c1_100x200"])
def test_fast_against_highs_where_the_reference_breaks_down(wl):
    """On these fixtures the reference's tolerance-free arithmetic ends part of the LPs in a false
    'unbounded' or in the safe_divide panic (tests/cases.py).  The generator makes every LP feasible
    and bounded, so HiGHS is the yardstick: whatever the fast path calls optimal must carry the
    HiGHS objective; it may also break down (same algorithm), it must not invent another optimum."""
    from dantzig_b200 import Template, solve_batch
    from dantzig_b200.model import model_from_theta
    from tests import cases

    w = cases.GOLDEN_WORKLOADS[wl]()
    t = Template(w.structure)
    fa = solve_batch(t, w.theta, numerics="fast")
    ex = solve_batch(t, w.theta)
    n_opt = 0
    for i in range(w.B):
        if fa.status[i] != 0:
            continue
        best = _highs_max(model_from_theta(w.structure, w.theta[i]))
        assert abs(fa.objective[i] - best) <= 1e-7 * max(1.0, abs(best)), (i, fa.objective[i], best)
        n_opt += 1
    assert n_opt >= (ex.status == 0).sum()            # at least as many solved as the exact path


@pytest.mark.gpu
def test_fast_single_lp_entry_on_config1():
    """dz_solve_model with numerics = FAST: BASELINE configs[0] (100 x 200, mixed rows, free variables;
    lowered 283 x 683).  The exact path reproduces the reference's outcome on these seeds (optimal,
    safe_divide panic, false 'unbounded'); the fast path ends at the HiGHS optimum on all three."""
    from dantzig_b200 import solve_model
    from dantzig_b200.model import model_from_theta
    from tests import cases

    w = cases.GOLDEN_WORKLOADS["c1_100x200"]()
    for i in range(w.B):
        m = model_from_theta(w.structure, w.theta[i])
        s = solve_model(m, numerics="fast")
        best = _highs_max(m)
        assert s.status == 0 and abs(s.objective - best) <= 1e-7 * max(1.0, abs(best)), (i, s.status, s.objective, best)


@pytest.mark.gpu
def test_fast_through_the_rust_module():
    """rust.solve_batch(..., numerics="fast") -- the batched front door of the drop-in module."""
    import dantzig_b200.rust as rs
    from tests.test_gpu_parity import _frontend

    _frontend(rs)                                       # dantzig.exceptions for the module's error mapping

    def lp(k):
        x, y = rs.Variable(lb=0.0, ub=None), rs.Variable(lb=0.0, ub=None)
        obj = rs.PyAffExpr(linexpr=rs.PyLinExpr([3.0 + k, 2.0], [x, y]), constant=0.0)
        cons = [rs.PyInequality(linexpr=rs.PyLinExpr([1.0, 1.0], [x, y]), b=4.0),
                rs.PyInequality(linexpr=rs.PyLinExpr([1.0, 3.0], [x, y]), b=6.0)]
        return obj, cons

    lps = [lp(k) for k in range(5)]
    ex = rs.solve_batch([o for o, _ in lps], [c for _, c in lps])
    fa = rs.solve_batch([o for o, _ in lps], [c for _, c in lps], numerics="fast")
    for a, b in zip(ex, fa):
        assert abs(a.objective_value - b.objective_value) <= 1e-9 * max(1.0, abs(a.objective_value))
    with pytest.raises(ValueError):
        rs.solve_batch([lps[0][0]], [lps[0][1]], numerics="sloppy")
