"""Tiny runs of every launch shape (target for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, solve_batch
w = generate.mixed_batch(6, 9, 12)
t = Template(w.structure)
ref = None
for ww, home in ((0, 0), (-1, 0), (1, 2), (3, 1), (2, 3)):
    r = solve_batch(t, w.theta, worker_warps=ww, basis_home=home)
    if ref is None:
        ref = r
    assert np.array_equal(ref.trace_hash, r.trace_hash) and np.array_equal(ref.objective, r.objective)
    print("shape", ww, home, "ok", r.pivots.tolist())
w = generate.small_batch(3, 40, 48)   # M = 88: exercises multi-chunk rows in the warp fast path
t = Template(w.structure)
a = solve_batch(t, w.theta, worker_warps=-1)
b = solve_batch(t, w.theta, worker_warps=3, basis_home=1)
assert np.array_equal(a.trace_hash, b.trace_hash)
print("M=%d ok" % t.m, a.pivots.tolist())
