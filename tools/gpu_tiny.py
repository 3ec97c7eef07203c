"""Tiny-LP latency: 32 LPs of 8x16 (the smoke() workload) and single small models, per launch shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch, solve_model
from tests import kat
w = generate.small_batch(32, 8, 16)
t = Template(w.structure)
for kw in (dict(), dict(basis_home=4), dict(worker_warps=-1), dict(worker_warps=1, basis_home=1)):
    b = Batch(t, w.B, **kw)
    b.upload(w.theta)
    best = 1e9
    for _ in range(5):
        b.solve(); b.sync(); best = min(best, b.kernel_ms())
    r = b.download(light=True)
    print("tiny 32x(8x16)", kw, b.launch_info(), "kernel %.3f ms" % best, "pivots", int(r.pivots.sum()), flush=True)
    b.close()
name, model, expect = kat.rust_kats()[3]
for _ in range(2):
    t0 = time.perf_counter(); s = solve_model(model); dt = time.perf_counter() - t0
    print("solve_model", name, "status", s.status, "wall %.3f ms" % (dt * 1e3))
