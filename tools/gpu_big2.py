"""Whole-GPU single-LP kernel on configs 3 and 4: pivot-prefix timings, checked against the
sparse oracle where the cap is small enough."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
from dantzig_b200.model import model_from_theta
from oracle import dzo_py
def run(name, t, theta, cap, lower_fn, check, **kw):
    b = Batch(t, 1, max_pivots=cap, **kw)
    b.upload(np.ascontiguousarray(theta).reshape(1, -1)); b.solve(); b.sync()
    r = b.download(light=True); ms = b.kernel_ms()
    msg = "%s lowered %dx%d cap %d %s launch %s: pivots %d status %d ms %.1f pivots/s %.1f Gflop %.2f" % (
        name, t.m, t.n_int, cap, kw, b.launch_info(), r.pivots[0], r.status[0], ms, r.pivots[0] / ms * 1e3, r.work[:, :4].sum() / 1e9)
    b.close()
    if check:
        t0 = time.time()
        o = lower_fn().solve(dzo_py.SPARSE, max_pivots=cap)
        msg += " | oracle %.1fs match %s" % (time.time() - t0, (o.status, o.pivots, o.trace_hash, o.objective) ==
                                             (int(r.status[0]), int(r.pivots[0]), int(r.trace_hash[0]), float(r.objective[0])))
    print(msg, flush=True)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
caps = [int(a) for a in sys.argv[2:]] or [60, 400]
if which in ("c3", "both"):
    w = generate.packing(1, 2000, 4000)
    t = Template(w.structure)
    low = lambda: dzo_py.lower(model_from_theta(w.structure, w.theta[0]))
    for cap in caps:
        run("c3", t, w.theta[0], cap, low, cap <= 100)
if which in ("c4", "both"):
    model = generate.transportation_model(0, 10000, 10000, 50000, 10)
    t = Template(model)
    theta = t.pack_theta(model)
    low = lambda: dzo_py.lower(model)
    for cap in caps:
        run("c4", t, theta, cap, low, cap <= 500)
