"""Throughput of the batched kernel on the other BASELINE shapes (C5 unit, C1, packing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
def run(w, **kw):
    t = Template(w.structure)
    b = Batch(t, w.B, **kw)
    b.upload(w.theta); b.solve(); b.sync()
    r = b.download(light=True)
    ms = b.kernel_ms()
    print(w.name, "B", w.B, "lowered %dx%d" % (t.m, t.n_int), kw, b.launch_info(), "ms %.1f" % ms, "LP/s %.1f" % (w.B/ms*1e3),
          "pivots/s %.0f" % (r.pivots.sum()/ms*1e3), "status", np.bincount(r.status, minlength=5).tolist(), "pivots/LP %.0f" % r.pivots.mean(), flush=True)
    b.close()
    return r
run(generate.config2(4096))
run(generate.config2(4096), basis_home=1)
run(generate.config5(1184))
run(generate.config5(1184), worker_warps=4, ctas_per_sm=4)
run(generate.config1(range(8)))
run(generate.config1(range(1)))
run(generate.packing(8, 100, 200))
