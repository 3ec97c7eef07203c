"""Large single LPs: parity of the HBM-resident mode and pivots/s on config-3-like LPs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
from dantzig_b200.model import model_from_theta
from oracle import dzo_py
def run(w, check=0, **kw):
    t0 = time.time()
    t = Template(w.structure)
    t1 = time.time()
    b = Batch(t, w.B, **kw)
    b.upload(w.theta); b.solve(); b.sync()
    r = b.download(light=True)
    ms = b.kernel_ms()
    print(w.name, "B", w.B, "lowered %dx%d nnz %d" % (t.m, t.n_int, t.nnz), kw, b.launch_info(), "template %.1fs" % (t1-t0), "ms %.1f" % ms,
          "pivots", r.pivots.tolist()[:4], "pivots/s %.1f" % (r.pivots.sum()/ms*1e3), "status", r.status.tolist()[:4], flush=True)
    b.close()
    for i in range(check):
        o = dzo_py.lower(model_from_theta(w.structure, w.theta[i])).solve(dzo_py.SKIP, max_pivots=kw.get("max_pivots", 0))
        ok = (o.status == r.status[i] and o.pivots == r.pivots[i] and o.trace_hash == int(r.trace_hash[i]) and o.objective == r.objective[i])
        print("   oracle lp", i, "status", o.status, "pivots", o.pivots, "MATCH" if ok else "MISMATCH", flush=True)
    return r
w = generate.packing(2, 200, 400)
run(w, check=2)
run(w, check=2, basis_home=3)
w = generate.mixed_batch(2, 40, 80)
run(w, check=2, basis_home=3, worker_warps=5)
if len(sys.argv) > 1:
    m = int(sys.argv[1])
    w = generate.packing(1, m, 2 * m)
    run(w, check=1 if m <= 1000 else 0, max_pivots=int(sys.argv[2]) if len(sys.argv) > 2 else 30)
