"""Config 3 (m=2000 x n=4000 dense packing LP, lowered 6000x14000): pivots/s over a pivot
prefix on one CTA with all state in HBM, checked against the oracle's prefix."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
from dantzig_b200.model import model_from_theta
from oracle import dzo_py
m = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 60
w = generate.packing(1, m, 2 * m)
t0 = time.time(); t = Template(w.structure); t1 = time.time()
b = Batch(t, 1, max_pivots=cap)
b.upload(w.theta); b.solve(); b.sync()
r = b.download(light=True); ms = b.kernel_ms()
print("config3-like m=%d lowered %dx%d nnz %d template %.1fs launch %s" % (m, t.m, t.n_int, t.nnz, t1 - t0, b.launch_info()))
print("GPU prefix: pivots %d status %d ms %.1f pivots/s %.2f" % (r.pivots[0], r.status[0], ms, r.pivots[0] / ms * 1e3), flush=True)
b.close()
if os.environ.get('DZ_SKIP_ORACLE'):
    sys.exit(0)
t0 = time.time()
o = dzo_py.lower(model_from_theta(w.structure, w.theta[0])).solve(dzo_py.SKIP, max_pivots=cap)
dt = time.time() - t0
print("oracle(skip) prefix: pivots %d status %d %.1fs pivots/s %.2f  trace match %s objective match %s" % (
    o.pivots, o.status, dt, o.pivots / dt, o.trace_hash == int(r.trace_hash[0]), o.objective == r.objective[0]), flush=True)
