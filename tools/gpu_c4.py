"""Config-4 family at 1/10 scale (1000 supply + 1000 demand rows, 5000 arcs with 10+10
nonzeros: lowered 7000x17000): pivots/s over a pivot prefix, checked against the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
from oracle import dzo_py
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 40
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
arcs = int(sys.argv[4]) if len(sys.argv) > 4 else int(2.5 * S)
model = generate.transportation_model(0, S, S, arcs, k)
t0 = time.time(); t = Template(model); theta = t.pack_theta(model); t1 = time.time()
b = Batch(t, 1, max_pivots=cap)
b.upload(theta[None, :]); b.solve(); b.sync()
r = b.download(light=True); ms = b.kernel_ms()
print("config4-like S=D=%d arcs=%d k=%d lowered %dx%d nnz %d template %.1fs launch %s" % (S, arcs, k, t.m, t.n_int, t.nnz, t1 - t0, b.launch_info()))
print("GPU prefix: pivots %d status %d ms %.1f pivots/s %.2f" % (r.pivots[0], r.status[0], ms, r.pivots[0] / ms * 1e3), flush=True)
b.close()
if t.m > 12000:
    print('oracle check skipped at this size (dense %d x %d basis per pivot on one CPU core)' % (t.m, t.m)); sys.exit(0)
if os.environ.get('DZ_SKIP_ORACLE'):
    sys.exit(0)
t0 = time.time()
o = dzo_py.lower(model).solve(dzo_py.SKIP, max_pivots=cap)
dt = time.time() - t0
print("oracle(skip) prefix: pivots %d status %d %.1fs pivots/s %.2f  trace match %s objective match %s" % (
    o.pivots, o.status, dt, o.pivots / dt, o.trace_hash == int(r.trace_hash[0]), o.objective == r.objective[0]), flush=True)
