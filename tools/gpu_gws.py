import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
def run(w, **kw):
    t = Template(w.structure)
    b = Batch(t, w.B, **kw)
    b.upload(w.theta); b.solve(); b.sync()
    r = b.download(light=True)
    ms = b.kernel_ms()
    print(w.name, "B", w.B, kw, b.launch_info(), "ms %.1f" % ms, "LP/s %.1f" % (w.B/ms*1e3), "pivots/s %.0f" % (r.pivots.sum()/ms*1e3), "nonopt", int((r.status != 0).sum()), flush=True)
    b.close()
    return r
w5 = generate.config5(2368)
a = run(w5)
b = run(w5, worker_warps=-1)
print("same:", np.array_equal(a.trace_hash, b.trace_hash))
run(generate.config1(range(1)))
run(generate.config1(range(8)))
run(generate.config2(4096), worker_warps=3)
