import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
w = generate.config2(8192)
t = Template(w.structure)
for kw in ({"worker_warps": -1, "ctas_per_sm": 4}, {"worker_warps": -1, "ctas_per_sm": 6}, {"worker_warps": -1}):
    b = Batch(t, w.B, **kw)
    b.upload(w.theta)
    for rep in range(2):
        b.solve(); b.sync()
    r = b.download(light=True)
    ms = b.kernel_ms()
    print(os.environ.get("DZ_LIB", "default")[-12:], "c2 B=%d" % w.B, kw, b.launch_info()["ctas_per_sm"], "ms %.2f" % ms, "LP/s %.0f" % (w.B / ms * 1e3), "nonopt", int((r.status != 0).sum()), flush=True)
    b.close()
