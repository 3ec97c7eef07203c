import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
w = generate.config2(8192)
t = Template(w.structure)
for G, cps in ((-1, 2), (-1, 3), (-1, 5)):
    b = Batch(t, w.B, worker_warps=G, ctas_per_sm=cps)
    b.upload(w.theta)
    for rep in range(2):
        b.solve(); b.sync()
    r = b.download(light=True)
    ms = b.kernel_ms()
    print(os.environ.get("DZ_LIB", "default")[-12:], "c2 B=%d G" % w.B, G, "cps", cps, "ms %.2f" % ms, "LP/s %.0f" % (w.B / ms * 1e3), "nonopt", int((r.status != 0).sum()), flush=True)
    b.close()
