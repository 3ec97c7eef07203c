import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
w = generate.config2(4096)
t = Template(w.structure)
for G, cps in ((1, 8), (1, 12), (1, 16), (1, 24), (1, 32), (2, 12), (2, 20)):
    b = Batch(t, w.B, worker_warps=G, ctas_per_sm=cps)
    b.upload(w.theta)
    for rep in range(2):
        b.solve(); b.sync()
    r = b.download(light=True)
    ms = b.kernel_ms()
    print("c2 B=4096 G", G, b.launch_info(), "ms %.2f" % ms, "LP/s %.0f" % (w.B / ms * 1e3), "nonopt", int((r.status != 0).sum()), "pivots", r.pivots.sum())
    b.close()
