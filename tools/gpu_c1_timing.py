"""Config 1 (one 100 x 200 LP, lowered 283 x 683): kernel time per solve, exact path against the opt-in fast numerics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import Template, Batch
from tests import cases
w = cases.GOLDEN_WORKLOADS["c1_100x200"]()
t = Template(w.structure)
for numerics in ("exact", "fast"):
    for i in range(w.B):
        b = Batch(t, 1, numerics=numerics)
        b.upload(w.theta[i:i + 1]); b.solve(); b.sync(); b.solve(); b.sync()
        r = b.download(light=True)
        print("C1", numerics, "seed", i, "status", int(r.status[0]), "pivots", int(r.pivots[0]), "kernel ms %.1f" % b.kernel_ms(), b.launch_info(), flush=True)
        b.close()
