"""config-5-shaped LPs at small per-GPU batch sizes: which launch shape has the lowest batch time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
def tput(name, w, **kw):
    b = Batch(Template(w.structure), w.B, **kw)
    b.upload(w.theta); b.solve(); b.sync()
    ms = b.kernel_ms(); r = b.download(light=True)
    print("TPUT B=%d %-10s" % (w.B, name), b.launch_info(), "ms %.1f LP/s %.0f" % (ms, w.B / ms * 1e3), flush=True)
    b.close()
for B in (296, 592, 1184, 4736):
    w5 = generate.config5(B)
    tput("warp", w5, worker_warps=-1)
    tput("cta-3w", w5, worker_warps=3, basis_home=2)
    if B <= 1184:
        tput("core-2", w5, basis_home=4, ctas_per_sm=2)
