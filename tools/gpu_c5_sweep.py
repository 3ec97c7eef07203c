"""config-5-shaped LPs: launch-shape sweep of the general kernel (worker warps x CTAs per SM) and
the core kernel with one / two CTAs per SM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
def tput(name, w, **kw):
    b = Batch(Template(w.structure), w.B, **kw)
    b.upload(w.theta); b.solve(); b.sync()
    ms = b.kernel_ms(); r = b.download(light=True)
    print("TPUT", name, kw, b.launch_info(), "ms %.1f LP/s %.0f" % (ms, w.B / ms * 1e3), "status", np.bincount(r.status, minlength=5).tolist(), flush=True)
    b.close()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
w5 = generate.config5(B)
for ww in (2, 3, 4, 5):
    for cps in (8, 12):
        tput("c5-general", w5, worker_warps=ww, ctas_per_sm=cps, basis_home=2)
tput("c5-warp", w5, worker_warps=-1)
tput("c5-core-1", generate.config5(592), basis_home=4, ctas_per_sm=1)
tput("c5-core-2", generate.config5(592), basis_home=4, ctas_per_sm=2)
