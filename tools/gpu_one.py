"""One launch of a chosen kernel on a chosen workload (ncu target).

    gpu_one.py c2|c5 <B> [worker_warps] [basis_home] [ctas_per_sm] [fast]   batched kernels
    gpu_one.py c3|c4 <prefix>                                          whole-GPU single-LP kernel
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 296
if which in ("c3", "c4"):
    if which == "c3":
        w = generate.packing(1, 2000, 4000); t = Template(w.structure); theta = w.theta[:1]
    else:
        model = generate.transportation_model(0, 10000, 10000, 50000, 10); t = Template(model); theta = t.pack_theta(model)[None, :]
    b = Batch(t, 1, max_pivots=n)
    b.upload(theta); b.solve(); r = b.download(light=True)
    print(which, "prefix", n, b.launch_info(), "ms %.2f" % b.kernel_ms(), "pivots", int(r.pivots[0]), "status", int(r.status[0]))
else:
    ww = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    home = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    cps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    numerics = "fast" if len(sys.argv) > 6 and sys.argv[6] == "fast" else "exact"
    w = generate.config2(n) if which == "c2" else generate.config5(n)
    b = Batch(Template(w.structure), w.B, worker_warps=ww, basis_home=home, ctas_per_sm=cps, numerics=numerics)
    b.upload(w.theta); b.solve(); r = b.download(light=True)
    print(which, "B", n, b.launch_info(), "ms %.2f" % b.kernel_ms(), "optimal", int((r.status == 0).sum()), "pivots", int(r.pivots.sum()))
