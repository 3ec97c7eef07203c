"""Cycle profile of the whole-GPU single-LP kernel's master CTA (opt.profile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
names = ["status", "lists", "zero+scatter", "bookkeeping+fast", "search(slow)", "compact+update", "epoch jobs", "back prep",
         "back chains", "price", "ratio", "vecupd"]
def run(name, t, theta, cap):
    b = Batch(t, 1, max_pivots=cap, profile=True)
    b.upload(np.ascontiguousarray(theta).reshape(1, -1)); b.solve(); r = b.download(light=True)
    ms = b.kernel_ms(); p = r.prof[0].astype(np.float64); piv = int(r.pivots[0])
    tot = p[:12].sum()
    print("%s cap %d: %.1f ms, %.1f pivots/s; master cycles/pivot %.0f (%.2f ms at 1.9 GHz)" % (name, cap, ms, piv / ms * 1e3, tot / piv, tot / piv / 1.9e6))
    for i, n in enumerate(names):
        print("   %-18s %10.0f cyc/pivot %5.1f%%" % (n, p[i] / piv, 100 * p[i] / tot))
    print("   per pivot: fast commits %.0f, slow steps %.0f, epochs %.1f, pending back-sub rows %.0f; grid-wide steps %.0f" % (
        p[12] / piv, p[13] / piv, p[14] / piv, p[15] / piv, r.work[0, 6] / piv))
    b.close()
which = sys.argv[1]; cap = int(sys.argv[2])
if which == "c3":
    w = generate.packing(1, 2000, 4000); run("c3", Template(w.structure), w.theta[0], cap)
else:
    model = generate.transportation_model(0, 10000, 10000, 50000, 10); t = Template(model); run("c4", t, t.pack_theta(model), cap)
