"""Per-phase cycle breakdown of the on-chip core kernel (opt.profile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
names = ["status", "lists", "gather", "elim", "back", "price", "ratio", "update"]
def run(w, **kw):
    b = Batch(Template(w.structure), w.B, profile=True, basis_home=4, **kw)
    b.upload(w.theta); b.solve(); r = b.download(light=True)
    ms = b.kernel_ms()
    p = r.prof.astype(np.float64)
    piv = r.pivots.sum(); solves = max(p[:, 9].sum(), 1); steps = max(p[:, 8].sum(), 1)
    print(w.name, kw, b.launch_info(), "ms %.2f LP/s %.0f" % (ms, w.B / ms * 1e3), "pivots/LP %.1f" % (piv / w.B),
          "handed over", int(p[:, 13].sum()))
    tot = p[:, :8].sum()
    for i, n in enumerate(names):
        print("  %-8s %9.0f cyc/pivot %5.1f%%" % (n, p[:, i].sum() / piv, 100 * p[:, i].sum() / tot))
    print("  total cyc/pivot %.0f | real steps/solve %.1f | nr avg %.1f | overflow rows/solve %.2f" % (
        tot / piv, steps / solves, p[:, 14].sum() / solves, p[:, 15].sum() / solves))
    print("  per real step: bookkeeping %.0f  search+barrierA %.0f  update+barrierB %.0f cycles" % (
        p[:, 10].sum() / steps, p[:, 11].sum() / steps, p[:, 12].sum() / steps))
    b.close()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
which = sys.argv[2] if len(sys.argv) > 2 else "c2"
w = generate.config2(B) if which == "c2" else generate.config5(B)
for cps in [int(a) for a in sys.argv[3:]] or [0]:
    run(w, ctas_per_sm=cps)
