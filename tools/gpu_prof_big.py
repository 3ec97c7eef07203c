"""Phase breakdown (opt.profile) of a large single LP prefix: config 3 or config 4."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
names = ["status","gather","elim","back_a","back_b","price","ratio","update","#nontriv","#pending","#solves"]
sub = ["e_search","e_b2wait","e_div+upd","e_b1wait"]
which, cap = sys.argv[1], int(sys.argv[2])
if which == "c3":
    w = generate.packing(1, 2000, 4000); t = Template(w.structure); theta = w.theta
else:
    model = generate.transportation_model(0, 10000, 10000, 50000, 10); t = Template(model); theta = t.pack_theta(model)[None, :]
b = Batch(t, 1, profile=True, max_pivots=cap)
b.upload(theta); b.solve(); r = b.download(light=True)
ms = b.kernel_ms(); p = r.prof.astype(np.float64); piv = r.pivots.sum()
print(which, "lowered %dx%d" % (t.m, t.n_int), "pivots", piv, "ms %.1f" % ms, "pivots/s %.2f" % (piv/ms*1e3))
tot = p[:, :8].sum()
for i, n in enumerate(names):
    if i < 8: print("  %-9s %12.0f cyc/pivot  %5.1f%%" % (n, p[:, i].sum()/piv, 100*p[:, i].sum()/tot))
    else: print("  %-9s %10.2f per solve" % (n, p[:, i].sum()/max(p[:, 10].sum(),1)))
for i, n in enumerate(sub):
    print("  %-9s %10.0f cyc per nontrivial step" % (n, p[:, 11+i].sum()/max(p[:,8].sum(),1)))
