"""Per-phase cycle breakdown of the batched kernel (opt.profile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch
names = ["status","gather","elim","back_a","back_b","price","ratio","update","#nontriv","#pending","#solves"]
sub = ["e_search","e_b2wait","e_div+upd","e_b1wait"]
def run(w, **kw):
    t = Template(w.structure)
    b = Batch(t, w.B, profile=True, **kw)
    b.upload(w.theta); b.solve(); r = b.download(light=True)
    ms = b.kernel_ms()
    p = r.prof.astype(np.float64)
    piv = r.pivots.sum()
    print(w.name, kw, b.launch_info(), "ms %.2f LP/s %.0f" % (ms, w.B/ms*1e3), "pivots/LP %.1f" % (piv/w.B))
    tot = p[:, :8].sum()
    for i, n in enumerate(names):
        if i < 8:
            print("  %-9s %10.0f cyc/pivot  %5.1f%%" % (n, p[:, i].sum()/piv, 100*p[:, i].sum()/tot))
        else:
            print("  %-9s %10.2f per solve" % (n, p[:, i].sum()/max(p[:, 10].sum(),1)))
    for i, n in enumerate(sub):
        print("  %-9s %10.0f cyc per nontrivial step" % (n, p[:, 11+i].sum()/max(p[:,8].sum(),1)))
    print("  total cyc/pivot %.0f ; cyc per nontrivial step %.0f ; cyc per pending row %.0f" % (tot/piv, p[:,2].sum()/max(p[:,8].sum(),1), p[:,4].sum()/max(p[:,9].sum(),1)))
    b.close()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
w = generate.config2(B)
if len(sys.argv) > 2 and sys.argv[2] == "auto":
    run(w)  # the shipped launch shape only
elif len(sys.argv) > 2 and sys.argv[2] == "load":
    # same kernel, increasing load: 4 warps on 37 SMs ... the full single wave
    for b in (148, 1184, 2368, 4096):
        run(generate.config2(b), worker_warps=-1)
else:
    run(w, worker_warps=-1, ctas_per_sm=8)
    run(w, worker_warps=-1, ctas_per_sm=2)
