import sys, os, json, struct, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch, solve_batch
from tests import cases
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
bits = lambda x: struct.pack("<d", float(x)).hex()
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
bad = 0
for wl in sorted(cases.GOLDEN_WORKLOADS):
    w = cases.GOLDEN_WORKLOADS[wl]()
    g = json.load(open(os.path.join(GOLD, wl + ".json")))
    res = solve_batch(Template(w.structure), w.theta, worker_warps=-1)
    nb = 0
    for i, e in enumerate(g["lps"]):
        ok = (res.status[i], res.pivots[i], res.n_primal[i], int(res.trace_hash[i])) == (e["status"], e["pivots"], e["n_primal"], e["trace_hash"]) \
            and bits(res.objective[i]) == e["objective_bits"] and sha(res.values[i]) == e["values_sha"]
        nb += (not ok)
    print(wl, "warp mode mismatches", nb, "of", len(g["lps"]), flush=True)
    bad += nb
w = generate.config2(8192)
t = Template(w.structure)
for G, cps in ((-1, 6), (-1, 8), (-1, 12), (3, 12)):
    b = Batch(t, w.B, worker_warps=G, ctas_per_sm=cps)
    b.upload(w.theta)
    for rep in range(2):
        b.solve(); b.sync()
    r = b.download(light=True)
    ms = b.kernel_ms()
    print("c2 B=%d G" % w.B, G, b.launch_info(), "ms %.2f" % ms, "LP/s %.0f" % (w.B / ms * 1e3), "nonopt", int((r.status != 0).sum()), "pivots", r.pivots.sum(), "work", r.work.sum(axis=0), flush=True)
    b.close()
print("TOTAL MISMATCHES", bad)
