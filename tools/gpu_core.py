"""On-chip core kernel: golden parity in the automatic launch shape, then throughput sweeps
(CTAs per SM) on config 2 and on config-5-shaped LPs."""
import os, sys, json, struct, hashlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch, solve_batch
from tests import cases
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gold = os.path.join(ROOT, "tests", "golden")
bits = lambda x: struct.pack("<d", float(x)).hex()
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
if "--no-parity" not in sys.argv:
    for wl in sorted(cases.GOLDEN_WORKLOADS):
        w = cases.GOLDEN_WORKLOADS[wl]()
        g = json.load(open(os.path.join(gold, wl + ".json")))
        t0 = time.time()
        res = solve_batch(Template(w.structure), w.theta)
        bad = []
        for i, e in enumerate(g["lps"]):
            ok = (res.status[i], res.pivots[i], res.n_primal[i], int(res.trace_hash[i])) == \
                 (e["status"], e["pivots"], e["n_primal"], e["trace_hash"]) \
                 and bits(res.objective[i]) == e["objective_bits"] and sha(res.values[i]) == e["values_sha"]
            if not ok:
                bad.append((i, int(res.status[i]), e["status"], int(res.pivots[i]), e["pivots"]))
        print("PARITY", wl, "mismatches", len(bad), bad[:4], "%.1fs" % (time.time() - t0), flush=True)
def tput(name, w, **kw):
    b = Batch(Template(w.structure), w.B, **kw)
    b.upload(w.theta)
    best = 1e30
    for _ in range(2):
        b.solve(); b.sync()
        best = min(best, b.kernel_ms())
    r = b.download(light=True)
    print("TPUT", name, kw, b.launch_info(), "ms %.1f LP/s %.0f" % (best, w.B / best * 1e3),
          "status", np.bincount(r.status, minlength=5).tolist(), "Gflop exec %.2f" % (r.work[:, :4].sum() / 1e9), flush=True)
    b.close()
w2 = generate.config2(4096)
for cps in (0, 2, 4):
    tput("c2", w2, ctas_per_sm=cps, basis_home=4)
tput("c2-warp", w2, worker_warps=-1)
w5 = generate.config5(1184)
tput("c5", w5, basis_home=4)
tput("c5-old", generate.config5(592), basis_home=2)
