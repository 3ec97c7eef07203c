"""Opt-in fast numerics (dz_fast.cu) on the GPU: agreement with the exact path on the golden
workloads, then throughput and the per-phase cycle profile on configs 2 and 5."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch, solve_batch
from tests import cases

PH = ["status", "lists", "build", "gj", "ftran", "btran", "price", "ratio", "update"]

def agree(name, w, n):
    t = Template(w.structure)
    ex = solve_batch(t, w.theta[:n])
    fa = solve_batch(t, w.theta[:n], numerics="fast")
    opt = (ex.status == 0) & (fa.status == 0)
    rel = np.abs(ex.objective[opt] - fa.objective[opt]) / np.maximum(1.0, np.abs(ex.objective[opt]))
    dv = np.abs(ex.values[opt] - fa.values[opt]).max() if opt.any() else 0.0
    print("AGREE", name, "n", n, "status equal", int((ex.status == fa.status).sum()), "both optimal", int(opt.sum()),
          "obj rel max %.2e" % (rel.max() if rel.size else 0), "values abs max %.2e" % dv,
          "same pivot count", int((ex.pivots == fa.pivots).sum()), "same trace", int((ex.trace_hash == fa.trace_hash).sum()),
          "exact status", np.bincount(ex.status, minlength=5).tolist(), "fast status", np.bincount(fa.status, minlength=5).tolist(), flush=True)

def tput(name, w, reps=2, **kw):
    b = Batch(Template(w.structure), w.B, numerics="fast", **kw)
    b.upload(w.theta)
    best = 1e30
    for _ in range(reps):
        b.solve(); b.sync()
        best = min(best, b.kernel_ms())
    r = b.download(light=True)
    print("TPUT", name, kw, b.launch_info(), "ms %.2f LP/s %.0f" % (best, w.B / best * 1e3),
          "status", np.bincount(r.status, minlength=5).tolist(), "Gflop exec %.2f" % (r.work[:, :4].sum() / 1e9),
          "pivots/LP %.1f" % r.pivots.mean(), flush=True)
    if kw.get("profile"):
        p = r.prof.astype(float).sum(0); piv = r.pivots.sum()
        tot = p[:9].sum()
        print("   cycles/pivot %.0f:" % (tot / piv), " ".join("%s %.0f" % (PH[i], p[i] / piv) for i in range(9)),
              "| k avg %.1f max %d" % (p[9] / p[10], r.prof[:, 13].max()),
              "| gj per step: search %.0f div+publish+bar %.0f update(warp0) %.0f bar %.0f" % tuple(p[i] / p[9] for i in (11, 12, 14, 15)), flush=True)
    b.close()

if "--no-agree" not in sys.argv:
    for wl in ("small_8x16", "mixed_20x40", "c2_32x64", "packing_24x48", "small_40x80", "mixed_60x120", "c2_false_unbounded", "c5_64x128"):
        w = cases.GOLDEN_WORKLOADS[wl]()
        agree(wl, w, w.B)
    agree("c2 first 1024", generate.config2(1024), 1024)
    agree("c5 first 296", generate.config5(296), 296)
w2 = generate.config2(4096)
tput("c2-prof", generate.config2(592), profile=True)
if "--prof-only" in sys.argv:
    tput("c5-prof", generate.config5(148), profile=True)
    sys.exit(0)
tput("c2", w2)
tput("c2-blocked-dmma", w2, worker_warps=2)
w5 = generate.config5(2368)
tput("c5-prof", generate.config5(296), profile=True)
tput("c5", w5)
tput("c5-blocked-dmma", w5, worker_warps=2)
tput("c5-1184", generate.config5(1184))
