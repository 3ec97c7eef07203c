"""First GPU contact: parity of the batched kernel against the oracle + rough timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dantzig_b200 import generate, Template, Batch, device_info, measure_fp64_peak
from dantzig_b200.model import model_from_theta
from oracle import dzo_py

print(device_info(0))
print("fp64 peak (mul+sub, fma) GFLOP/s:", measure_fp64_peak(0))

def parity(w, nchk, **kw):
    t = Template(w.structure)
    b = Batch(t, w.B, trace_cap=0, **kw)
    b.upload(w.theta); b.solve(); r = b.download()
    print(w.name, "launch", b.launch_info(), "kernel_ms", b.kernel_ms())
    bad = 0
    for i in range(min(nchk, w.B)):
        lo = dzo_py.lower(model_from_theta(w.structure, w.theta[i]))
        o = lo.solve(dzo_py.SKIP)
        same = (o.status == r.status[i] and o.pivots == r.pivots[i] and o.trace_hash == int(r.trace_hash[i])
                and o.objective == r.objective[i] and np.array_equal(o.values, r.values[i]) and np.array_equal(o.x_basic, r.x_basic[i]))
        if not same:
            bad += 1
            if bad < 5:
                print("MISMATCH lp", i, "oracle", o.status, o.pivots, hex(o.trace_hash), o.objective, "gpu", r.status[i], r.pivots[i], hex(int(r.trace_hash[i])), r.objective[i])
    print(w.name, "checked", min(nchk, w.B), "mismatches", bad, "status hist", np.bincount(r.status, minlength=5), "pivots mean", r.pivots.mean(), "work/LP", r.work.mean(axis=0))
    b.close()
    return bad

bad = 0
bad += parity(generate.small_batch(16, 4, 6), 16)
bad += parity(generate.small_batch(64, 8, 16), 64)
bad += parity(generate.mixed_batch(64, 9, 12), 64)
bad += parity(generate.config2(64), 64)
for tpr in (1, 2, 4):
    bad += parity(generate.config2(64), 16, worker_warps=tpr)
w = generate.config2(4096)
t = Template(w.structure)
for tpr in (1, 2, 4):
    for cps in (1, 2):
        b = Batch(t, w.B, worker_warps=tpr, ctas_per_sm=cps)
        b.upload(w.theta)
        for rep in range(2):
            b.solve(); b.sync()
        r = b.download(light=True)
        ms = b.kernel_ms()
        print("c2 B=4096 tpr", tpr, "cps", cps, b.launch_info(), "ms %.2f" % ms, "LP/s %.0f" % (w.B / ms * 1e3), "optimal", (r.status == 0).sum(), "pivots", r.pivots.sum())
        b.close()
print("TOTAL MISMATCHES", bad)
sys.exit(1 if bad else 0)
