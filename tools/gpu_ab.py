"""A/B of library builds: python tools/gpu_ab.py build/ab/a.so build/ab/b.so ...

Each build runs in its own process (DZ_LIB selects the library): golden parity in
the warp-per-LP and auto launch shapes, then config-2 and config-5-unit throughput.
"""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import json, struct, hashlib
    import numpy as np
    from dantzig_b200 import generate, Template, Batch, solve_batch
    from tests import cases
    gold = os.path.join(ROOT, "tests", "golden")
    bits = lambda x: struct.pack("<d", float(x)).hex()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    bad = 0
    for wl in sorted(cases.GOLDEN_WORKLOADS):
        w = cases.GOLDEN_WORKLOADS[wl]()
        if os.environ.get("AB_FAST") and w.m >= 100:
            continue  # config-1 size on one warp per LP takes most of a minute
        g = json.load(open(os.path.join(gold, wl + ".json")))
        for G in ((-1,) if os.environ.get("AB_FAST") else (0, -1, 3)):
            res = solve_batch(Template(w.structure), w.theta, worker_warps=G)
            for i, e in enumerate(g["lps"]):
                ok = (res.status[i], res.pivots[i], res.n_primal[i], int(res.trace_hash[i])) == \
                     (e["status"], e["pivots"], e["n_primal"], e["trace_hash"]) \
                     and bits(res.objective[i]) == e["objective_bits"] and sha(res.values[i]) == e["values_sha"]
                bad += (not ok)
    # ceil(m_int/32) == 4 has no golden fixture: check it against the oracle directly
    from oracle import dzo_py
    from dantzig_b200.model import model_from_theta
    w = generate.small_batch(24, 40, 80)
    res = solve_batch(Template(w.structure), w.theta, worker_warps=-1)
    for i in range(w.B):
        o = dzo_py.lower(model_from_theta(w.structure, w.theta[i])).solve(dzo_py.SKIP, 0, 0)
        ok = (res.status[i], res.pivots[i], int(res.trace_hash[i])) == (o.status, o.pivots, int(o.trace_hash)) \
            and bits(res.objective[i]) == bits(o.objective)
        bad += (not ok)
    out = {"lib": os.environ.get("DZ_LIB"), "mismatches": int(bad)}
    for name, w, G, reps in (("c2", generate.config2(4096), 0, 3),
                             ("c5w", generate.config5(1024), -1, 1))[:1 if os.environ.get("AB_FAST") else 2]:
        b = Batch(Template(w.structure), w.B, worker_warps=G)
        b.upload(w.theta)
        best = 1e30
        for _ in range(reps):
            b.solve(); b.sync()
            best = min(best, b.kernel_ms())
        out[name] = round(w.B / best * 1e3, 1)
        b.close()
    print("AB", json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--child":
        child()
    else:
        for lib in sys.argv[1:]:
            env = dict(os.environ, DZ_LIB=os.path.join(ROOT, lib), DZ_LIB_TEST_ONLY="1")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False)
