"""TEST INFRASTRUCTURE ONLY: runs in a subprocess whose DZ_LIB points at an
emulator build (tests/emu/build_emu.py) and prints one JSON line of parity
results against the golden fixtures / the oracle.

    run_child.py golden <shapes: e.g. "-1:0,0:0,1:2"> <workload:count> ...
    run_child.py kats
    run_child.py fast <workload:count> ...
"""
import hashlib
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from dantzig_b200 import Template, device_info, solve_batch, solve_model  # noqa: E402
from tests import cases, kat  # noqa: E402

bits = lambda x: struct.pack("<d", float(x)).hex()
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def golden(shapes, specs):
    out = {}
    for spec in specs:
        wl, n = spec.split(":")
        w = cases.GOLDEN_WORKLOADS[wl]()
        g = json.load(open(os.path.join(ROOT, "tests", "golden", wl + ".json")))
        n = min(int(n), w.B)
        for shape in shapes:                      # worker_warps : basis_home [: ctas_per_sm]
            G, H = shape[0], shape[1]
            cps = shape[2] if len(shape) > 2 else 0
            res = solve_batch(Template(w.structure), w.theta[:n], worker_warps=G, basis_home=H, ctas_per_sm=cps)
            bad = 0
            for i in range(n):
                e = g["lps"][i]
                ok = (res.status[i], res.pivots[i], res.n_primal[i], int(res.trace_hash[i])) == (
                    e["status"], e["pivots"], e["n_primal"], e["trace_hash"]) \
                    and bits(res.objective[i]) == e["objective_bits"] and sha(res.values[i]) == e["values_sha"]
                bad += (not ok)
            out["%s@%s" % (wl, ":".join(str(v) for v in shape))] = [bad, n]
    return out


def fast(specs):
    """Opt-in fast numerics (dz_fast.cu) against the golden fixtures: NOT bit parity -- status equal,
    objective within 1e-9 relative and (on these well-posed workloads) the same pivot count; two runs
    bit-identical; the blocked tensor-core variant equal to 1e-9 with the same pivot count.  (Primal
    values against the exact path are compared on the GPU, tests/test_fast_mode.py.)"""
    out = {}
    for spec in specs:
        wl, n = spec.split(":")
        w = cases.GOLDEN_WORKLOADS[wl]()
        g = json.load(open(os.path.join(ROOT, "tests", "golden", wl + ".json")))
        n = min(int(n), w.B)
        t = Template(w.structure)
        res = solve_batch(t, w.theta[:n], numerics="fast")
        again = solve_batch(t, w.theta[:n], numerics="fast")
        dmma = solve_batch(t, w.theta[:n], numerics="fast", worker_warps=2)   # blocked tensor-core elimination
        bad = 0
        for i in range(n):
            e = g["lps"][i]
            gold_obj = struct.unpack("<d", bytes.fromhex(e["objective_bits"]))[0]
            ok = res.status[i] == e["status"] and res.pivots[i] == e["pivots"]
            if e["status"] == 0:
                ok = ok and abs(res.objective[i] - gold_obj) <= 1e-9 * max(1.0, abs(gold_obj))
            ok = ok and bits(res.objective[i]) == bits(again.objective[i]) and res.trace_hash[i] == again.trace_hash[i]
            ok = ok and dmma.status[i] == res.status[i] and dmma.pivots[i] == res.pivots[i] and \
                abs(dmma.objective[i] - res.objective[i]) <= 1e-9 * max(1.0, abs(res.objective[i]))
            bad += (not ok)
        out["fast:" + wl] = [bad, n]
    return out


def kats():
    from oracle import dzo_py

    out = {}
    for name, model, expect in kat.rust_kats():
        s = solve_model(model)
        o = dzo_py.lower(model).solve(dzo_py.LITERAL)
        ok = (s.status, s.pivots, s.trace_hash) == (o.status, o.pivots, o.trace_hash)
        if o.status == 0:
            ok = ok and bits(s.objective) == bits(o.objective)
        out[name] = [int(not ok), 1]
    for name, model, minimize, expect in kat.python_kats():          # tests/test_optimize.py: exact ==
        s = solve_model(model)
        ok = s.status == {"optimal": 0, "unbounded": 1, "infeasible": 2}[expect[0]]
        if ok and expect[0] == "optimal":
            ok = (-s.objective if minimize else s.objective) == expect[1] and \
                all(s.values[v] == val for v, val in expect[2].items())
        out["py:" + name] = [int(not ok), 1]
    for name, model in cases.ragged_models():                        # empty / ragged / degenerate inputs
        s = solve_model(model)
        o = dzo_py.lower(model).solve(dzo_py.LITERAL)
        ok = (s.status, s.pivots, s.trace_hash, bits(s.objective)) == (o.status, o.pivots, o.trace_hash, bits(o.objective))
        out["ragged:" + name] = [int(not ok), 1]
    return out


if __name__ == "__main__":
    assert "emulator" in device_info(0)["name"].lower(), "run_child.py must run on an emulator build"
    if sys.argv[1] == "fast":
        print("EMU " + json.dumps(fast(sys.argv[2:])))
    elif sys.argv[1] == "golden":
        shapes = [tuple(int(x) for x in s.split(":")) for s in sys.argv[2].split(",")]
        print("EMU " + json.dumps(golden(shapes, sys.argv[3:])))
    else:
        print("EMU " + json.dumps(kats()))
