"""Regenerates tests/golden/*.json from the CPU oracle (LITERAL variant).

Run from the repo root:  python tests/golden/make_golden.py
The fixtures pin, per LP of each seeded workload: status, pivot count, trace
hash, objective bits, a checksum of the primal values, and a checksum of the
generated inputs (so generator drift is told apart from solver drift).
"""
import hashlib
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from dantzig_b200.model import model_from_theta  # noqa: E402
from oracle import dzo_py  # noqa: E402
from tests.cases import GOLDEN_WORKLOADS  # noqa: E402


def bits(x: float) -> str:
    return struct.pack("<d", float(x)).hex()


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main() -> None:
    only = set(sys.argv[1:])          # optional: regenerate just the named workloads
    for name, make in GOLDEN_WORKLOADS.items():
        if only and name not in only:
            continue
        w = make()
        rows = []
        for i in range(w.B):
            lo = dzo_py.lower(model_from_theta(w.structure, w.theta[i]))
            r = lo.solve(dzo_py.LITERAL)
            rows.append(dict(status=int(r.status), pivots=int(r.pivots), n_primal=int(r.n_primal),
                             trace_hash=int(r.trace_hash), objective_bits=bits(r.objective),
                             values_sha=digest(r.values)))
        out = dict(workload=name, B=w.B, m=w.m, n=w.n, theta_sha=digest(w.theta), lps=rows)
        path = os.path.join(os.path.dirname(__file__), name + ".json")
        json.dump(out, open(path, "w"), indent=0)
        hist = np.bincount([r["status"] for r in rows], minlength=5).tolist()
        print(name, "B", w.B, "status hist", hist, "pivots", sum(r["pivots"] for r in rows))


if __name__ == "__main__":
    main()
