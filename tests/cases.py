"""Shared seeded test workloads (CPU oracle tests and GPU parity tests use the
same ones so the golden fixtures apply to both)."""
from __future__ import annotations

import numpy as np

from dantzig_b200 import generate
from dantzig_b200.model import ModelBuilder

# name -> (factory, how many LPs the oracle is asked to check)
GOLDEN_WORKLOADS = {
    "tiny_4x6": lambda: generate.small_batch(32, 4, 6),
    "small_8x16": lambda: generate.small_batch(64, 8, 16),
    "mixed_9x12": lambda: generate.mixed_batch(64, 9, 12),
    "mixed_20x40": lambda: generate.mixed_batch(48, 20, 40),
    "c2_32x64": lambda: generate.config2(48),
    "packing_24x48": lambda: generate.packing(16, 24, 48),
    # lowered 120x280: ceil(m_int/32) == 4, the widest warp-per-LP fast path
    "small_40x80": lambda: generate.small_batch(12, 40, 80),
    # the fragile family at a size where the reference's own arithmetic breaks
    # down on part of the seeds (false infeasible/unbounded, safe_divide panic)
    "mixed_60x120": lambda: generate.mixed_batch(12, 60, 120),
    # config-2 LPs on which the reference's arithmetic ends in a FALSE "unbounded"
    "c2_false_unbounded": lambda: by_ids("c2", [287, 2142, 3300, 286]),
    # 80x160 mixed: false infeasible (5, 6), safe_divide panic (23), optimal (0);
    # lowered 226x546, too large for shared memory -> exercises the HBM workspace path
    "mixed_80x160_breakdown": lambda: by_ids("mixed80", [5, 6, 23, 0]),
    # BASELINE.json configs[0]: 100x200 dense, mixed ==/<=/>= rows, 25% free variables
    # (lowered 283x683).  Seeds 0 and 1: one optimal, one where the reference's own
    # arithmetic breaks down -- parity includes reproducing that outcome.
    "c1_100x200": lambda: generate.config1(seeds=(0, 1, 2)),
    # BASELINE.json configs[4] unit: 64x128 LPs (lowered 192x448) of the config-5 batch.  Ids 184
    # (safe_divide panic), 201 and 223 (false unbounded) are the non-optimal ones among the
    # first 320 LPs of the batch; the others are optimal.
    "c5_64x128": lambda: by_ids("c5", [0, 1, 2, 3, 184, 201, 223, 5]),
}


def by_ids(kind: str, ids):
    ids = np.asarray(ids)
    if kind in ("c2", "c5"):
        m, n, fam = (32, 64, 2) if kind == "c2" else (64, 128, 5)
        senses = np.full(m, generate.LE, np.int32)
        free = np.zeros(n, bool)
    else:
        m, n, fam = 80, 160, 7
        senses = np.array([[generate.LE, generate.GE, generate.EQ][i % 3] for i in range(m)], np.int32)
        free = np.array([j % 4 == 3 for j in range(n)])
    return generate._general_workload(f"{kind}_ids", ids, m, n, senses, free, family=fam)


def ragged_models():
    """Edge cases of the lowering: variables that only appear in rows, duplicate
    terms, boxed/free/nonpositive variables, empty rows, zero coefficients."""
    out = []
    m = ModelBuilder()
    x, y, z = m.var(-1.0, 2.0), m.free(), m.var(None, 0.0)
    m.maximize([(1.0, x), (2.0, x), (-1.0, y)], 3.0)       # duplicate objective term
    m.leq([(1.0, x), (1.0, y), (0.0, z)], 4.0)              # explicit zero coefficient
    m.leq([(1.0, y), (2.0, y), (-1.0, z)], 5.0)             # duplicate row term
    m.geq([(1.0, y)], -6.0)
    out.append(("dups_zero_bounds", m.build()))
    m = ModelBuilder()
    a, b, c = m.nonneg(), m.nonneg(), m.var(0.0, 10.0)
    m.maximize([(1.0, b)])
    m.leq([(1.0, a), (1.0, b), (1.0, c)], 7.0)              # a, c only in rows
    m.leq([], 1.0)                                           # empty row
    m.eq([(1.0, a), (-1.0, c)], 0.0)
    out.append(("rows_only_vars_empty_row", m.build()))
    m = ModelBuilder()
    v = [m.nonneg() for _ in range(5)]
    m.minimize([(float(i + 1), v[i]) for i in range(5)])
    for i in range(4):
        m.geq([(1.0, v[i]), (1.0, v[i + 1])], float(i + 1))
    out.append(("chain_cover", m.build()))
    return out
