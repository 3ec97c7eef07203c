"""Host-side multi-rank logic (world_size 2, gloo, CPU): shard ranges tile the
batch and the final gather reassembles per-LP results in LP-id order."""
import os
import socket
import sys

import numpy as np
import pytest

from dantzig_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_the_batch():
    for n in (0, 1, 7, 4096, 262144, 262145):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank: int, world: int, port: int, n_lps: int, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from dantzig_b200 import generate
    from dantzig_b200.sharding import gather_results, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n_lps, rank, world)
    w = generate.small_batch(hi - lo, 4, 6, first=lo)            # this rank's LPs, by id
    local = dict(lp_id=np.arange(lo, hi, dtype=np.int64),
                 theta0=w.theta[:, :5].copy(), status=np.full(hi - lo, rank, np.int32))
    full = gather_results(local, n_lps, dist)
    dist.barrier()
    if rank == 0:
        q.put({k: v for k, v in full.items()})
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    import multiprocessing as mp

    from dantzig_b200 import generate

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_lps, world = 37, 2                                          # ragged: 19 + 18
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_lps, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(full["lp_id"], np.arange(n_lps))
    ref = generate.small_batch(n_lps, 4, 6)                       # the unsharded batch
    assert np.array_equal(full["theta0"], ref.theta[:, :5])       # same LPs, same order
    assert list(full["status"]) == [0] * 19 + [1] * 18
