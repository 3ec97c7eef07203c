"""Partitioning of a batch of independent LPs across ranks (one process per GPU).

LPs are the shardable unit of this path; there is no exchange step, so the only
communication is the final gather of per-LP results.  `shard_range` is the
contiguous split used by bench.py; `gather_results` is the one collective.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_lps: int, rank: int, world: int) -> tuple[int, int]:
    """Half-open LP-id range of `rank`: sizes differ by at most one, ranks with
    smaller index get the extra LP, the union over ranks is [0, n_lps)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_lps), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local: dict[str, np.ndarray], n_lps: int, dist=None, device=None) -> dict | None:
    """All-gather per-LP result arrays (first axis = this rank's LPs) into
    arrays over all `n_lps` LPs in LP-id order.  `dist` is torch.distributed (or
    None for a single process).  Shards may be ragged (sizes differ by one)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: np.asarray(v) for k, v in local.items()}
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_lps, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    if cap == 0:  # an empty batch: nothing to exchange
        return {k: np.asarray(v) for k, v in local.items()}
    out = {}
    for name, arr in local.items():
        arr = np.ascontiguousarray(arr)
        pad = np.zeros((cap,) + arr.shape[1:], arr.dtype)
        pad[: arr.shape[0]] = arr
        t = torch.from_numpy(pad.view(np.uint8).reshape(cap, -1))
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        full = []
        for (lo, hi), p in zip(sizes, parts):
            raw = p.cpu().numpy().reshape(-1).view(arr.dtype).reshape((cap,) + arr.shape[1:])
            full.append(raw[: hi - lo])
        out[name] = np.concatenate(full, axis=0)
    return out
