"""Multi-GPU plumbing: reads are independent units, so they shard across ranks with no data-path collective
(SURVEY.md §8e).  Training adds exactly one collective per iteration: an all-reduce (sum) of the pooled
sufficient statistics.  torch.distributed is used for rendezvous/collectives only (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_indices(costs, rank: int, world: int) -> np.ndarray:
    """Indices of the reads rank `rank` of `world` processes: reads sorted by cost (in-band cells) descending
    and dealt round-robin in serpentine order, so every rank gets the same number of reads (+-1) and nearly the
    same total cost.  Deterministic; the union over ranks is a partition of range(len(costs))."""
    costs = np.asarray(costs)
    order = np.argsort(-costs, kind="stable")
    n = order.size
    pos = np.arange(n)
    rnd, col = np.divmod(pos, world)
    owner = np.where(rnd % 2 == 0, col, world - 1 - col)
    return np.sort(order[owner == rank])


def allreduce_stats(stats: dict, device=None) -> dict:
    """Sum the pooled training statistics {w, x, xx, xi, Z, n} over all ranks (one all-reduce of 3K+4 doubles,
    6.3 MB for 9-mers).  A no-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    K = stats["w"].size
    flat = np.concatenate([stats["w"], stats["x"], stats["xx"], stats["xi"], [stats.get("Z", 0.0), stats.get("n", 0.0)]])
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    flat = t.cpu().numpy()
    return {"w": flat[:K], "x": flat[K:2 * K], "xx": flat[2 * K:3 * K], "xi": flat[3 * K:3 * K + 2],
            "Z": float(flat[3 * K + 2]), "n": float(flat[3 * K + 3])}
