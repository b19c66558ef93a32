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


class MultiDeviceAligner:
    """Product-level multi-GPU front: the reference fans reads out to an ``mp.Pool`` of single-read workers
    (segment.py:296-324); here the reads of a batch are dealt over the GPUs of one box (``shard_indices``: sorted by
    in-band cells, serpentine), every GPU runs its shard through its own ``Aligner`` handle from its own host thread (the C
    ABI call releases the GIL), and the results are merged back in input order.  No data-path collective: reads are
    independent."""

    def __init__(self, model_file: str, pore, devices, mode: str = "basic", band: int = 400, **kw):
        from .aligner import Aligner
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("MultiDeviceAligner: empty device list")
        self.aligners = [Aligner(model_file, pore, mode=mode, band=band, device=d, **kw) for d in self.devices]

    def set_option(self, key: str, value: float) -> None:
        for al in self.aligners:
            al.set_option(key, value)

    def align_batch(self, signals, sequences, calc_probabilities: bool = False, raise_errors: bool = False):
        import threading
        n = len(signals)
        assert len(sequences) == n
        world = len(self.aligners)
        costs = [self.aligners[0].read_cells(len(s), len(q)) for s, q in zip(signals, sequences)]
        shards = [shard_indices(costs, r, world) for r in range(world)]
        out = [None] * n
        errors = [None] * world

        def work(r):
            try:
                idx = shards[r]
                if idx.size == 0:
                    return
                res = self.aligners[r].align_batch([signals[i] for i in idx], [sequences[i] for i in idx], calc_probabilities)
                for i, v in zip(idx, res):
                    out[int(i)] = v
            except Exception as e:  # a CUDA / runtime error of one device fails the call
                errors[r] = e
        threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        if raise_errors:
            for v in out:
                if isinstance(v, Exception):
                    raise v
        return out
