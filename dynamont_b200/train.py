"""Pooled Baum-Welch training on top of Aligner.train_batch ("next" row N4).

The reference's ``dynamont-train`` runs a per-read M-step in C++ and averages the last <= 100 per-read estimates
in Python (train.py:19-46,185-223).  The data-parallel formulation here pools the expected counts instead:
every rank accumulates w[k] = sum gamma, x[k] = sum gamma*x, xx[k] = sum gamma*x^2 (NT_aligner_api.cpp:510-512) and
the expected transition counts over its shard, one all-reduce sums them, and the M-step (NT:519-535, 703-722) is
replicated on every rank.  Per-read parity with the reference is tested separately (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import numpy as np

from .parallel import allreduce_stats, shard_indices


def m_step(stats: dict, old_mean: np.ndarray, old_stdev: np.ndarray, min_weight: float = 0.0):
    """Emission update of NT_aligner_api.cpp:519-535 applied to pooled statistics (variance floor 1e-12, kmers
    without weight keep the old model) and the normalised transition update of NT:703-722."""
    w, x, xx = stats["w"], stats["x"], stats["xx"]
    mean, sd = old_mean.copy(), old_stdev.copy()
    ok = w > min_weight
    mu = x[ok] / w[ok]
    var = np.maximum(xx[ok] / w[ok] - mu * mu, 1e-12)
    mean[ok] = mu
    sd[ok] = np.sqrt(var)
    xm, xe = stats["xi"]
    norm = xm + xe
    trans = {"m1": xm / norm if norm > 0 else 0.0, "e1": 1.0, "e2": xe / norm if norm > 0 else 0.0}
    return mean, sd, trans


class PooledTrainer:
    def __init__(self, aligner, rank: int = 0, world: int = 1, device=None):
        self.al = aligner
        self.rank, self.world, self.device = rank, world, device

    def iteration_device(self, signals, sequences):
        """One EM iteration with the pooled statistics resident on the device: every rank adds its shard's expected counts
        to a float64 tensor of 3K + 4 entries on ``self.device`` (dyn_train_accumulate), ONE all-reduce sums it in place
        (NCCL over NVLink on GPUs; the tensor never visits the host), and the M-step runs on the device on every rank
        (dyn_train_mstep_device), which also installs the new model.  Returns (transitions, stats tensor)."""
        import torch
        import torch.distributed as dist
        K = self.al.num_kmers
        stats = torch.zeros(3 * K + 4, dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        costs = [self.al.read_cells(len(s), len(q)) for s, q in zip(signals, sequences)]
        mine = shard_indices(costs, self.rank, self.world)
        self.al.train_accumulate([signals[i] for i in mine], [sequences[i] for i in mine], stats.data_ptr())
        if stats.is_cuda:
            torch.cuda.synchronize(stats.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        if stats.is_cuda:
            torch.cuda.synchronize(stats.device)
        trans = self.al.mstep_device(stats.data_ptr())
        return trans, stats

    def iteration(self, signals, sequences, update: bool = True):
        """One EM iteration over the given reads (the full set is passed on every rank; each rank processes its
        shard).  Returns (mean, stdev, transitions, stats) — identical on all ranks."""
        costs = [self.al.read_cells(len(s), len(q)) for s, q in zip(signals, sequences)]
        mine = shard_indices(costs, self.rank, self.world)
        res, pooled = self.al.train_batch([signals[i] for i in mine], [sequences[i] for i in mine])
        ok = [r for r in res if not isinstance(r, Exception)]
        pooled["Z"] = float(sum(r["Z"] for r in ok))
        pooled["n"] = float(len(ok))
        pooled = allreduce_stats(pooled, self.device)
        mean0, sd0 = self.al.model()
        mean, sd, trans = m_step(pooled, mean0, sd0)
        if update:
            self.al.set_model(mean, sd)
        return mean, sd, trans, pooled
