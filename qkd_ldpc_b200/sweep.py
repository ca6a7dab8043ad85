"""Frame-batch scheduler for QBER sweeps: the data-parallel replacement of the reference's trial loop
(QKD_LDPC_batch_simulation, src/simulation.cpp:192-316).

Trials are independent, so a sweep shards trial indices over ranks (one process per GPU) with no exchange while
decoding; the only collective is ONE all-reduce (sum) of integer statistics at the end of the sweep. Statistics are
kept as integer histograms of iterations_num over successful frames, so the reduced result -- and everything derived
from it the way the reference derives it (src/simulation.cpp:252-312) -- is independent of the number of ranks and
of the reduction order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of trial indices [lo, hi) for `rank`; blocks differ by at most one trial."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PointStats:
    """Integer statistics of one (matrix, QBER) point: [hist[0..max_it], n_sp, n_ldpc, n_trials, n_iter_total]."""

    def __init__(self, max_it: int):
        self.max_it = int(max_it)
        self.vec = np.zeros(self.max_it + 1 + 4, np.int64)

    @property
    def width(self) -> int:
        return self.vec.size

    def add(self, iterations, result) -> None:
        it = np.asarray(iterations, np.int64)
        res = np.asarray(result, np.uint8)
        ok = (res & 1) != 0
        self.vec[: self.max_it + 1] += np.bincount(it[ok], minlength=self.max_it + 1)[: self.max_it + 1]
        self.vec[self.max_it + 1] += int(ok.sum())
        self.vec[self.max_it + 2] += int((ok & ((res & 2) != 0)).sum())
        self.vec[self.max_it + 3] += it.size
        self.vec[self.max_it + 4] += int(it.sum())


@dataclass
class PointResult:
    """What the reference puts in one CSV row (struct sim_result, src/simulation.hpp:29-43)."""
    mean: float
    std_dev: float
    it_min: int
    it_max: int
    ratio_sp: float
    ratio_ldpc: float
    fer: float
    n_trials: int
    frame_iterations: int


def derive(vec: np.ndarray, max_it: int) -> PointResult:
    """src/simulation.cpp:252-312 from the integer histogram (population std-dev over successful frames)."""
    hist = np.asarray(vec[: max_it + 1], np.int64)
    n_sp, n_ldpc, n_trials, n_iter = (int(v) for v in vec[max_it + 1: max_it + 5])
    its = np.arange(max_it + 1, dtype=np.float64)
    mean = std = 0.0
    it_min, it_max = max_it, 0
    if n_sp > 0:
        mean = float((hist * its).sum() / n_sp)
        std = math.sqrt(float((hist * (its - mean) ** 2).sum() / n_sp))
        nz = np.flatnonzero(hist)
        it_min, it_max = int(nz[0]), int(nz[-1])
    if it_min == max_it:
        it_min = 0  # src/simulation.cpp:306
    ratio_sp = n_sp / n_trials if n_trials else 0.0
    ratio_ldpc = n_ldpc / n_trials if n_trials else 0.0
    return PointResult(mean, std, it_min, it_max, ratio_sp, ratio_ldpc, 1.0 - ratio_ldpc, n_trials, n_iter)


def allreduce_stats(stats: np.ndarray, device=None) -> np.ndarray:
    """The sweep's single collective: sum of the [points x width] int64 statistics over all ranks.
    NCCL when the process group is NCCL (tensor on this rank's GPU), gloo on CPU."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    t = torch.from_numpy(np.ascontiguousarray(stats, np.int64))
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def binomial_ci95(k: int, n: int) -> tuple[float, float]:
    """Clopper-Pearson 95 % interval for a proportion (used for the fp32 FER bar)."""
    from scipy.stats import beta
    lo = 0.0 if k == 0 else float(beta.ppf(0.025, k, n - k + 1))
    hi = 1.0 if k == n else float(beta.ppf(0.975, k + 1, n - k))
    return lo, hi
