"""The N > 1 path of the sweep on CPU: world_size-2 gloo processes shard the trials, all-reduce the integer statistics
once, and every rank derives the same numbers as a single process over all trials -- and as the oracle's restatement of
the reference's statistics loop (src/simulation.cpp:252-312)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from qkd_ldpc_b200 import sweep

MAX_IT = 100
POINTS = 3
TRIALS = 1001


def synthetic_results(point):
    rng = np.random.default_rng(100 + point)
    it = rng.integers(3, 60, TRIALS).astype(np.int64)
    ok = rng.random(TRIALS) < (0.9, 0.5, 0.0)[point]
    keys = ok & (rng.random(TRIALS) < 0.98)
    it = np.where(ok, it, MAX_IT)
    res = ok.astype(np.uint8) | (keys.astype(np.uint8) << 1)
    return it, res


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = np.zeros((POINTS, MAX_IT + 5), np.int64)
    for pt in range(POINTS):
        it, res = synthetic_results(pt)
        lo, hi = sweep.shard_range(TRIALS, rank, world)
        ps = sweep.PointStats(MAX_IT)
        ps.add(it[lo:hi], res[lo:hi])
        stats[pt] = ps.vec
    total = sweep.allreduce_stats(stats)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), total)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_trials():
    for total in (1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [sweep.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_allreduce_matches_single_process_and_oracle(tmp_path, oracle, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    reduced = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    for r in reduced[1:]:
        assert (r == reduced[0]).all(), "every rank must hold the same reduced statistics"
    for pt in range(POINTS):
        it, res = synthetic_results(pt)
        single = sweep.PointStats(MAX_IT)
        single.add(it, res)
        assert (single.vec == reduced[0][pt]).all(), "result must not depend on the number of ranks"
        d = sweep.derive(reduced[0][pt], MAX_IT)
        out3 = np.stack([it, res & 1, (res >> 1) & 1], 1).astype(np.uint64)
        want = oracle.point_stats(out3, MAX_IT)
        assert d.n_trials == TRIALS
        assert d.ratio_sp == want["ratio_sp"] and d.ratio_ldpc == want["ratio_ldpc"]
        assert d.it_min == want["min"] and d.it_max == want["max"]
        assert abs(d.mean - want["mean"]) <= 1e-12 * max(1.0, want["mean"])
        assert abs(d.std_dev - want["std_dev"]) <= 1e-9 * max(1.0, want["std_dev"])


def test_binomial_ci():
    lo, hi = sweep.binomial_ci95(0, 100)
    assert lo == 0.0 and 0.03 < hi < 0.04
    lo, hi = sweep.binomial_ci95(50, 100)
    assert 0.39 < lo < 0.41 and 0.59 < hi < 0.61
