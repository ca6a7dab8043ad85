"""Synthetic workload for the reconciliation hot path, with the reference's input distribution
(src/array_and_matrix_operations.cpp:424-460): Alice = i.i.d. fair bits, Bob = Alice with EXACTLY floor(N*q) flipped
positions chosen uniformly. torch is used only as plumbing (device memory, RNG, pinned host buffers).

The keys here are drawn from torch's generator, not from the reference's xoshiro stream: throughput does not depend on
which fair bits are drawn. Parity runs (tests/, smoke) use the oracle's generator, which reproduces the reference's keys.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# BASELINE.json configs[1]: QBER sweep 0.03 ... 0.11 (end-exclusive grid 0.03 + 0.01*j, j < 9; src/simulation.cpp:55-61)
QBER_GRID = [0.03 + 0.01 * j for j in range(9)]


def exact_qber(n: int, q: float) -> float:
    return float(int(n * q)) / n


def _pack(bits_u8: torch.Tensor, words: int) -> torch.Tensor:
    """[F][n] uint8 0/1 (device) -> [F][words] int32 holding the uint32 words, bit i at word i//32 position i%32."""
    f, n = bits_u8.shape
    if n != words * 32:
        bits_u8 = torch.nn.functional.pad(bits_u8, (0, words * 32 - n))
    w = torch.tensor([1 << k for k in range(32)], dtype=torch.int64, device=bits_u8.device)
    v = (bits_u8.view(f, words, 32).to(torch.int64) * w).sum(-1)           # 0 .. 2^32-1
    v = torch.where(v >= 2 ** 31, v - 2 ** 32, v)
    return v.to(torch.int32).contiguous()


def make_frames(n: int, words: int, n_frames: int, q: float, seed: int, device, chunk: int = 4096):
    """Returns (alice_packed, bob_packed) int32 [F][words] on `device`, and the exact QBER."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_err = int(n * q)
    if n_err == 0:
        raise RuntimeError(f"Key size '{n}' is too small for QBER.")
    a_out = torch.empty((n_frames, words), dtype=torch.int32, device=device)
    b_out = torch.empty((n_frames, words), dtype=torch.int32, device=device)
    for lo in range(0, n_frames, chunk):
        f = min(chunk, n_frames - lo)
        alice = torch.randint(0, 2, (f, n), dtype=torch.uint8, device=device, generator=gen)
        if n <= 65536:
            pos = torch.rand((f, n), device=device, generator=gen).topk(n_err, dim=1).indices
            flip = torch.zeros((f, n), dtype=torch.uint8, device=device)
            flip.scatter_(1, pos, 1)
        else:  # very long keys (throughput probes only): i.i.d. flips with probability q instead of an exact count
            flip = (torch.rand((f, n), device=device, generator=gen) < (n_err / n)).to(torch.uint8)
        a_out[lo:lo + f] = _pack(alice, words)
        b_out[lo:lo + f] = _pack(alice ^ flip, words)
    return a_out, b_out, n_err / n


def log_prior(q_exact: float) -> float:
    return math.log((1.0 - q_exact) / q_exact)


# ---------------------------------------------------------------------------------------------------------------------
# The reference's trial seeds (src/simulation.cpp:222-228): seeds[k] = k-th raw output of Xoshiro256PlusPlus(simulation_seed)
# (Reputeless/Xoshiro-cpp 1.1: state = four SplitMix64 outputs of the seed). Trial k of sweep point `pt` is seeded with
# seeds[k] + pt (:247); fed to qlb_generate_device / qlb_run_trials these seeds re-create the reference's own frames bit for bit.
_M64 = (1 << 64) - 1


def _rotl(x: int, k: int) -> int:
    return ((x << k) | (x >> (64 - k))) & _M64


def trial_seeds(simulation_seed: int, count: int) -> np.ndarray:
    x = int(simulation_seed) & _M64
    s = []
    for _ in range(4):  # SplitMix64
        x = (x + 0x9E3779B97F4A7C15) & _M64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        s.append(z ^ (z >> 31))
    s0, s1, s2, s3 = s
    out = np.empty(int(count), np.uint64)
    for k in range(int(count)):
        out[k] = (_rotl((s0 + s3) & _M64, 23) + s0) & _M64
        t = (s1 << 17) & _M64
        s2 ^= s0
        s3 ^= s1
        s1 ^= s2
        s0 ^= s3
        s2 ^= t
        s3 = _rotl(s3, 45)
    return out
