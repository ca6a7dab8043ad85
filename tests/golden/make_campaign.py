"""Generates tests/golden/campaign_n10240.npz from the UNMODIFIED reference (oracle/_ref/libqkdref.so): per-frame outcomes of
the reference's own run_trial (src/simulation.cpp:161-189) for the frames of BASELINE.json configs[1]/[2].

Run in the build container only (needs /root/reference for oracle/_ref):

    python tests/golden/make_campaign.py [frames_per_point=4096] [threads]

Frames are the reference's: trial k of point `pt` is seeded with seeds[k] + pt, seeds = the first raw draws of
Xoshiro256PlusPlus(777) (src/simulation.cpp:222-228,247) -- so frame (pt, k) here IS trial k of QBER point pt of the config2
sweep, and the GPU side re-creates the very same keys from the same seeds with its bit-exact on-device generator.
  grid      : the 9 points 0.03 ... 0.11 of configs[1], point index pt = 0..8 (seed offset pt, as curr_sim in the sweep)
  waterfall : 0.0825, 0.085, 0.0875 -- where frames converge late or fail -- with seed offsets 100, 101, 102
Stored per frame: iterations_num (uint8; max_it = 100) and flags (bit 0 syndromes_match, bit 1 keys_match); ~2 bytes per frame
before compression. The reference discards the decoded key (src/qkd_ldpc_algorithm.cpp:444); keys_match pins it for every
converged frame.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.bindings import Reference  # noqa: E402
from qkd_ldpc_b200 import codes  # noqa: E402

GRID = [0.03 + 0.01 * j for j in range(9)]  # src/simulation.cpp:55-61 for {0.03, 0.12, 0.01}
WATERFALL = [0.0825, 0.085, 0.0875]
SEED = 777


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    ref = Reference(max_it=100, thr=100.0, enable_thr=True, threads=threads)
    h = ref.load(codes.materialize()[codes.NORTH_STAR], dense=False)
    seeds = ref.trial_seeds(SEED, per)
    points = [(q, pt) for pt, q in enumerate(GRID)] + [(q, 100 + j) for j, q in enumerate(WATERFALL)]
    its, flags = [], []
    for q, off in points:
        t = time.time()
        out = ref.run_trials(h, q, seeds + np.uint64(off), threads=threads)
        its.append(out[:, 0].astype(np.uint8))
        flags.append((out[:, 1] | (out[:, 2] << np.uint64(1))).astype(np.uint8))
        print(f"q={q:.4f} offset={off}: {per} frames in {time.time() - t:.1f} s, success {int(out[:, 1].sum())}, "
              f"mean iterations {out[:, 0].mean():.2f}", flush=True)
    np.savez_compressed(ROOT / "tests" / "golden" / "campaign_n10240.npz", simulation_seed=SEED, frames_per_point=per,
                        qber=np.array([p[0] for p in points]), seed_offset=np.array([p[1] for p in points], np.uint64),
                        iterations=np.stack(its), flags=np.stack(flags), max_it=100, thr=100.0)


if __name__ == "__main__":
    main()
