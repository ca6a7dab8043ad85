"""Throughput of the on-device key generator and of qlb_run_trials (seeds in, results out)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from qkd_ldpc_b200 import capi, codes
mat = codes.load_npz(codes.NORTH_STAR); code = capi.Code.from_graph(mat); ctx = capi.Context(0)
dev = torch.device("cuda:0")
for frames in (10000, 100000):
    seeds = torch.randint(0, 2**62, (frames,), dtype=torch.int64, device=dev)
    a = torch.zeros((frames, code.words_n), dtype=torch.int32, device=dev); b = torch.zeros_like(a)
    torch.cuda.synchronize()
    for q in (0.03, 0.11):
        for rep in range(2):
            ctx.timer_start(); ctx.generate_device(mat.n, frames, seeds.data_ptr(), q, a.data_ptr(), b.data_ptr()); ms = ctx.timer_stop()
        print(f"generate_device frames={frames} q={q}: {ms:.3f} ms -> {frames/ms*1e3/1e6:.2f} M frames/s", flush=True)
p = capi.make_params(32, 100, 100.0, True, fast_math=True)
hs = np.random.default_rng(1).integers(0, 2**62, 10000).astype(np.uint64)
for q in (0.03, 0.07, 0.09):
    for rep in range(2):
        t = time.perf_counter(); it, res, ex = ctx.run_trials(code, p, hs, q); dt = time.perf_counter() - t
    print(f"run_trials 10000 frames q={q}: {dt*1e3:.2f} ms -> {10000/dt:.0f} frames/s (ok {int((res&1).sum())})", flush=True)
