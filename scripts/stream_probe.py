"""Throughput of the decoders on a code of block length n (PEG up to N = 100 000, permutation code above; the streaming
kernels when a frame does not fit one SM).
usage: stream_probe.py <n> <m> <frames> <qber> <max_it> [fast|acc|f64|f64fused] [tier]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from qkd_ldpc_b200 import capi, codes, workload
n, m, frames, q, max_it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
rule = sys.argv[6] if len(sys.argv) > 6 else "fast"
fast = rule in ("fast", "f64fused")
prec, bpe = (64, 32) if rule.startswith("f64") else (32, 16)
tier = int(sys.argv[7]) if len(sys.argv) > 7 else None
mat = codes.load_npz(codes.NORTH_STAR) if (n, m) == (10240, 5231) else (codes.peg_code(n, m, 3, 666, bfs_limit=2000) if n <= 100000 else codes.permutation_code(n, m, 3, 666))
code = capi.Code.from_graph(mat); ctx = capi.Context(0); dev = torch.device("cuda:0")
a, b, qe = workload.make_frames(mat.n, code.words_n, frames, q, 7, dev, chunk=max(1, 2**26 // mat.n))
lp = torch.full((frames,), workload.log_prior(qe), dtype=torch.float64, device=dev)
it = torch.zeros(frames, dtype=torch.int32, device=dev); res = torch.zeros(frames, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
p = capi.make_params(prec, max_it, 100.0, True, fast_math=fast, tier=tier)
for rep in range(3):
    ctx.timer_start(); ctx.reconcile_device(code, p, frames, a.data_ptr(), b.data_ptr(), lp.data_ptr(), it.data_ptr(), res.data_ptr()); ms = ctx.timer_stop()
    iters = int(it.sum().item()); ok = int((res & 1).sum().item())
    gbs = iters * mat.e * bpe / (ms * 1e-3) / 1e9
    print(f"n={n} frames={frames} q={q} max_it={max_it} rule={rule} tier={tier}: {ms:.2f} ms, mean it {iters/frames:.2f}, ok {ok}, {iters/(ms*1e-3)/1e6:.3f} M frame-it/s, "
          f"{iters*mat.e/(ms*1e-3)/1e9:.1f} G edge-it/s, algorithmic {gbs:.0f} GB/s = {gbs/6537.3:.3f} of HBM peak", flush=True)
