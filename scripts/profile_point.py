"""One launch of the bench's own shape for ncu: `frames` reference frames (trial seeds of Xoshiro256PlusPlus(777) + point index, keys
from the on-device generator) of ONE QBER point of configs[1], one precision. A warm launch first (ncu: --launch-skip accordingly).

    python scripts/profile_point.py f64 0.09 10000 [block_threads]
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from qkd_ldpc_b200 import capi, codes, workload

prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
q = float(sys.argv[2]) if len(sys.argv) > 2 else 0.09
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
bt = int(sys.argv[4]) if len(sys.argv) > 4 else 0
grid = [0.03 + 0.01 * j for j in range(9)]
pt = min(range(9), key=lambda j: abs(grid[j] - q))
mat = codes.load_npz(codes.NORTH_STAR)
code = capi.Code.from_graph(mat)
ctx = capi.Context(0)
dev = torch.device("cuda:0")
seeds = torch.from_numpy(workload.trial_seeds(777, frames).view(np.int64)).to(dev)
a = torch.empty((frames, code.words_n), dtype=torch.int32, device=dev)
b = torch.empty_like(a)
qe = ctx.generate_device(mat.n, frames, seeds.data_ptr(), q, a.data_ptr(), b.data_ptr(), seed_offset=pt)
lp = torch.full((frames,), workload.log_prior(qe), dtype=torch.float64, device=dev)
it = torch.zeros(frames, dtype=torch.int32, device=dev)
res = torch.zeros(frames, dtype=torch.uint8, device=dev)
p = capi.make_params(64 if prec.startswith("f64") else 32, 100, 100.0, True, fast_math=prec in ("f32fast", "f64fused"), block_threads=bt)
for rep in range(2):
    ctx.timer_start()
    ctx.reconcile_device(code, p, frames, a.data_ptr(), b.data_ptr(), lp.data_ptr(), it.data_ptr(), res.data_ptr())
    ms = ctx.timer_stop()
iters = int(it.sum().item())
bpe = 32 if prec.startswith("f64") else 16
print(f"{prec} q={q} frames={frames} ms={ms:.3f} frame-iterations={iters} frame-it/s={iters / ms / 1e3:.3f}M algorithmic bytes={iters * mat.e * bpe} "
      f"({iters * mat.e * bpe / ms / 1e6 / 6537.3:.3f} of HBM peak) ok={int((res & 1).sum().item())}")
