"""Quick device-resident throughput probe (not the bench contract): per QBER point, per precision."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from qkd_ldpc_b200 import capi, codes, workload

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f32", "f32fast", "f64"]
qs = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.03, 0.07, 0.09]
mat = codes.load_npz(codes.NORTH_STAR)
code = capi.Code.from_graph(mat)
ctx = capi.Context(0)
dev = torch.device("cuda:0")
for q in qs:
    a, b, qe = workload.make_frames(mat.n, code.words_n, frames, q, 1234, dev)
    lp = torch.full((frames,), workload.log_prior(qe), dtype=torch.float64, device=dev)
    it = torch.zeros(frames, dtype=torch.int32, device=dev)
    res = torch.zeros(frames, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for v in variants:
        p = capi.make_params(64 if v.startswith("f64") else 32, 100, 100.0, True, fast_math=v in ("f32fast", "f64fused"))
        nf = frames if not v.startswith("f64") else max(frames // 4, 148)
        for rep in range(2):
            ctx.timer_start()
            ctx.reconcile_device(code, p, nf, a.data_ptr(), b.data_ptr(), lp.data_ptr(), it.data_ptr(), res.data_ptr())
            ms = ctx.timer_stop()
        iters = int(it[:nf].sum().item()); ok = int((res[:nf] & 1).sum().item()); km = int(((res[:nf] >> 1) & 1).sum().item())
        fi = iters / (ms * 1e-3)
        bytes_per = 16 if not v.startswith("f64") else 32
        print(f"q={q:.3f} {v:8s} frames={nf} ms={ms:9.3f} frames/s={nf/(ms*1e-3):12.1f} mean_it={iters/nf:6.2f} "
              f"frame-it/s={fi/1e6:8.3f}M  alg GB/s={fi*mat.e*bytes_per/1e9:9.1f} ({fi*mat.e*bytes_per/6537.3e9:.3f} of HBM peak) ok={ok} keys={km}",
              flush=True)
