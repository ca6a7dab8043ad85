"""On-device key generator throughput against the batch size (one thread replays one trial: small batches leave the GPU empty)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from qkd_ldpc_b200 import capi, codes
mat = codes.load_npz(codes.NORTH_STAR); code = capi.Code.from_graph(mat); ctx = capi.Context(0)
dev = torch.device("cuda:0")
for frames in (4096, 16384, 32768, 52428, 65536, 131072, 262144):
    seeds = torch.randint(0, 2**62, (frames,), dtype=torch.int64, device=dev)
    a = torch.zeros((frames, code.words_n), dtype=torch.int32, device=dev); b = torch.zeros_like(a)
    torch.cuda.synchronize()
    for q in (0.03,):
        for rep in range(3):
            ctx.timer_start(); ctx.generate_device(mat.n, frames, seeds.data_ptr(), q, a.data_ptr(), b.data_ptr()); ms = ctx.timer_stop()
        print(f"generate_device frames={frames} q={q}: {ms:.3f} ms -> {frames/ms*1e3/1e6:.2f} M frames/s", flush=True)
