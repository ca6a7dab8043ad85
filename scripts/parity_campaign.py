"""Large parity campaign on a GPU box: reference-generator frames around the waterfall and on the benchmark grid, decoded by
the CPU oracle (pinned restatement of the reference, fp64) and by every GPU variant; prints per-variant mismatch counts.
usage: parity_campaign.py <frames_per_point> [threads]"""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from oracle.bindings import Graph, Restatement
from qkd_ldpc_b200 import capi, codes

per = int(sys.argv[1]) if len(sys.argv) > 1 else 256
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
mat = codes.load_npz(codes.NORTH_STAR)
g = Graph(mat.n, mat.m, mat.row_ptr, mat.col_idx, mat.col_ptr, mat.row_idx)
orc = Restatement(); code = capi.Code.from_graph(mat); ctx = capi.Context(0)
qs = [0.03, 0.05, 0.07, 0.08, 0.0825, 0.085, 0.0875, 0.09, 0.11]
variants = {"f64": capi.make_params(64, 100, 100.0, True), "f64fused": capi.make_params(64, 100, 100.0, True, fast_math=True), "f32": capi.make_params(32, 100, 100.0, True),
            "f32fast": capi.make_params(32, 100, 100.0, True, fast_math=True),
            "f32fast_stream": capi.make_params(32, 100, 100.0, True, fast_math=True, tier=3)}
tot = {v: dict(frames=0, flags=0, iters=0, keys_conv=0, conv=0, it1=0, bits_all=0) for v in variants}
rows = []
seeds0 = orc.trial_seeds(20261018, per)
for pt, q in enumerate(qs):
    seeds = seeds0 + np.uint64(1000 * pt)
    t = time.time(); want, wdec = orc.run_trials(g, q, seeds, threads=threads, want_decoded=True); t_cpu = time.time() - t
    a, b, exact = ctx.generate(mat.n, seeds, q)
    row = {"q": q, "exact": exact, "ref_success": int(want[:, 1].sum()), "cpu_s": round(t_cpu, 1)}
    for name, p in variants.items():
        it, res, dec, _ = ctx.reconcile_packed(code, p, a, b, np.full(per, exact))
        same_flags = ((res & 1) == want[:, 1]) & (((res >> 1) & 1) == want[:, 2])
        same_it = it == want[:, 0]
        conv = (want[:, 1] == 1) & same_flags
        keys = (capi.unpack_bits(dec, mat.n)[conv] == wdec[conv]).all(axis=1) if conv.any() else np.zeros(0, bool)
        d = tot[name]
        d["frames"] += per; d["flags"] += int(same_flags.sum()); d["iters"] += int(same_it.sum()); d["conv"] += int(conv.sum())
        d["bits_all"] += int((capi.unpack_bits(dec, mat.n) == wdec).all(axis=1).sum())  # incl. the last decision of failed frames
        d["keys_conv"] += int(keys.sum()); d["it1"] += int((np.abs(it.astype(int) - want[:, 0].astype(int)) <= 1).sum())
        row[name] = {"same_flags": int(same_flags.sum()), "same_iterations": int(same_it.sum()), "fer": float(1 - ((res & 3) == 3).mean())}
    row["fer_ref"] = float(1 - (want[:, 1] * want[:, 2]).mean())
    rows.append(row)
    print(json.dumps(row), flush=True)
print("TOTALS", json.dumps(tot))
