"""BASELINE.json configs[4]: multi-rate sweep R = 0.3 ... 0.8 at N = 10240, CW = 3 (seeded PEG codes), fp32 vs fp64 messages,
efficiency f = (1-R)/h2(q) and FER against the reference (CPU oracle on the same reference-generator frames).
usage: multirate_sweep.py <frames_per_point>"""
import json, math, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from oracle.bindings import Graph, Restatement
from qkd_ldpc_b200 import capi, codes

per = int(sys.argv[1]) if len(sys.argv) > 1 else 256
threads = os.cpu_count() or 8
orc = Restatement(); ctx = capi.Context(0)
h2 = lambda q: -q * math.log2(q) - (1 - q) * math.log2(1 - q)
plan = {7168: [0.10, 0.12, 0.13, 0.14, 0.16], 5231: [0.06, 0.075, 0.0825, 0.0875, 0.095], 3072: [0.03, 0.038, 0.042, 0.046, 0.055], 2048: [0.012, 0.018, 0.021, 0.024, 0.03]}
variants = {"f64": capi.make_params(64, 100, 100.0, True), "f64fused": capi.make_params(64, 100, 100.0, True, fast_math=True), "f32": capi.make_params(32, 100, 100.0, True), "f32fast": capi.make_params(32, 100, 100.0, True, fast_math=True)}
seeds0 = orc.trial_seeds(424242, per)
for m, qs in plan.items():
    mat = codes.peg_code(10240, m, 3, 666)
    g = Graph(mat.n, mat.m, mat.row_ptr, mat.col_idx, mat.col_ptr, mat.row_idx)
    code = capi.Code.from_graph(mat)
    for pt, q in enumerate(qs):
        seeds = seeds0 + np.uint64(7 * pt + m)
        want = orc.run_trials(g, q, seeds, threads=threads)
        a, b, exact = ctx.generate(mat.n, seeds, q)
        row = {"M": m, "rate": round(1 - m / mat.n, 4), "q": q, "q_exact": exact, "f_efficiency": round((m / mat.n) / h2(exact), 4), "fer_ref": float(1 - (want[:, 1] * want[:, 2]).mean()),
               "mean_it_ref": float(want[:, 0].mean())}
        for name, p in variants.items():
            for rep in range(2):
                t = time.perf_counter(); it, res, dec, _ = ctx.reconcile_packed(code, p, a, b, np.full(per, exact), want_decoded=False); dt = time.perf_counter() - t
            same = ((res & 1) == want[:, 1]) & (((res >> 1) & 1) == want[:, 2])
            row[name] = {"fer": float(1 - ((res & 3) == 3).mean()), "same_flags": int(same.sum()), "same_it": int((it == want[:, 0]).sum()), "frames_per_s_e2e": round(per / dt)}
        print(json.dumps(row), flush=True)
