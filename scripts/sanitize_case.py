"""Small launches of every decoder family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): the SM-resident
fp32 (both rules) and fp64 (both rules) kernels, the generic kernel's tiers, the streaming decoder in fp32 and fp64 with several
on-device repacks, the key generator and qlb_run_trials. Outcomes are cross-checked between kernels, so a run that the tool
disturbs still has to produce the right answers.     compute-sanitizer --tool racecheck python scripts/sanitize_case.py [small]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from qkd_ldpc_b200 import capi, codes, workload

small = len(sys.argv) > 1 and sys.argv[1] == "small"
mat = codes.load_npz(codes.NORTH_STAR)
code = capi.Code.from_graph(mat)
ctx = capi.Context(0)
A, B, Q = [], [], []
mix = ((0.03, 40), (0.07, 30), (0.085, 40), (0.095, 30)) if small else ((0.03, 120), (0.07, 90), (0.085, 100), (0.095, 60))
for pt, (q, cnt) in enumerate(mix):
    a, b, ex = ctx.generate(mat.n, workload.trial_seeds(100 + pt, cnt), q)
    A.append(a); B.append(b); Q.append(np.full(cnt, ex))
A, B, Q = np.concatenate(A), np.concatenate(B), np.concatenate(Q)
perm = np.random.default_rng(5).permutation(len(Q))
A, B, Q = np.ascontiguousarray(A[perm]), np.ascontiguousarray(B[perm]), Q[perm]
max_it = 24 if small else 40
out = {}
for name, kw in (("resident f64", dict(precision=64)), ("resident f64 fused", dict(precision=64, fast_math=True)),
                 ("stream f64", dict(precision=64, tier=3)), ("stream f64 fused", dict(precision=64, fast_math=True, tier=3)),
                 ("generic f64 tier 1", dict(precision=64, tier=1)), ("generic f64 tier 2", dict(precision=64, tier=2)),
                 ("resident f32", dict(precision=32)), ("resident f32 fast", dict(precision=32, fast_math=True)),
                 ("stream f32", dict(precision=32, tier=3)), ("stream f32 fast", dict(precision=32, fast_math=True, tier=3)),
                 ("generic f32 tier 0", dict(precision=32, tier=0)), ("generic f32 tier 2", dict(precision=32, tier=2))):
    prec = kw.pop("precision")
    out[name] = ctx.reconcile_packed(code, capi.make_params(prec, max_it, 100.0, True, **kw), A, B, Q, want_decoded=True, want_syndrome=True)
    # a data race shows up as run-to-run or block-size-to-block-size differences: every outcome must repeat bit for bit
    for bt in (0, 512, 640):
        again = ctx.reconcile_packed(code, capi.make_params(prec, max_it, 100.0, True, block_threads=bt, **kw), A, B, Q, want_decoded=True, want_syndrome=True)
        assert all((x == y).all() for x, y in zip(out[name], again)), (name, bt)
    print(f"{name:22s} frames {len(Q)}  converged {int((out[name][1] & 1).sum())}  frame-iterations {int(out[name][0].sum())}", flush=True)
for a, b in (("resident f64", "stream f64"), ("resident f64", "generic f64 tier 1"), ("resident f64", "generic f64 tier 2"),
             ("resident f64 fused", "stream f64 fused"), ("resident f32", "stream f32"), ("resident f32 fast", "stream f32 fast")):
    assert all((x == y).all() for x, y in zip(out[a], out[b])), (a, b)
it, res = ctx.run_trials(code, capi.make_params(64, max_it, 100.0, True), workload.trial_seeds(100, 40), 0.03, 0)[:2]
assert ((res & 3) == 3).all()
print("outcomes agree between the kernels")
