#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

Metric   : decoded frames/s (and sifted-key Mbit/s = frames/s * N / 1e6) over the QBER sweep 0.03 ... 0.11 of the
           N=10240, M=5231, CW=3 code (configs[1]); FER per point reported beside it.
A "step" : one pass of the hot path over one batch = the whole 9-point sweep, `--frames-per-point` frames per point
           per GPU (default 10 000, as configs[1]), every frame through the fused reconcile kernel
           (Alice syndrome + LLR init + sum-product decode with per-frame early termination + key compare).
value    : whole-job frames/s with the packed keys already resident in HBM (device-timed, max over ranks).
e2e      : the same sweep through the C-ABI call a user makes (qlb_reconcile_batch_packed) with pinned HOST buffers:
           H2D of keys + QBERs and D2H of results inside the timed region.
roofline : SURVEY.md 8d: 16 B (fp32) / 32 B (fp64) of algorithmic message traffic per edge-iteration, against the measured
           HBM copy bandwidth in MEASURED_PEAKS.json. The fp32 kernel keeps a frame's messages in shared memory, so
           its DRAM traffic is far below the algorithmic bytes and `frac` may exceed 1; `roofline.secondary`
           is the roof that does bound it (instruction issue). `variants.stream_n100k*` are the HBM-bound design point
           (configs[3]) measured in the same run.
cpu_baseline / --impl reference : the reference's own CPU implementation (oracle/_ref, the unmodified sources compiled in
           place) on the box's host cores, on a bounded sample of the same sweep.

Launch: `python bench.py --gpus 1 ...` or, for N > 1,
`python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...`
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MAX_IT = 100
THR = 100.0
BYTES_PER_EDGE_IT = {"f32": 16, "f32fast": 16, "f64": 32, "f64fused": 32}


def qber_grid():
    return [0.03 + 0.01 * j for j in range(9)]  # end-exclusive grid of src/simulation.cpp:55-61 for 0.03..0.12 step 0.01


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in Path(self.tmp.name).read_text().splitlines() if r.count(",") >= 8]
        os.unlink(self.tmp.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline
def cpu_reference_sweep(trials_per_point: int, threads: int, seed: int = 777):
    """The reference's CPU path on this host: run_trial (generate + QKD_LDPC_irregular) for `trials_per_point` trials
    at each of the 9 QBER points, `threads` host threads. Returns (kind, seconds, per-point out3 arrays)."""
    from oracle.bindings import REF_SO, Reference, Restatement
    from qkd_ldpc_b200 import codes
    mats = codes.materialize()
    grid = qber_grid()
    if REF_SO.exists():
        ref = Reference(max_it=MAX_IT, thr=THR, enable_thr=True, threads=threads)
        h = ref.load(mats[codes.NORTH_STAR], dense=False)
        seeds = ref.trial_seeds(seed, trials_per_point)
        run = lambda q, s: ref.run_trials(h, q, s, threads=threads)
        kind = "reference"
    else:
        from oracle.bindings import Graph
        orc = Restatement()
        m = codes.load_npz(codes.NORTH_STAR)
        g = Graph(m.n, m.m, m.row_ptr, m.col_idx, m.col_ptr, m.row_idx)
        seeds = orc.trial_seeds(seed, trials_per_point)
        run = lambda q, s: orc.run_trials(g, q, s, threads=threads)
        kind = "port"
    outs = []
    t0 = time.perf_counter()
    for pt, q in enumerate(grid):
        outs.append(run(q, seeds + np.uint64(pt)))  # trial seed = seeds[k] + curr_sim (src/simulation.cpp:247)
    return kind, time.perf_counter() - t0, outs


def run_reference_arm(args, guard):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    tpp = args.cpu_trials_per_point or max(64, 4 * cores)  # ~2 s wall = ~30 core-seconds per sweep
    grid = qber_grid()
    for _ in range(args.warmup):
        cpu_reference_sweep(max(1, tpp // 8), cores)
    times = []
    for _ in range(args.steps):
        kind, sec, outs = cpu_reference_sweep(tpp, cores)
        times.append(sec)
    frames = tpp * len(grid)
    sec = float(np.mean(times))
    value = frames / sec
    fer = [1.0 - float((o[:, 1] * o[:, 2]).mean()) for o in outs]
    sample = f"{tpp} trials per QBER point x {len(grid)} points per step (run_trial: key generation + reconciliation)"
    line = {
        "impl": "reference", "metric": "decoded_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.frames_per_point, args.precision),  # the measured arm's config; the sample is below
        "sifted_mbit_s": value * 10240 / 1e6,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "per_qber": [{"qber": q, "fer": f} for q, f in zip(grid, fer)],
        "gpu_launches": 0,
    }
    guard.emit(line)
    return 0


def workload_config(frames_per_point, precision):
    return {
        "workload": "configs[1]: alist (N=10240,M=5231,R=0.49,CW=3,SEED=666) QBER sweep 0.03-0.11 (9 points), "
                    f"{frames_per_point} frames per point per GPU",
        "code": "N=10240 M=5231 E=30720 CW=3", "qber_grid": [round(q, 4) for q in qber_grid()],
        "frames_per_point_per_gpu": frames_per_point, "max_iterations": MAX_IT, "msg_threshold": THR,
        "precision": precision, "parallelism": "trial-sharded (one process per GPU, no data-path collective); the 9 launches of a sweep alternate between two streams of the GPU",
        "l2_policy": "inputs larger than L2 (packed keys of one sweep > 126 MB) and a fresh key set per QBER point",
    }


# ---------------------------------------------------------------------------------------------------------------------
class StdoutGuard:
    """The contract is ONE JSON line on stdout: anything libraries print to fd 1 meanwhile (e.g. NCCL's version banner) is
    sent to stderr, and the line is written to the real stdout at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: dict):
        sys.stdout.flush()
        os.write(self.real, (json.dumps(line) + "\n").encode())


def main():
    guard = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-point", type=int, default=10000)
    ap.add_argument("--precision", default="f32fast", choices=["f32", "f32fast", "f64", "f64fused"])
    ap.add_argument("--cpu-trials-per-point", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)

    if args.impl == "reference":
        return run_reference_arm(args, guard)

    import torch
    import torch.distributed as dist
    from qkd_ldpc_b200 import capi, codes, sweep, workload

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    mat = codes.load_npz(codes.NORTH_STAR)
    code = capi.Code.from_graph(mat)
    ctx = capi.Context(local_rank)
    grid = qber_grid()
    fpp = args.frames_per_point
    n, e = mat.n, mat.e

    def params_for(prec):
        return capi.make_params(64 if prec.startswith("f64") else 32, MAX_IT, THR, True, fast_math=prec in ("f32fast", "f64fused"))

    # ---- synthetic inputs, resident in HBM (each rank draws its own keys) ------------------------------------------
    keys = []
    for pt, q in enumerate(grid):
        a, b, qe = workload.make_frames(n, code.words_n, fpp, q, 1000 * (rank + 1) + pt, dev)
        lp = torch.full((fpp,), workload.log_prior(qe), dtype=torch.float64, device=dev)
        keys.append((a, b, lp, qe))
    d_it = torch.zeros((len(grid), fpp), dtype=torch.int32, device=dev)
    d_res = torch.zeros((len(grid), fpp), dtype=torch.uint8, device=dev)
    h_it = torch.zeros((len(grid), fpp), dtype=torch.int32).pin_memory()
    h_res = torch.zeros((len(grid), fpp), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # a second context (stream, frame queue) on the same GPU: the points of a sweep alternate between the two streams, so the
    # tail of one point's launch (SMs running out of frames) is filled by the next point's CTAs. Fork / join through events,
    # so that the CUDA events timed() records on ctx's stream still bracket everything.
    ctx_b = capi.Context(local_rank)
    ext_b = torch.cuda.ExternalStream(ctx_b.stream, device=dev)
    launch_order = sorted(range(len(grid)), key=lambda pt: -grid[pt])  # the long (non-converging) points first

    def sweep_step(prec, frames=fpp, collect=True):
        """One step on device-resident inputs: 9 launches over the two streams, then results -> host -> statistics."""
        p = params_for(prec)
        fork = torch.cuda.Event()
        fork.record(ext)
        ext_b.wait_event(fork)
        for i, pt in enumerate(launch_order):
            a, b, lp, _ = keys[pt]
            (ctx if i % 2 == 0 else ctx_b).reconcile_device(code, p, frames, a.data_ptr(), b.data_ptr(), lp.data_ptr(), d_it[pt].data_ptr(), d_res[pt].data_ptr())
        join = torch.cuda.Event()
        join.record(ext_b)
        ext.wait_event(join)
        if not collect:
            return None
        with torch.cuda.stream(ext):
            h_it.copy_(d_it, non_blocking=True)
            h_res.copy_(d_res, non_blocking=True)
        ctx.synchronize()
        stats = np.zeros((len(grid), MAX_IT + 5), np.int64)
        for pt in range(len(grid)):
            ps = sweep.PointStats(MAX_IT)
            ps.add(h_it[pt, :frames].numpy(), h_res[pt, :frames].numpy())
            stats[pt] = ps.vec
        return sweep.allreduce_stats(stats, dev)  # the sweep's only collective (NCCL when world > 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Device time of `steps` calls: CUDA events on the launching stream, bracketed by barrier + synchronize."""
        barrier()
        ctx.timer_start()
        t0 = time.perf_counter()
        out = None
        for _ in range(steps):
            out = fn()
        ms = ctx.timer_stop()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall, out

    prec = args.precision
    for _ in range(args.warmup):
        sweep_step(prec)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.counters(reset=True)
    ctx_b.counters(reset=True)
    ms_total, wall_total, stats = timed(lambda: sweep_step(prec), args.steps)
    launches, frame_iters = (x + y for x, y in zip(ctx.counters(reset=True), ctx_b.counters(reset=True)))
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    frames_step = fpp * len(grid) * world
    value = frames_step / (ms_step * 1e-3)
    results = [sweep.derive(stats[pt], MAX_IT) for pt in range(len(grid))]

    # ---- per-point kernel durations (events tightly around each launch) -> roofline ----------------------------------
    def per_point(precision, frames):
        p = params_for(precision)
        out = []
        for pt in range(len(grid)):
            a, b, lp, _ = keys[pt]
            best = None
            for _ in range(2):
                ctx.timer_start()
                ctx.reconcile_device(code, p, frames, a.data_ptr(), b.data_ptr(), lp.data_ptr(), d_it[pt].data_ptr(), d_res[pt].data_ptr())
                ms = ctx.timer_stop()
                best = ms if best is None else min(best, ms)
            iters = int(d_it[pt, :frames].sum().item())
            ok = int(((d_res[pt, :frames] & 3) == 3).sum().item())
            out.append((best, iters, ok, frames))
        return out

    hbm_peak, peak_src = measured_peaks()

    def roofline_of(pp, precision):
        kern_ms = sum(x[0] for x in pp)
        alg_bytes = sum(x[1] for x in pp) * e * BYTES_PER_EDGE_IT[precision]
        ach = alg_bytes / (kern_ms * 1e-3) / 1e9
        return kern_ms, alg_bytes, ach

    pp = per_point(prec, fpp)
    kern_ms, alg_bytes, ach = roofline_of(pp, prec)
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(prec, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
        "peak_source": peak_src, "kernel": ("qlb::decode_resident_f64_kernel" if prec.startswith("f64") else "qlb::decode_resident_f32_kernel") + " (fused reconcile: prior init + Alice syndrome + BP iterations + early termination + key compare)",
        "algorithmic_bytes_per_edge_iteration": BYTES_PER_EDGE_IT[prec],
        "frame_iterations_per_s": sum(x[1] for x in pp) / (kern_ms * 1e-3),
        "edge_iterations_per_s": sum(x[1] for x in pp) * e / (kern_ms * 1e-3),
        "kernel_ms_per_step": kern_ms, "kernel_share_of_step": kern_ms / ms_step,
        "note": "fp32 messages live in shared memory: DRAM traffic << algorithmic bytes, so frac can exceed 1" if not prec.startswith("f64")
                else "fp64 messages live in shared memory (92 %) + a small L2-resident per-CTA tail; the kernel is FP64-pipe bound",
    }
    # the resident kernels never touch HBM inside an iteration: what bounds them is instruction issue. Secondary roof from the
    # ncu-measured warp instructions per frame-iteration (profiles/traffic.json) against 4 issue slots per SM per clock.
    try:
        wi = json.loads(tp.read_text()).get(prec, {}).get("warp_instructions_per_frame_iteration")
    except Exception:
        wi = None
    if wi:
        sm_mhz = float(clocks.get("sm_mhz") or 0.0) if isinstance(clocks, dict) else 0.0
        issue_peak = 4.0 * ctx.sm_count * sm_mhz * 1e6 if sm_mhz > 0 else None
        issue_ach = roofline["frame_iterations_per_s"] * wi
        roofline["secondary"] = {"bound": "issue", "unit": "G warp-instructions/s", "achieved": issue_ach / 1e9,
                                 "peak": issue_peak / 1e9 if issue_peak else None, "frac": issue_ach / issue_peak if issue_peak else None,
                                 "warp_instructions_per_frame_iteration": wi,
                                 "source": "ncu smsp__inst_executed.sum of the profiled launch / its frame-iterations; peak = 4 x SMs x median SM clock under load"}
    per_qber = []
    for pt, q in enumerate(grid):
        ms_pt, iters, ok, fr = pp[pt]
        r = results[pt]
        per_qber.append({"qber": round(q, 4), "qber_exact": keys[pt][3], "frames_per_s": fr / (ms_pt * 1e-3),
                         "sifted_mbit_s": fr / (ms_pt * 1e-3) * n / 1e6, "mean_iterations": iters / fr, "fer": r.fer,
                         "mean_iterations_successful": r.mean, "roofline_frac": iters * e * BYTES_PER_EDGE_IT[prec] / (ms_pt * 1e-3) / 1e9 / hbm_peak})

    # ---- end to end through the C-ABI with pinned host buffers -------------------------------------------------------
    h_keys = []
    for pt in range(len(grid)):
        a, b, lp, qe = keys[pt]
        h_keys.append((a.cpu().pin_memory(), b.cpu().pin_memory(), torch.full((fpp,), qe, dtype=torch.float64).pin_memory()))
    p_e2e = params_for(prec)

    # Two host threads, each with a context (stream + staging buffers) of its own, take alternate QBER points -- what the C++
    # scheduler does with its two workers per GPU: one point's H2D / D2H copies run under the other point's decode. The calls
    # are the blocking C-ABI entry point; ctypes releases the GIL while they run.
    from concurrent.futures import ThreadPoolExecutor
    e2e_pool = ThreadPoolExecutor(max_workers=2)

    def e2e_points(c, points):
        for pt in points:
            ha, hb, hq = h_keys[pt]
            c.reconcile_packed_ptrs(code, p_e2e, fpp, ha.data_ptr(), hb.data_ptr(), hq.data_ptr(), h_it[pt].data_ptr(), h_res[pt].data_ptr())

    def e2e_step():
        # longest points first on each worker so that the two finish together
        order = sorted(range(len(grid)), key=lambda pt: -grid[pt])
        jobs = [e2e_pool.submit(e2e_points, ctx, order[0::2]), e2e_pool.submit(e2e_points, ctx_b, order[1::2])]
        for j in jobs:
            j.result()
        return float(h_it.sum())  # the step's result is read on the host

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    e2e_ms, e2e_wall, _ = timed(e2e_step, args.steps)
    e2e_value = frames_step / (e2e_wall / args.steps * 1e-3)
    h2d = len(grid) * fpp * (2 * code.words_n * 4 + 8)
    d2h = len(grid) * fpp * 5
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_wall / args.steps, "sifted_mbit_s": e2e_value * n / 1e6,
           "api": "qlb_reconcile_batch_packed (pinned host buffers), two host threads / contexts per GPU taking alternate QBER points"}

    # ---- the other precisions, shorter (explanatory numbers, same JSON line) -----------------------------------------
    variants = {}
    if not args.no_variants and rank == 0:
        for v in ("f32", "f32fast", "f64", "f64fused"):
            if v == prec:
                continue
            fr = fpp if not v.startswith("f64") else max(148, fpp // 4)
            for pt in (0, 6):
                ctx.reconcile_device(code, params_for(v), min(fr, 1024), keys[pt][0].data_ptr(), keys[pt][1].data_ptr(), keys[pt][2].data_ptr(),
                                     d_it[pt].data_ptr(), d_res[pt].data_ptr())
            ppv = per_point(v, fr)
            kms, ab, achv = roofline_of(ppv, v)
            variants[v] = {"frames_per_point": fr, "frames_per_s": fr * len(grid) / (kms * 1e-3), "achieved_GBps": achv, "roofline_frac": achv / hbm_peak,
                           "frame_iterations_per_s": sum(x[1] for x in ppv) / (kms * 1e-3),
                           "fer": [1.0 - x[2] / x[3] for x in ppv]}
    # ---- the HBM-bound design point (configs[3]): N = 100 000 through the streaming decoder, same JSON line ---------------------
    if not args.no_variants and rank == 0 and world == 1:
        try:
            big = codes.peg_code(100000, 51080, 3, 666, bfs_limit=2000)  # committed copy under data/codes/
            big_code = capi.Code.from_graph(big)
            fr, q_big, it_big = 18944, 0.10, 20  # 148 groups of 128 frames; nothing converges at this QBER: 20 full iterations
            ba, bb, bq = workload.make_frames(big.n, big_code.words_n, fr, q_big, 4242, dev, chunk=max(1, 2 ** 26 // big.n))
            blp = torch.full((fr,), workload.log_prior(bq), dtype=torch.float64, device=dev)
            bit_ = torch.zeros(fr, dtype=torch.int32, device=dev)
            bres = torch.zeros(fr, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            pbig = capi.make_params(32, it_big, THR, True, fast_math=True)
            best = None
            for _ in range(5):
                ctx.timer_start()
                ctx.reconcile_device(big_code, pbig, fr, ba.data_ptr(), bb.data_ptr(), blp.data_ptr(), bit_.data_ptr(), bres.data_ptr())
                ms = ctx.timer_stop()
                best = ms if best is None else min(best, ms)
            its = int(bit_.sum().item())
            gbs = its * big.e * 16 / (best * 1e-3) / 1e9
            variants["stream_n100k"] = {
                "workload": "configs[3]: PEG N=100000 M=51080 CW=3 SEED=666, 18944 frames, QBER 0.10, 20 iterations, fp32 fast rule",
                "kernels": "qlb::stream_{setup,init,check,update,bit,finalize}_kernel (one kernel per pass, 4-group bundles)",
                "ms": best, "frame_iterations_per_s": its / (best * 1e-3), "edge_iterations_per_s": its * big.e / (best * 1e-3),
                "achieved_GBps": gbs, "roofline_frac": gbs / hbm_peak, "bound": "hbm",
                "note": "whole call incl. set-up and result kernels, CUDA events; DRAM bytes measured by ncu = algorithmic bytes (profiles/r01_stream_split.md)"}
            # the same code at a waterfall QBER: most frames converge around round 45, ~7 % run to 100 -- what the on-device
            # frame compaction is for (algorithmic bytes count only the rounds each frame needed)
            wa, wb, wq = workload.make_frames(big.n, big_code.words_n, fr, 0.085, 4343, dev, chunk=max(1, 2 ** 26 // big.n))
            blp.fill_(workload.log_prior(wq))
            pw = capi.make_params(32, MAX_IT, THR, True, fast_math=True)
            best = None
            for _ in range(2):
                ctx.timer_start()
                ctx.reconcile_device(big_code, pw, fr, wa.data_ptr(), wb.data_ptr(), blp.data_ptr(), bit_.data_ptr(), bres.data_ptr())
                ms = ctx.timer_stop()
                best = ms if best is None else min(best, ms)
            its = int(bit_.sum().item())
            gbs = its * big.e * 16 / (best * 1e-3) / 1e9
            variants["stream_n100k_waterfall"] = {
                "workload": "same code, QBER 0.085, max 100 iterations, 18944 frames", "ms": best, "mean_iterations": its / fr,
                "fer": 1.0 - float(((bres & 3) == 3).sum().item()) / fr, "frame_iterations_per_s": its / (best * 1e-3),
                "achieved_GBps": gbs, "roofline_frac": gbs / hbm_peak, "bound": "hbm"}
            del ba, bb, wa, wb, blp, bit_, bres, big_code
        except Exception as ex:  # the headline line must not depend on the side measurement
            variants["stream_n100k"] = {"error": str(ex)[:200]}
    if world > 1:
        dist.barrier()

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------------------------
    cpu = None
    fer_parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        tpp = args.cpu_trials_per_point or max(64, 4 * cores)  # ~2 s wall = ~30 core-seconds per sweep
        kind, sec, outs = cpu_reference_sweep(tpp, cores)
        cpu_val = tpp * len(grid) / sec
        cpu = {"value": cpu_val, "unit": "frames/s", "cores": cores, "kind": kind, "seconds": sec,
               "sample": f"{tpp} trials per QBER point x {len(grid)} points (run_trial: key generation + reconciliation), {cores} threads",
               "per_qber_fer": [1.0 - float((o[:, 1] * o[:, 2]).mean()) for o in outs],
               "per_qber_mean_iterations": [float(o[:, 0].mean()) for o in outs]}
        fer_parity = []
        for pt in range(len(grid)):
            k_ref = int(tpp - (outs[pt][:, 1] * outs[pt][:, 2]).sum())
            lo, hi = sweep.binomial_ci95(k_ref, tpp)
            fer_parity.append({"qber": round(grid[pt], 4), "fer_gpu": results[pt].fer, "fer_ref": k_ref / tpp, "ref_ci95": [lo, hi],
                               "inside": bool(lo <= results[pt].fer <= hi)})

    if rank == 0:
        line = {
            "metric": "decoded_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if prec.startswith("f64") else "f32", "data": "synthetic",
            "config": workload_config(fpp, prec),
            "sifted_mbit_s": value * n / 1e6,
            "frame_iterations_per_step": int(frame_iters // args.steps) * 1, "wall_ms_per_step": wall_total / args.steps,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "per_qber": per_qber, "fer_parity_vs_cpu_sample": fer_parity, "variants": variants,
        }
        guard.emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
