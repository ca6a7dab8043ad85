"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/libqkdref.so).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  data/codes/*.npz                     the reference's shipped parity-check matrices, both adjacency halves as its
                                       own loaders produce them (input data, not code) -- the GPU box has no
                                       /root/reference, so the codes travel in this compact form
  tests/golden/frames_n10240.npz       full frames (packed Alice/Bob/syndrome/decoded bits + results) on the N=10240 code
  tests/golden/waterfall_n10240.npz    seeds + reference results for many frames around the waterfall (inputs are
                                       regenerated from the seeds by the pinned restatement generator)
  tests/golden/small_codes.npz         exhaustive single-error cases on the dense N=6/7/10 codes
  tests/golden/kat_n6.json             the textbook example shipped in example/qkd_ldpc_example.cpp:34-39
  tests/golden/sweep_*.csv             CSVs written by the reference's own main() (oracle/_ref/QKD_LDPC_ref)
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.bindings import REF_MAIN, REFERENCE_ROOT, Reference, Restatement, fnv1a64_bits  # noqa: E402

GOLD = ROOT / "tests" / "golden"
CODES = ROOT / "data" / "codes"
ALIST = REFERENCE_ROOT / "alist_sparse_matrices" / "(N=10240,M=5231,R=0.49,CW=3,SEED=666).txt"
DENSE = {
    "dense_n6_m4": "(N=6,K=2,M=4,R=0.34).txt",
    "dense_n7_m3": "(N=7,K=4,M=3,R=0.57).txt",
    "dense_n10_m5": "(N=10,K=5,M=5,R=0.5).txt",
}
THREADS = os.cpu_count() or 1


def save_code(ref, name, path, dense):
    h = ref.load(path, dense)
    g = ref.graph(h)
    idx_t = np.uint16 if max(g.n, g.m) < 65536 else np.uint32
    np.savez_compressed(
        CODES / f"{name}.npz",
        n=g.n, m=g.m, is_regular=g.is_regular, max_bit_w=g.max_bit_w, max_check_w=g.max_check_w,
        check_w=np.diff(g.row_ptr).astype(np.uint8), col_idx=g.col_idx.astype(idx_t),
        bit_w=np.diff(g.col_ptr).astype(np.uint8), row_idx=g.row_idx.astype(idx_t),
        source=str(Path(path).name), dense=bool(dense),
    )
    return h, g


def run_ref_main(config: dict, matrix_files: list[Path], dense: bool) -> str:
    """Runs the reference's own main() in a scratch SOURCE_DIR and returns the CSV it wrote."""
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "config.json").write_text(json.dumps(config))
        sub = td / ("dense_matrices" if dense else "alist_sparse_matrices")
        sub.mkdir()
        for f in matrix_files:
            shutil.copy(f, sub / f.name)
        subprocess.run([str(REF_MAIN)], cwd=td, check=True, capture_output=True)
        out = sorted((td / "results").glob("*.csv"))
        assert len(out) == 1
        return out[0].read_text()


def base_config(**kw):
    cfg = {
        "threads_number": THREADS, "trials_number": 64, "use_config_simulation_seed": True, "simulation_seed": 777,
        "interactive_mode": False, "sum_product_max_iterations": 100, "use_dense_matrices": False,
        "trace_qkd_ldpc": False, "trace_sum_product": False, "trace_sum_product_llr": False,
        "enable_sum_product_msg_llr_threshold": True, "sum_product_msg_llr_threshold": 100.0,
        "code_rate_QBER_parameters": [{"code_rate": 0.5, "QBER_begin": 0.03, "QBER_end": 0.12, "QBER_step": 0.01}],
    }
    cfg.update(kw)
    return cfg


def main():
    CODES.mkdir(parents=True, exist_ok=True)
    ref = Reference(max_it=100, thr=100.0, enable_thr=True)
    orc = Restatement()

    # ---- codes -------------------------------------------------------------------------------------------------
    h, g = save_code(ref, "n10240_m5231_cw3_seed666", ALIST, False)
    small = {}
    for name, fn in DENSE.items():
        small[name] = save_code(ref, name, REFERENCE_ROOT / "dense_matrices" / fn, True)

    # ---- full frames on the N=10240 code -------------------------------------------------------------------------
    seeds = ref.trial_seeds(777, 8)
    spec = [(0.03, 0), (0.03, 1), (0.04, 2), (0.05, 0), (0.06, 0), (0.07, 0), (0.07, 3), (0.08, 0), (0.08, 1),
            (0.0825, 4), (0.085, 0), (0.085, 1), (0.0875, 5), (0.0875, 6), (0.09, 0), (0.09, 1), (0.10, 2), (0.11, 0),
            (0.11, 1)]
    rec = {k: [] for k in ("q_req", "seed", "q_exact", "alice", "bob", "syndrome", "decoded", "iterations",
                            "syndromes_match", "keys_match")}
    for q, k in spec:
        a, b, ex = ref.generate(seeds[k], g.n, q)
        it, sm, km = ref.qkd_ldpc(h, a, b, ex)
        syn = ref.syndrome(h, a, g.m)
        log_p = np.log((1.0 - ex) / ex)
        llr = np.where(b != 0, -log_p, log_p).astype(np.float64)
        it2, sm2, dec = ref.sum_product(h, g.n, llr, syn)
        assert (it, sm) == (it2, sm2) and km == bool((a == dec).all())
        for key, v in zip(rec, (q, seeds[k], ex, np.packbits(a.astype(np.uint8), bitorder="little"),
                                np.packbits(b.astype(np.uint8), bitorder="little"),
                                np.packbits(syn.astype(np.uint8), bitorder="little"),
                                np.packbits(dec.astype(np.uint8), bitorder="little"), it, sm, km)):
            rec[key].append(v)
        print(f"frame q={q} k={k} exact={ex:.7f} it={it} sm={sm} km={km} alice={fnv1a64_bits(a)}")
    np.savez_compressed(
        GOLD / "frames_n10240.npz",
        q_req=np.array(rec["q_req"]), seed=np.array(rec["seed"], np.uint64), q_exact=np.array(rec["q_exact"]),
        alice=np.stack(rec["alice"]), bob=np.stack(rec["bob"]), syndrome=np.stack(rec["syndrome"]),
        decoded=np.stack(rec["decoded"]), iterations=np.array(rec["iterations"], np.int32),
        syndromes_match=np.array(rec["syndromes_match"], np.uint8), keys_match=np.array(rec["keys_match"], np.uint8),
        max_it=100, thr=100.0, enable_thr=True, bitorder="little",
    )

    # ---- waterfall statistics: seeds + results only ------------------------------------------------------------
    wf_q = [0.05, 0.07, 0.08, 0.0825, 0.085, 0.0875, 0.09]
    per = 96
    wf_seeds = np.arange(1000, 1000 + per, dtype=np.uint64)
    wf = {}
    for q in wf_q:
        out3 = ref.run_trials(h, q, wf_seeds, threads=THREADS)
        # decoded-bit hashes come from the pinned restatement (the reference API discards the decoded key);
        # the restatement must agree with the reference on (iterations, flags) for every one of these frames
        o3, dec = orc.run_trials(g, q, wf_seeds, threads=THREADS, want_decoded=True)
        assert (o3 == out3).all(), f"restatement != reference at q={q}"
        wf[f"res_{q}"] = out3.astype(np.int32)
        wf[f"dechash_{q}"] = np.array([int(fnv1a64_bits(d), 16) for d in dec], np.uint64)  # FNV-1a, one byte per bit
        print(f"waterfall q={q}: ok {int(out3[:, 1].sum())}/{per}, mean it {out3[out3[:, 1] == 1, 0].mean() if out3[:, 1].any() else 0:.2f}")
    np.savez_compressed(GOLD / "waterfall_n10240.npz", q=np.array(wf_q), seeds=wf_seeds, max_it=100, thr=100.0, **wf)

    # ---- exhaustive small codes: every Alice x every single-bit error ----------------------------------------------
    sm_out = {}
    for name, (hh, gg) in small.items():
        n = gg.n
        q = 1.0 / n
        rows = []
        for av in range(1 << n):
            a = np.array([(av >> i) & 1 for i in range(n)], np.int32)
            for epos in range(n):
                b = a.copy()
                b[epos] ^= 1
                for variant in ((0, 1) if gg.is_regular else (0,)):
                    it, s, k = ref.qkd_ldpc(hh, a, b, q, variant=variant)
                    syn = ref.syndrome(hh, a, gg.m, variant=variant)
                    llr = np.where(b != 0, -np.log((1 - q) / q), np.log((1 - q) / q))
                    _, _, dec = ref.sum_product(hh, n, llr, syn, variant=variant)
                    dv = int(sum(int(v) << i for i, v in enumerate(dec)))
                    sv = int(sum(int(v) << i for i, v in enumerate(syn)))
                    rows.append((av, epos, variant, it, int(s), int(k), dv, sv))
        sm_out[name] = np.array(rows, np.int32)
        print(name, "cases", len(rows), "success", int(sm_out[name][:, 5].sum()))
    np.savez_compressed(GOLD / "small_codes.npz", **sm_out)

    # ---- textbook KAT -------------------------------------------------------------------------------------------
    hh, gg = small["dense_n6_m4"]
    a = np.array([0, 0, 1, 0, 1, 1], np.int32)
    b = np.array([1, 0, 1, 0, 1, 1], np.int32)
    it, s, k = ref.qkd_ldpc(hh, a, b, 0.2, variant=1)
    kat = {
        "source": "example/qkd_ldpc_example.cpp:34-39 (Johnson, Introducing LDPC codes, example 2.5); values in "
                  "'survey_trace' are the reference's own TRACE output recorded in SURVEY.md section 4",
        "alice": a.tolist(), "bob": b.tolist(), "qber": 0.2, "variant": "regular",
        "iterations": it, "syndromes_match": s, "keys_match": k,
        "survey_trace": {"r_abs": 1.386, "E_abs": 0.7538, "L": [0.1212, 1.386, -2.894, 1.386, -1.386, -1.386],
                         "z": [0, 0, 1, 0, 1, 1], "s": [0, 0, 0, 0]},
    }
    (GOLD / "kat_n6.json").write_text(json.dumps(kat, indent=1))
    print("KAT", it, s, k)

    # ---- the reference's own main(): CSV goldens -------------------------------------------------------------------
    csv = run_ref_main(base_config(trials_number=64), [ALIST], dense=False)
    (GOLD / "sweep_n10240_t64_seed777.csv").write_text(csv)
    print(csv)
    cfg = base_config(trials_number=1000, use_dense_matrices=True,
                      code_rate_QBER_parameters=[{"code_rate": 0.58, "QBER_begin": 0.15, "QBER_end": 0.35, "QBER_step": 0.1}])
    csv = run_ref_main(cfg, [REFERENCE_ROOT / "dense_matrices" / DENSE["dense_n7_m3"]], dense=True)
    (GOLD / "sweep_dense_n7_t1000_seed777.csv").write_text(csv)
    print(csv)


if __name__ == "__main__":
    main()
