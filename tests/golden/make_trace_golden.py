"""Console-trace fixtures from the reference's own main() (oracle/_ref/QKD_LDPC_ref, unmodified sources):
stdout of small batch / interactive runs with trace_qkd_ldpc / trace_sum_product / trace_sum_product_llr enabled.
Run in the build container (needs /root/reference): python tests/golden/make_trace_golden.py

  tests/golden/trace/<case>.json        config.json of the run, the matrix file name, stdin
  tests/golden/trace/<case>.stdout.gz   the reference's stdout (or .sha256 + head for the N=10240 case)
  tests/golden/trace/peg_n96_m48.alist  the sparse test matrix (seeded PEG), as fed to the reference

tests/test_trace.py replays each case through qkd_ldpc_b200_sim on the GPU and compares the blue (trace) part byte for byte.
"""
import gzip
import hashlib
import json
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.bindings import REF_MAIN, REFERENCE_ROOT  # noqa: E402
from qkd_ldpc_b200 import codes  # noqa: E402

OUT = ROOT / "tests" / "golden" / "trace"
BLUE = "\x1b[38;2;000;000;255m"
RESET = "\x1b[0m"


def trace_body(stdout: str) -> str:
    """From the first blue print to the end of the last one (what the TRACE_* keys add to the console)."""
    a = stdout.find(BLUE)
    if a < 0:
        return ""
    b = stdout.rfind(BLUE)
    return stdout[a:stdout.index(RESET, b) + len(RESET)]


def config(**kw):
    cfg = {
        "threads_number": 1, "trials_number": 2, "use_config_simulation_seed": True, "simulation_seed": 777,
        "interactive_mode": False, "sum_product_max_iterations": 10, "use_dense_matrices": True,
        "trace_qkd_ldpc": True, "trace_sum_product": True, "trace_sum_product_llr": True,
        "enable_sum_product_msg_llr_threshold": True, "sum_product_msg_llr_threshold": 100.0,
        "code_rate_QBER_parameters": [{"code_rate": 0.99, "QBER_begin": 0.15, "QBER_end": 0.3, "QBER_step": 0.1}],
    }
    cfg.update(kw)
    return cfg


def run(case: str, cfg: dict, matrix: Path, stdin: str = "", digest_only: bool = False):
    dense = cfg["use_dense_matrices"]
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "config.json").write_text(json.dumps(cfg))
        sub = td / ("dense_matrices" if dense else "alist_sparse_matrices")
        sub.mkdir()
        shutil.copy(matrix, sub / matrix.name)
        r = subprocess.run([str(REF_MAIN)], cwd=td, capture_output=True, input=stdin.encode())
        if r.returncode != 0:
            raise RuntimeError(f"{case}: reference main failed: {r.stderr.decode()[-400:]}")
    out = r.stdout.decode()
    meta = {"config": cfg, "matrix": matrix.name, "stdin": stdin, "stdout_bytes": len(out)}
    if digest_only:
        body = trace_body(out)  # MAX_LLR is printed with 17 digits: kept apart, compared numerically
        meta["max_llr"] = [float(v) for v in re.findall(r"MAX_LLR = (\S+)", body)]
        body = re.sub(r"MAX_LLR = \S+", "MAX_LLR = #", body)
        meta["trace_sha256"] = hashlib.sha256(body.encode()).hexdigest()
        meta["trace_bytes_blanked"] = len(body)
        meta["trace_head"] = body[:4000]
    else:
        with open(OUT / f"{case}.stdout.gz", "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", compresslevel=9, mtime=0) as f:
            f.write(out.encode())
    (OUT / f"{case}.json").write_text(json.dumps(meta, indent=1))
    print(case, "stdout", len(out), "bytes; trace", len(trace_body(out)))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    dense = REFERENCE_ROOT / "dense_matrices"
    n7 = dense / "(N=7,K=4,M=3,R=0.57).txt"
    n10 = dense / "(N=10,K=5,M=5,R=0.5).txt"
    n6 = dense / "(N=6,K=2,M=4,R=0.34).txt"
    run("batch_dense_n7_full", config(), n7)
    run("batch_dense_n6_regular_full", config(code_rate_QBER_parameters=[{"code_rate": 0.99, "QBER_begin": 0.2, "QBER_end": 0.25, "QBER_step": 0.05}]), n6)
    run("batch_dense_n10_llr_only", config(trace_sum_product=False, sum_product_max_iterations=100, trials_number=3,
                                           code_rate_QBER_parameters=[{"code_rate": 0.99, "QBER_begin": 0.1, "QBER_end": 0.35, "QBER_step": 0.1}]), n10)
    run("batch_dense_n10_noclamp", config(enable_sum_product_msg_llr_threshold=False, sum_product_max_iterations=30, trials_number=2,
                                          code_rate_QBER_parameters=[{"code_rate": 0.99, "QBER_begin": 0.1, "QBER_end": 0.15, "QBER_step": 0.05}]), n10)
    peg = OUT / "peg_n96_m48.alist"
    codes.write_alist(codes.peg_code(96, 48, 3, seed=5), peg)
    run("batch_alist_n96_full", config(use_dense_matrices=False, sum_product_max_iterations=20,
                                       code_rate_QBER_parameters=[{"code_rate": 0.99, "QBER_begin": 0.05, "QBER_end": 0.06, "QBER_step": 0.01}]), peg)
    run("interactive_dense_n7", config(interactive_mode=True, trace_sum_product=False, trace_sum_product_llr=False), n7, stdin="1\n")
    run("interactive_alist_n96_full", config(interactive_mode=True, use_dense_matrices=False, sum_product_max_iterations=20,
                                             code_rate_QBER_parameters=[{"code_rate": 0.99, "QBER_begin": 0.03, "QBER_end": 0.08, "QBER_step": 0.02}]),
        peg, stdin="1\n")
    big = REFERENCE_ROOT / "alist_sparse_matrices" / "(N=10240,M=5231,R=0.49,CW=3,SEED=666).txt"
    run("batch_alist_n10240_full", config(use_dense_matrices=False, trials_number=1, sum_product_max_iterations=100,
                                          code_rate_QBER_parameters=[{"code_rate": 0.99, "QBER_begin": 0.03, "QBER_end": 0.04, "QBER_step": 0.01}]),
        big, digest_only=True)


if __name__ == "__main__":
    main()
