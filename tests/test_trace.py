"""the reference's console traces (CFG.TRACE_QKD_LDPC / TRACE_SUM_PRODUCT / TRACE_SUM_PRODUCT_LLR) and its
interactive mode, reproduced with the GPU decoder's own intermediates.

Golden data: stdout of the reference's unmodified main() (tests/golden/trace/*, made by tests/golden/make_trace_golden.py).
CPU: the oracle's trace copy-out is pinned against that text. GPU: qkd_ldpc_b200_sim replays every case and must print the
same bytes (the values are printed with 4 significant digits; MAX_LLR, printed with 17, is compared to 1e-9), and
qlb_sum_product_trace must agree with the oracle's intermediates and with the throughput kernels' results."""
import gzip
import hashlib
import json
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLD, NS
from qkd_ldpc_b200 import build, capi, codes

TRACE = GOLD / "trace"
BLUE = "\x1b[38;2;000;000;255m"
RESET = "\x1b[0m"
ANSI = re.compile(r"\x1b\[[0-9;]*m")
CASES = sorted(p.stem for p in TRACE.glob("*.json"))
DENSE_BY_FILE = {"(N=6,K=2,M=4,R=0.34).txt": "dense_n6_m4", "(N=7,K=4,M=3,R=0.57).txt": "dense_n7_m3", "(N=10,K=5,M=5,R=0.5).txt": "dense_n10_m5"}


def trace_body(stdout: str) -> str:
    a = stdout.find(BLUE)
    if a < 0:
        return ""
    b = stdout.rfind(BLUE)
    return stdout[a:stdout.index(RESET, b) + len(RESET)]


def split_max_llr(text: str):
    """-> (text with the 17-digit MAX_LLR values blanked, the values)."""
    vals = [float(v) for v in re.findall(r"MAX_LLR = (\S+)", text)]
    return re.sub(r"MAX_LLR = \S+", "MAX_LLR = #", text), vals


def golden_stdout(case: str) -> str:
    return gzip.open(TRACE / f"{case}.stdout.gz").read().decode()


def matrix_path(meta):
    name = meta["matrix"]
    if name.endswith(".alist"):
        return TRACE / name
    if name in DENSE_BY_FILE:
        return codes.materialize()[DENSE_BY_FILE[name]]
    return codes.materialize()[NS]


# ---- CPU: the oracle's trace against the reference's printed trace --------------------------------------------------
def parse_frames(stdout: str):
    """Splits a full-trace stdout into frames -> list of dict(r, syndrome, iterations=[dict(E, L, z, s, M?)], bits, performed)."""
    plain = ANSI.sub("", stdout)
    frames = []
    for chunk in plain.split("\nr:\n")[1:]:
        fr = {"iterations": []}
        head, rest = chunk.split("\n\nAlice syndrome:\n", 1)
        fr["r"] = [float(x) for x in head.split()]
        syn, rest = rest.split("\n\nIteration: ", 1) if "\n\nIteration: " in rest else (rest, "")
        fr["syndrome"] = [int(x) for x in syn.split("\n")[0].split()]
        rest, tail = rest.split("\nBob corrected bit array:\n", 1)
        for it_chunk in rest.split("\n\nIteration: "):
            if not it_chunk.strip():
                continue
            it = {}
            body = it_chunk.split("\n", 1)[1]
            e_txt, body = body.split("\nE:\n", 1)[1].split("\nL:\n", 1)
            it["E"] = [float(x) for x in e_txt.split()]
            l_txt, body = body.split("\n\nz:\n", 1)
            it["L"] = [float(x) for x in l_txt.split()]
            z_txt, body = body.split("\n\ns:\n", 1)
            it["z"] = [int(x) for x in z_txt.split()]
            if "\n\nM:\n" in body:
                s_txt, m_txt = body.split("\n\nM:\n", 1)
                it["M"] = [float(x) for x in m_txt.split("\n\nMAX_LLR")[0].split()]
            else:
                s_txt = body
            it["s"] = [int(x) for x in s_txt.split("\n\nMAX_LLR")[0].split()]
            fr["iterations"].append(it)
        fr["bits"] = [int(x) for x in tail.split("\n\nIterations performed: ")[0].split()]
        fr["performed"] = int(tail.split("\n\nIterations performed: ")[1].split("\n")[0])
        frames.append(fr)
    return frames


def g4(values):
    return [float("%.4g" % v) for v in values]


@pytest.mark.parametrize("case", ["batch_alist_n96_full", "interactive_alist_n96_full", "batch_dense_n7_full", "batch_dense_n10_noclamp"])
def test_oracle_trace_matches_reference_console(oracle, case):
    """orc_sum_product_f64_trace reproduces every number the reference printed (4 significant digits), frame by frame."""
    from oracle.bindings import Graph
    meta = json.loads((TRACE / f"{case}.json").read_text())
    cfg = meta["config"]
    mat = codes.read_dense(matrix_path(meta)) if cfg["use_dense_matrices"] else codes.read_alist(matrix_path(meta))
    g = Graph(mat.n, mat.m, mat.row_ptr, mat.col_idx, mat.col_ptr, mat.row_idx, is_regular=mat.is_regular, max_bit_w=mat.max_bit_w,
              max_check_w=mat.max_check_w)
    frames = parse_frames(golden_stdout(case))
    assert frames
    for fr in frames:
        assert len(fr["r"]) == mat.n and len(fr["syndrome"]) == mat.m
        # the printed prior has 4 digits; its exact value is +-log((1-q)/q) with q = errors/n, recovered from the frame itself
        mag = abs(fr["r"][0])
        q_candidates = [k / mat.n for k in range(1, mat.n) if float("%.4g" % np.log((1 - k / mat.n) / (k / mat.n))) == mag]
        assert len(q_candidates) == 1
        lp = np.log((1 - q_candidates[0]) / q_candidates[0])
        llr = np.where(np.array(fr["r"]) < 0, -lp, lp)
        tr = oracle.sum_product_trace(g, llr, fr["syndrome"], cfg["sum_product_max_iterations"], max_it=cfg["sum_product_max_iterations"],
                                      thr=cfg["sum_product_msg_llr_threshold"], enable_thr=cfg["enable_sum_product_msg_llr_threshold"])
        assert tr["iterations"] == fr["performed"] == len(fr["iterations"])
        assert tr["bits"].tolist() == fr["bits"]
        for t, it in enumerate(fr["iterations"]):
            assert g4(tr["E"][t]) == it["E"] and g4(tr["L"][t]) == it["L"]
            assert tr["z"][t].tolist() == it["z"] and tr["s"][t].tolist() == it["s"]
            if "M" in it:
                assert g4(tr["M"][t]) == it["M"]
            else:
                assert t == len(fr["iterations"]) - 1 and tr["result"] == 1


# ---- GPU: the simulator's console against the reference's -----------------------------------------------------------
@pytest.fixture(scope="module")
def sim():
    build.build_all()
    assert build.SIM_PATH.exists()
    return build.SIM_PATH


def replay(sim, tmp_path, meta, **extra_cfg):
    cfg = dict(meta["config"], **extra_cfg)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    sub = tmp_path / ("dense_matrices" if cfg["use_dense_matrices"] else "alist_sparse_matrices")
    sub.mkdir()
    src = matrix_path(meta)
    shutil.copy(src, sub / meta["matrix"])
    p = subprocess.run([str(sim), str(tmp_path)], capture_output=True, input=meta["stdin"].encode())
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c.startswith("batch") and "n10240" not in c])
def test_batch_trace_equals_reference_console(sim, tmp_path, case):
    meta = json.loads((TRACE / f"{case}.json").read_text())
    mine, mine_llr = split_max_llr(trace_body(replay(sim, tmp_path, meta)))
    ref, ref_llr = split_max_llr(trace_body(golden_stdout(case)))
    assert mine == ref
    assert np.allclose(mine_llr, ref_llr, rtol=1e-9, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c.startswith("interactive")])
def test_interactive_mode_equals_reference_console(sim, tmp_path, case):
    """Interactive mode end to end: the whole stdout (menu, per-frame summary, traces) equals the reference's."""
    meta = json.loads((TRACE / f"{case}.json").read_text())
    mine, mine_llr = split_max_llr(replay(sim, tmp_path, meta))
    ref, ref_llr = split_max_llr(golden_stdout(case))
    assert mine == ref
    assert np.allclose(mine_llr, ref_llr, rtol=1e-9, atol=0)


@pytest.mark.gpu
def test_batch_trace_north_star_digest(sim, tmp_path):
    """One traced frame of the N=10240 code (9.8 MB of console text): digest of the trace equals the reference's."""
    meta = json.loads((TRACE / "batch_alist_n10240_full.json").read_text())
    body, llr = split_max_llr(trace_body(replay(sim, tmp_path, meta)))
    assert len(body) == meta["trace_bytes_blanked"]
    assert body[:4000] == meta["trace_head"]
    assert hashlib.sha256(body.encode()).hexdigest() == meta["trace_sha256"]
    assert np.allclose(llr, meta["max_llr"], rtol=1e-9, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("frame", [0, 7, 13, 14])
def test_trace_api_matches_oracle_and_throughput_kernels(ctx, oracle, matrices, graphs, frame):
    """qlb_sum_product_trace on golden N=10240 frames: intermediates equal the oracle's to 1e-9, decisions exactly; its
    (iterations, result, bits) equal what qlb_sum_product_batch returns."""
    z = np.load(GOLD / "frames_n10240.npz")
    mat, g = matrices[NS], graphs[NS]
    code = capi.Code.from_graph(mat)
    unpack = lambda a, n: np.unpackbits(a, bitorder="little")[:n].astype(np.int32)  # noqa: E731
    bob, syn = unpack(z["bob"][frame], mat.n), unpack(z["syndrome"][frame], mat.m)
    q = float(z["q_exact"][frame])
    lp = np.log((1.0 - q) / q)
    llr = np.where(bob != 0, -lp, lp)
    p = capi.make_params(64, 100, 100.0, True)
    cap = 12
    dev = ctx.sum_product_trace(code, p, llr, syn, cap)
    orc = oracle.sum_product_trace(g, llr, syn, cap)
    it, res, bits = ctx.sum_product(code, p, llr, syn)
    assert dev["iterations"] == orc["iterations"] == int(it[0]) == int(z["iterations"][frame])
    assert dev["result"] == orc["result"] == int(res[0])
    assert (dev["bits"] == bits[0]).all() and (dev["bits"] == orc["bits"]).all()
    done = min(cap, dev["iterations"])
    for key in ("E", "L"):
        assert np.allclose(dev[key][:done], orc[key][:done], rtol=1e-9, atol=1e-9), key
    m_done = done - 1 if (dev["result"] and dev["iterations"] <= cap) else done
    assert np.allclose(dev["M"][:m_done], orc["M"][:m_done], rtol=1e-9, atol=1e-9)
    assert (dev["z"][:done] == orc["z"][:done]).all() and (dev["s"][:done] == orc["s"][:done]).all()
    # internal consistency of the device trace: L = prior + row sums of E; z = (L <= 0); s = H z
    col_ptr = mat.col_ptr
    sums = np.add.reduceat(dev["E"][0], col_ptr[:-1])
    assert np.allclose(dev["L"][0], llr + sums, rtol=1e-12, atol=1e-12)
    assert ((dev["L"][:done] <= 0).astype(np.int32) == dev["z"][:done]).all()
    assert (dev["s"][0] == oracle.syndrome(g, dev["z"][0])).all()
