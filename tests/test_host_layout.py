"""CPU tests of the host logic of libqkdldpc_b200.so: the library loads and exports every declared symbol, matrix
validation, and the device layout tables -- checked by running the kernel's in-place algorithm in numpy over those
tables and comparing with the oracle. No GPU compute is called here."""
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLD, NS, ROOT
from qkd_ldpc_b200 import capi


def test_exports_match_header(lib):
    header = (ROOT / "include" / "qkd_ldpc_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(qlb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/qkd_ldpc_b200.h but not exported"
    assert sorted(capi.EXPORTS) == declared
    assert lib.qlb_version() == int(re.search(r"#define QLB_VERSION (\d+)", header).group(1))


def test_no_cpu_fallback(lib):
    """Without a GPU the context constructor must fail loudly (there is no CPU decode path)."""
    if lib.qlb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.QlbError) as ei:
        capi.Context(0)
    assert "no CPU path" in str(ei.value)


def emulate(code: capi.Code, mat, bob, syn, log_p, max_it=100, thr=100.0, enable_thr=True):
    """The decode kernel's schedule (one in-place message array over physical slots) in numpy, fp64."""
    slot_of_edge, bit_slots, check_order = code.layout()
    n, m = mat.n, mat.m
    none = np.uint32(0xFFFFFFFF)
    msg = np.zeros(code.slots)
    prior = np.where(np.asarray(bob) != 0, -log_p, log_p)
    for a in range(bit_slots.shape[0]):
        ok = bit_slots[a] != none
        msg[bit_slots[a][ok]] = prior[ok]
    rows = [slot_of_edge[mat.row_ptr[j]:mat.row_ptr[j + 1]] for j in range(m)]
    z = np.zeros(n, np.int32)
    with np.errstate(all="ignore"):
        for it in range(max_it):
            for j in range(m):
                t = np.tanh(msg[rows[j]] / 2.0)
                row = -1.0 if syn[j] else 1.0
                for v in t:
                    row *= v
                out = 2.0 * np.arctanh(row / t)
                if enable_thr:
                    out = np.where(out > thr, thr, np.where(out < -thr, -thr, out))
                msg[rows[j]] = out
            total = prior.copy()
            for a in range(bit_slots.shape[0]):
                ok = bit_slots[a] != none
                total[ok] = total[ok] + msg[bit_slots[a][ok]]
            z = (total <= 0).astype(np.int32)
            for a in range(bit_slots.shape[0]):
                ok = bit_slots[a] != none
                v = total[ok] - msg[bit_slots[a][ok]]
                if enable_thr:
                    v = np.where(v > thr, thr, np.where(v < -thr, -thr, v))
                msg[bit_slots[a][ok]] = v
            par = np.array([np.bitwise_xor.reduce(z[mat.col_idx[mat.row_ptr[j]:mat.row_ptr[j + 1]]]) for j in range(m)])
            if (par == syn).all():
                return it + 1, True, z
    return max_it, False, z


@pytest.mark.parametrize("name", ["dense_n6_m4", "dense_n7_m3", "dense_n10_m5"])
def test_layout_emulation_small_codes(matrices, graphs, dev_codes, oracle, name):
    mat, g, code = matrices[name], graphs[name], dev_codes[name]
    rng = np.random.default_rng(3)
    for _ in range(40):
        a = rng.integers(0, 2, g.n).astype(np.int32)
        b = a.copy()
        b[rng.integers(0, g.n)] ^= 1
        q = 1.0 / g.n
        want = oracle.qkd_ldpc(g, a, b, q)
        got = emulate(code, mat, b, want[3], np.log((1 - q) / q))
        assert got[0] == want[0] and got[1] == want[1] and (got[2] == want[4]).all()


def test_layout_tables_north_star(matrices, dev_codes):
    mat, code = matrices[NS], dev_codes[NS]
    slot_of_edge, bit_slots, check_order = code.layout()
    assert len(set(slot_of_edge.tolist())) == mat.e and slot_of_edge.max() < code.slots, "edges must own distinct slots"
    assert sorted(check_order.tolist()) == list(range(mat.m))
    w = np.diff(mat.row_ptr)[check_order]
    assert (np.diff(w) <= 0).all(), "checks must be sorted by descending weight"
    # slot of (check at sorted position p, edge position k) = base[k] + p
    pos = np.empty(mat.m, np.int64)
    pos[check_order] = np.arange(mat.m)
    cnt = np.array([(np.diff(mat.row_ptr) > k).sum() for k in range(code.max_check_w)])
    base = np.concatenate([[0], np.cumsum((cnt + 31) // 32 * 32)[:-1]])   # rows of slots start on 32-slot boundaries
    naive, placed = code.gather_wavefronts()
    assert placed <= 0.72 * naive and placed <= 2.1, (naive, placed)            # bank-aware placement of the checks (4- and 8-byte objective)
    for j in (0, 1, 700, mat.m - 1):
        for k, p in enumerate(range(mat.row_ptr[j], mat.row_ptr[j + 1])):
            assert slot_of_edge[p] == base[k] + pos[j]
    # bit i's slots are exactly the slots of the edges that touch it, in ascending check order (the reference's
    # arrival order for sorted lists)
    edge_bit = mat.col_idx
    edge_check = np.repeat(np.arange(mat.m), np.diff(mat.row_ptr))
    for i in (0, 17, 5000, mat.n - 1):
        es = np.flatnonzero(edge_bit == i)
        es = es[np.argsort(edge_check[es], kind="stable")]
        assert bit_slots[: len(es), i].tolist() == slot_of_edge[es].tolist()


def test_layout_emulation_north_star_one_frame(matrices, graphs, dev_codes, oracle):
    mat, g, code = matrices[NS], graphs[NS], dev_codes[NS]
    z = np.load(GOLD / "frames_n10240.npz")
    k = 0
    a = capi.unpack_bits(z["alice"][k:k + 1].view(np.uint32), g.n)[0]
    b = capi.unpack_bits(z["bob"][k:k + 1].view(np.uint32), g.n)[0]
    q = float(z["q_exact"][k])
    syn = oracle.syndrome(g, a)
    got = emulate(code, mat, b, syn, np.log((1 - q) / q), max_it=6)
    assert got[0] == int(z["iterations"][k]) and got[1]
    assert (got[2] == capi.unpack_bits(z["decoded"][k:k + 1].view(np.uint32), g.n)[0]).all()


def test_code_create_rejects_bad_matrices(matrices):
    m = matrices["dense_n7_m3"]
    rp, ci, cp, ri = m.row_ptr.copy(), m.col_idx.copy(), m.col_ptr.copy(), m.row_idx.copy()
    with pytest.raises(capi.QlbError):
        capi.Code(m.n, m.m, rp, ci, cp[:-1].tolist() + [cp[-1] - 1], ri)  # edge counts differ
    bad = ci.copy(); bad[0] = m.n
    with pytest.raises(capi.QlbError):
        capi.Code(m.n, m.m, rp, bad, cp, ri)                               # index out of range
    # unsorted check list: the reference's positional routing would misroute -> rejected, not silently "fixed"
    sw = ci.copy(); sw[[0, 1]] = sw[[1, 0]]
    with pytest.raises(capi.QlbError) as ei:
        capi.Code(m.n, m.m, rp, sw, cp, ri)
    assert "consistent" in str(ei.value)
    with pytest.raises(capi.QlbError):
        capi.Code(0, m.m, rp, ci, cp, ri)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 100, 10240):
        b = rng.integers(0, 2, (3, n)).astype(np.int32)
        w = capi.pack_bits(b)
        assert w.shape == (3, (n + 31) // 32)
        assert (capi.unpack_bits(w, n) == b).all()


def test_committed_peg_code_n100k_is_what_it_says():
    """data/codes/peg_n100000_*.npz (configs[3], used by bench.py's stream_n100k variant and the N = 100 000 GPU tests) is the PEG
    generator's own output: column weight 3 everywhere, check weights within one of each other, both adjacency halves sorted,
    describing the same duplicate-free edge set, and accepted by qlb_code_create."""
    from qkd_ldpc_b200 import capi, codes
    mat = codes.peg_code(100000, 51080, 3, 666, bfs_limit=2000)
    assert (mat.n, mat.m, mat.e) == (100000, 51080, 300000)
    bw, cw = np.diff(mat.col_ptr), np.diff(mat.row_ptr)
    assert (bw == 3).all() and cw.max() - cw.min() <= 1 and set(np.unique(cw)) <= {5, 6}
    rows_of_edge = np.repeat(np.arange(mat.m), cw)          # CSR half: (check, bit)
    cols_of_edge = np.repeat(np.arange(mat.n), bw)          # CSC half: (bit, check)
    a = np.unique(rows_of_edge.astype(np.int64) * mat.n + mat.col_idx)
    b = np.unique(mat.row_idx.astype(np.int64) * mat.n + cols_of_edge)
    assert a.size == mat.e and (a == b).all()
    assert all((np.diff(mat.col_idx[mat.row_ptr[j]:mat.row_ptr[j + 1]]) > 0).all() for j in range(0, mat.m, 997))
    code = capi.Code.from_graph(mat)  # host-side layout only: no GPU needed
    assert code.n == mat.n and code.m == mat.m
