"""BASELINE.json configs[3] and configs[4] as parity cases: seeded PEG codes of other rates (R = 0.3 ... 0.8 at N = 10240,
check weights 4 ... 16) and of larger block length (N = 100 000, messages in global scratch), against the oracle."""
import numpy as np
import pytest

from qkd_ldpc_b200 import capi, codes

pytestmark = pytest.mark.gpu


def graph_of(mat):
    from oracle.bindings import Graph
    return Graph(mat.n, mat.m, mat.row_ptr, mat.col_idx, mat.col_ptr, mat.row_idx)


def frames_for(oracle, n, q, seeds):
    ab = [oracle.generate(int(s), n, q) for s in seeds]
    return np.stack([x[0] for x in ab]), np.stack([x[1] for x in ab]), np.array([x[2] for x in ab])


@pytest.mark.parametrize("m,qs", [(7168, (0.10, 0.13, 0.17)), (5231, (0.06, 0.085)), (3072, (0.03, 0.045, 0.06)), (2048, (0.012, 0.02, 0.03))])
def test_multi_rate_codes_match_oracle(ctx, oracle, m, qs):
    n = 10240
    mat = codes.peg_code(n, m, 3, 666)
    g, code = graph_of(mat), capi.Code.from_graph(mat)
    seeds = oracle.trial_seeds(4242, 12)
    seen = set()
    for pt, q in enumerate(qs):
        s = seeds + np.uint64(pt)
        want, wdec = oracle.run_trials(g, q, s, threads=8, want_decoded=True)
        A, B, Q = frames_for(oracle, n, q, s)
        it, res, dec, _ = ctx.reconcile_packed(code, capi.make_params(64, 100, 100.0, True), capi.pack_bits(A), capi.pack_bits(B), Q)
        assert (it == want[:, 0]).all(), (m, q, it, want[:, 0])
        assert ((res & 1) == want[:, 1]).all() and (((res >> 1) & 1) == want[:, 2]).all()
        assert (capi.unpack_bits(dec, n) == wdec).all()
        seen.update(want[:, 1].tolist())
        for fast in (False, True):
            it32, res32, dec32, _ = ctx.reconcile_packed(code, capi.make_params(32, 100, 100.0, True, fast_math=fast), capi.pack_bits(A), capi.pack_bits(B), Q)
            same = ((res32 & 1) == want[:, 1]) & (((res32 >> 1) & 1) == want[:, 2])
            assert same.mean() >= 0.9, (m, q, fast, same)
            ok = (want[:, 1] == 1) & same
            assert (capi.unpack_bits(dec32, n)[ok] == wdec[ok]).all()
    assert seen == {0, 1}, "the QBER points must straddle the code's threshold"


def test_large_block_length_matches_oracle(ctx, oracle):
    """N = 100 000 (E = 300 000): messages, indices and hard decisions live in per-CTA global scratch (tier 2)."""
    n, m = 100000, 51080
    mat = codes.peg_code(n, m, 3, 666, bfs_limit=2000)
    g, code = graph_of(mat), capi.Code.from_graph(mat)
    seeds = oracle.trial_seeds(99, 3)
    for q, mi in ((0.05, 100), (0.10, 12)):
        want, wdec = oracle.run_trials(g, q, seeds, threads=3, max_it=mi, want_decoded=True)
        A, B, Q = frames_for(oracle, n, q, seeds)
        for precision in (64, 32):
            it, res, dec, syn = ctx.reconcile_packed(code, capi.make_params(precision, mi, 100.0, True), capi.pack_bits(A), capi.pack_bits(B), Q,
                                                     want_syndrome=True)
            want_syn = np.stack([oracle.syndrome(g, a) for a in A])
            assert (capi.unpack_bits(syn, m) == want_syn).all()
            assert ((res & 1) == want[:, 1]).all() and (((res >> 1) & 1) == want[:, 2]).all()
            if precision == 64:
                assert (it == want[:, 0]).all()
                assert (capi.unpack_bits(dec, n) == wdec).all()
