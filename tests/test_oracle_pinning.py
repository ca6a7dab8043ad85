"""The trust chain of the parity tests, on the CPU: Restatement (oracle/restatement, plain C) == Reference (the unmodified
reference sources compiled in place, oracle/_ref) == the committed golden fixtures. The Reference tests skip where oracle/_ref
is not built (it needs /root/reference once; the built library then travels with the repo snapshot)."""
import numpy as np
import pytest

from conftest import GOLD, NS
from qkd_ldpc_b200 import codes


@pytest.fixture(scope="module")
def campaign():
    z = np.load(GOLD / "campaign_n10240.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def ref_h(reference):
    return reference.load(codes.materialize()[NS], dense=False)


def test_campaign_fixture_shape(campaign):
    """12 points x 4096 frames of the reference's own run_trial outcomes (tests/golden/make_campaign.py)."""
    per = int(campaign["frames_per_point"])
    assert campaign["iterations"].shape == (12, per) and campaign["flags"].shape == (12, per)
    assert np.allclose(campaign["qber"][:9], [0.03 + 0.01 * j for j in range(9)]) and campaign["seed_offset"][:9].tolist() == list(range(9))
    ok = campaign["flags"] & 1
    assert ok[:5].all() and not ok[7:9].any()           # 0.03 ... 0.07 always converge, 0.10 / 0.11 never
    assert 0 < ok[10].sum() < per                        # the waterfall points straddle
    assert ((campaign["flags"] >> 1) & 1 <= ok).all()    # keys can only match when the syndromes do
    assert (campaign["iterations"][ok == 0] == int(campaign["max_it"])).all()


def test_restatement_equals_fixture(oracle, graphs, campaign):
    """The plain-C restatement reproduces the reference's outcomes on the first frames of every campaign point."""
    g = graphs[NS]
    seeds = oracle.trial_seeds(int(campaign["simulation_seed"]), 12)
    for pt in range(12):
        k = 12 if campaign["qber"][pt] < 0.0825 else 4  # failing frames cost 100 iterations each
        out = oracle.run_trials(g, float(campaign["qber"][pt]), seeds[:k] + campaign["seed_offset"][pt], threads=4)
        assert (out[:, 0] == campaign["iterations"][pt, :k]).all(), pt
        assert ((out[:, 1] | (out[:, 2] << 1)) == campaign["flags"][pt, :k]).all(), pt


def test_reference_equals_fixture_and_restatement(reference, ref_h, oracle, graphs, campaign):
    """Live: the unmodified reference's run_trial (src/simulation.cpp:161-189) against the committed fixture and against the
    restatement, frame by frame, including the generator (same seeds -> same keys)."""
    g = graphs[NS]
    seeds = reference.trial_seeds(int(campaign["simulation_seed"]), 8)
    assert (seeds == oracle.trial_seeds(int(campaign["simulation_seed"]), 8)).all()
    for pt in (0, 4, 5, 6, 10):
        q, off = float(campaign["qber"][pt]), campaign["seed_offset"][pt]
        k = 8 if q < 0.0825 else 3
        out = reference.run_trials(ref_h, q, seeds[:k] + off, threads=4)
        assert (out[:, 0] == campaign["iterations"][pt, :k]).all() and ((out[:, 1] | (out[:, 2] << np.uint64(1))) == campaign["flags"][pt, :k]).all()
        assert (out == oracle.run_trials(g, q, seeds[:k] + off, threads=4)).all()
        a, b, ex = reference.generate(int(seeds[0] + off), g.n, q)
        a2, b2, ex2 = oracle.generate(int(seeds[0] + off), g.n, q)
        assert ex == ex2 and (a == a2).all() and (b == b2).all()


@pytest.mark.parametrize("name", ["dense_n6_m4", "dense_n7_m3", "dense_n10_m5"])
def test_reference_regular_and_irregular_entries_on_small_codes(reference, oracle, graphs, name):
    """QKD_LDPC_regular (src/qkd_ldpc_algorithm.cpp:347-396) and QKD_LDPC_irregular (:398-447) of the reference itself against the
    restatement, every single-error pattern of a few keys; on the regular N=6 code the two entries must also agree."""
    g = graphs[name]
    path = codes.materialize()[name]
    h = reference.load(path, dense=True)
    try:
        rng = np.random.default_rng(3)
        for _ in range(6):
            a = rng.integers(0, 2, g.n).astype(np.int32)
            for e in range(g.n):
                b = a.copy(); b[e] ^= 1
                variants = (0, 1) if g.is_regular else (0,)
                got = [reference.qkd_ldpc(h, a, b, 1.0 / g.n, variant=v) for v in variants]
                want = oracle.qkd_ldpc(g, a, b, 1.0 / g.n)
                assert all(r == (want[0], want[1], want[2]) for r in got), (name, a, e, got, want[:3])
    finally:
        reference.free(h)
