"""The fp64 building blocks of the check rule (csrc/qlb_f64_math.cuh) against the host's extended-precision libm, element-wise
through the qlb_test_f64_math probe: their accuracy is a test, not a claim."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulp_err(got, want_ld):
    """|got - want| in units of the last place of `got`'s binade, `want` in long double (64-bit mantissa here)."""
    want = np.asarray(want_ld, np.longdouble)
    ulp = np.spacing(np.abs(want).astype(np.float64)).astype(np.longdouble)
    return np.abs(got.astype(np.longdouble) - want) / ulp


def test_table_driven_exp(ctx):
    """e^-|m|, op 5: Tang's scheme with a 32-entry (hi, lo) table, <= 0.6 ulp measured on the CPU model; here <= 1 ulp on 2e6 arguments
    spread over the magnitudes the decoder produces (messages up to the clamp, 100) and beyond."""
    rng = np.random.default_rng(1)
    m = np.concatenate([rng.uniform(-100, 100, 1_000_000), rng.uniform(-1, 1, 500_000) * 10.0 ** rng.uniform(-12, 0, 500_000),
                        rng.uniform(-690, 690, 500_000)])  # (beyond ~693 the binary exponent is capped by design, see below)
    got = ctx.f64_math(5, m)
    want = np.exp(-np.abs(m.astype(np.longdouble)))
    err = ulp_err(got, want)
    assert err.max() <= 1.0, (err.max(), m[err.argmax()])
    assert (ctx.f64_math(5, np.array([0.0, -0.0])) == 1.0).all()
    big = ctx.f64_math(5, np.array([694.0, 800.0, 1e300, np.inf, -np.inf]))
    assert (big >= 0).all() and (big < 1e-300).all()   # "a positive number far below 2^-54": all tanh(|m| / 2) = 1 needs


def test_division_is_ieee(ctx):
    """op 0: the rule's own division (MUFU.RCP64H seed, one Newton step, exact-residual correction) equals the IEEE quotient."""
    rng = np.random.default_rng(2)
    a = rng.standard_normal(2_000_000) * 10.0 ** rng.uniform(-8, 8, 2_000_000)
    b = rng.standard_normal(2_000_000) * 10.0 ** rng.uniform(-8, 8, 2_000_000)
    assert (ctx.f64_math(0, a, b) == a / b).all()


def test_tanh_half_and_two_atanh(ctx):
    rng = np.random.default_rng(3)
    m = np.concatenate([rng.uniform(-80, 80, 1_000_000), rng.uniform(-1, 1, 500_000) * 10.0 ** rng.uniform(-10, 0, 500_000)])
    got = ctx.f64_math(3, m)
    want = np.tanh(m.astype(np.longdouble) / 2)
    # (1 - e) / (1 + e), e = e^-|m|: a few ulp where tanh is of order 1; for small |m| the subtraction 1 - e cancels and what the form
    # keeps is the ABSOLUTE accuracy of e, i.e. ~1.1e-16 -- unlike libm's tanh, which stays relatively accurate down to 0. Messages
    # that small do not occur in a decode other than as exact zeros (which give exactly 0 here too); the campaigns are the contract.
    abs_err = np.abs(got.astype(np.longdouble) - want)
    tol = np.maximum(2.5 * np.spacing(np.abs(want).astype(np.float64)), 2.3e-16)
    assert (abs_err <= tol).all(), (float((abs_err / tol).max()), m[(abs_err / tol).argmax()])
    assert (ctx.f64_math(3, np.array([0.0])) == 0.0).all()
    assert (np.abs(got) <= 1.0).all() and (np.sign(got) == np.sign(m)).all()
    p = np.concatenate([rng.uniform(-1, 1, 1_000_000), np.sign(rng.standard_normal(500_000)) * (1 - 10.0 ** rng.uniform(-15, 0, 500_000))])
    p = p[np.abs(p) < 1]
    got = ctx.f64_math(4, p, np.zeros_like(p))
    want = 2 * np.arctanh(p.astype(np.longdouble))
    # ln((1 + |p|) / (1 - |p|)): a few ulp for |p| of order 1; for small |p| the roundings of 1 + |p| and 1 - |p| bound the ABSOLUTE
    # error at ~4e-16 (libm's atanh stays relatively accurate instead) -- same remark as for tanh_half above
    abs_err = np.abs(got.astype(np.longdouble) - want)
    tol = np.maximum(3.0 * np.spacing(np.abs(want).astype(np.float64)), 4.5e-16)
    assert (abs_err <= tol).all(), (float((abs_err / tol).max()), p[(abs_err / tol).argmax()])
    inf = ctx.f64_math(4, np.array([1.0, -1.0]), np.array([1.0, 1.0]))  # want_inf: the literal expression's IEEE outcome
    assert np.isposinf(inf[0]) and np.isneginf(inf[1])


def test_log_ratio(ctx):
    rng = np.random.default_rng(4)
    den = 10.0 ** rng.uniform(-16, 0.3, 1_500_000)
    num = den * (1 + 10.0 ** rng.uniform(-12, 17, 1_500_000))
    got = ctx.f64_math(2, num, den)
    # reference: log1p of the exact difference over den in long double (num / den itself would lose the small ratios' digits)
    want = np.log1p((num.astype(np.longdouble) - den.astype(np.longdouble)) / den.astype(np.longdouble))
    err = ulp_err(got, want)
    assert err.max() <= 3.0, (err.max(), num[err.argmax()], den[err.argmax()])
