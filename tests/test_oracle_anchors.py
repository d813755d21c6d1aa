"""Pins the CPU oracle (test infrastructure) against every deterministic known answer the reference's tests hold
for this path: Black-Scholes, CRR, Carr-Madan, payoff, discount factors, ACT/365 ticks, and the
Random123 Philox4x32-10 known-answer vectors the native RNG stream is built on."""
import datetime as dt
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from oracle import anchors as A
from oracle import oracle as O


def test_black_scholes_known_answers():
    """test/unit/black_scholes.jl:75-129 (atol 1e-4)."""
    S, r, sig, T = 100.0, 0.05, 0.20, 1.0
    F = S * math.exp(r * T)
    assert A.bs_price(S, F, r, sig, T, +1) == pytest.approx(7.9655, abs=1e-4)     # :93
    assert A.bs_price(S, 90.0, r, sig, T, +1) == pytest.approx(16.6994, abs=1e-4)  # :103
    assert A.bs_price(S, 90.0, r, sig, T, -1) == pytest.approx(2.3101, abs=1e-4)   # :113
    assert A.bs_price(S, 110.0, r, sig, 91 / 365, -1) == pytest.approx(9.8237, abs=1e-4)  # :126


def test_crr_known_answers():
    """test/unit/binomial_tree.jl:18,26 (atol 1e-8)."""
    assert A.crr_price(1.0, 1.0, 0.2, 0.4, 1.0, 80, cp=+1, american=True, underlying="spot") == \
        pytest.approx(0.25225758542934945, abs=1e-8)
    assert A.crr_price(1.0, 1.0, 0.2, 0.4, 1.0, 80, cp=-1, american=True, underlying="forward") == \
        pytest.approx(0.07409148128021317, abs=1e-8)


def test_crr_and_carr_madan_agree_with_black_scholes():
    """test/agreement/price_agreement.jl: CRR vs BS atol 1e-3 (European), Carr-Madan vs BS atol 1e-6."""
    S, K, r, sig, T = 100.0, 100.0, 0.05, 0.2, 1.0
    bs = A.bs_price(S, K, r, sig, T)
    assert A.crr_price(S, K, r, sig, T, 10_000) == pytest.approx(bs, abs=1e-3)
    cm = A.carr_madan_price(lambda u: A.gbm_cf(u, S, r, sig, T), S, K, r, T, alpha=1.0, bound=32.0)
    assert cm == pytest.approx(bs, abs=1e-6)


def test_heston_carr_madan_literature_value():
    """Broadie-Kaya (2006) case 1 — the parameter set of test/unit/calibration.jl:39-40: true price 6.8061."""
    p = A.heston_price(100.0, 100.0, 0.0319, 1.0, 0.010201, 6.21, 0.019, 0.61, -0.7, bound=400.0)
    assert p == pytest.approx(6.8061, abs=1e-4)
    # the reference's tests integrate over (-32, 32) (montecarlo_heston.jl:47,199): truncation costs 5e-4 here
    p32 = A.heston_price(100.0, 100.0, 0.0319, 1.0, 0.010201, 6.21, 0.019, 0.61, -0.7, bound=32.0)
    assert p32 == pytest.approx(p, abs=1e-3)


def test_payoff_df_and_act365():
    """test/unit/payoff.jl:8-21, rate_curve.jl:44-62, date_functions.jl."""
    ref = dt.date(2020, 1, 1)
    call = hh.VanillaOption(100.0, ref + dt.timedelta(days=365), hh.European(), hh.Call(), hh.Spot())
    put = hh.VanillaOption(100.0, ref + dt.timedelta(days=365), hh.European(), hh.Put(), hh.Spot())
    assert call(120.0) == 20.0 and put(120.0) == 0.0 and put(80.0) == 20.0
    assert call.call_put() == 1.0 and put.call_put() == -1.0
    assert hh.to_ticks(dt.date(1970, 1, 1)) == 719163 * 86_400_000   # Dates.date2epochdays(1970-01-01)
    assert hh.yearfrac(ref, ref + dt.timedelta(days=365)) == 1.0
    assert hh.yearfrac(ref, dt.date(2021, 1, 1)) == 366 / 365          # 2020 is a leap year
    curve = hh.FlatRateCurve.of(0.025, dt.date(2025, 1, 1))
    assert hh.zero_rate(curve, dt.date(2026, 1, 1)) == 0.025
    assert hh.df(curve, dt.date(2026, 1, 1)) == pytest.approx(math.exp(-0.025 * 365 / 365), abs=1e-12)
    assert hh.add_yearfrac(ref, 0.5) == hh.to_ticks(ref) + 0.5 * 365 * 86_400_000


@pytest.mark.parametrize("ctr,key,out", [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox4x32_10_known_answers(ctr, key, out):
    """Random123 kat_vectors: philox4x32 10."""
    assert tuple(O.philox(ctr, key)) == out


def test_philox_matches_numpy_philox_free_implementation():
    """Independent cross-check of the round function on random counters."""
    rng = np.random.default_rng(1)

    def ref(c, k):
        c, k = list(map(int, c)), list(map(int, k))
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
            k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
        return c

    for _ in range(50):
        c = rng.integers(0, 2**32, 4)
        k = rng.integers(0, 2**32, 2)
        assert O.philox(c, k) == ref(c, k)


def test_normal_pair_is_box_muller_of_the_documented_bits():
    for key, idx, blk in [(42, 0, 0), (2**63 + 5, 123456789012, 251), (7, 2**40, 3)]:
        w = O.philox((idx & 0xffffffff, idx >> 32, blk, 0), (key & 0xffffffff, key >> 32))
        n1 = ((w[1] & 0xFFFFF) << 32) | (w[0] | 1)
        n2 = ((w[3] & 0xFFFFF) << 32) | w[2]
        u1, u2 = 1.0 - n1 * 2.0**-52, n2 * 2.0**-52
        r = math.sqrt(-2.0 * math.log(u1))
        z1, z2 = O.normal_pair(key, idx, blk)
        assert z1 == pytest.approx(r * math.cos(2 * math.pi * u2), rel=1e-13, abs=1e-15)
        assert z2 == pytest.approx(r * math.sin(2 * math.pi * u2), rel=1e-13, abs=1e-15)


def test_native_normals_are_standard_normal():
    from helpers import heston_model
    from hedgehog_jl_b200 import _abi as abi
    from hedgehog_jl_b200.engine import SimSpec
    z = O.OracleEngine().fill_normals(heston_model(), SimSpec(n_paths=20000, n_steps=16, scheme=abi.HH_SCHEME_EM, base_seed=1))
    z = z.ravel()
    n = z.size
    assert abs(z.mean()) < 4 / math.sqrt(n)
    assert abs(z.var() - 1) < 4 * math.sqrt(2 / n)
    assert abs((z**3).mean()) < 4 * math.sqrt(15 / n)
    assert abs((z**4).mean() - 3) < 4 * math.sqrt(96 / n)
    zz = z.reshape(-1, 2)
    assert abs(np.corrcoef(zz[:, 0], zz[:, 1])[0, 1]) < 4 / math.sqrt(n / 2)
