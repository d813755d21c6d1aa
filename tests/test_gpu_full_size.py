"""The five BASELINE.json configurations at their FULL sizes, through properties that do not need a per-path oracle run
(the CPU oracle needs minutes at these sizes): shard additivity of the payoff sums (a checksum of checksums), put-call parity
and convexity on common trajectories, monotonicity, determinism, agreement with the closed-form / Carr-Madan / CRR anchors of
tests/golden/config_anchors.json (written by tools/gen_anchors.py with the oracle's restatement of the reference's own
analytic pricers). Whole file: a few seconds of GPU time."""
import datetime as dt
import json
import math
import os

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ANCHORS = json.load(open(os.path.join(ROOT, "tests", "golden", "config_anchors.json")))
EULER_BIAS_252 = 0.005651   # price(252 steps) - Carr-Madan at the C2 parameters (tools/euler_bias.py, 2e8 paths, +-0.0005)


def test_c1_exact_gbm_1e6(cuda):
    g, D, n = gbm_model(), math.exp(-0.05), 1_000_000
    pay = [(100.0, 1.0), (100.0, -1.0)]
    res, _ = cuda.mc_european(g, SimSpec(n_paths=n, n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL, base_seed=42), pay, D)
    call, put = res
    assert abs(call.price - ANCHORS["c1_black_scholes_call"]) < 3.5 * call.std_error
    # put-call parity on common trajectories: call - put = D (mean S_T - K), and D mean S_T = S0 within the sampling error of S_T
    fwd = call.price - put.price + D * 100.0
    assert abs(fwd - 100.0) < 4 * 100.0 * 0.2 / math.sqrt(n) * 1.1
    again, _ = cuda.mc_european(g, SimSpec(n_paths=n, n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL, base_seed=42), pay, D)
    assert (again[0].sum, again[0].sumsq) == (call.sum, call.sumsq)


def test_c2_heston_em_1e8_x_252(cuda):
    m, D, n = heston_model(), math.exp(-0.03), 100_000_000
    pay = [(100.0, 1.0), (100.0, -1.0), (90.0, 1.0), (110.0, 1.0)]
    full, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=252, base_seed=42), pay, D)
    c, p, c90, c110 = full
    assert c.n == n and c.n_nonfinite == 0
    assert abs(c.price - ANCHORS["c2_carr_madan_call"] - EULER_BIAS_252) < 3.5 * c.std_error     # se ~ 1.2e-3
    # put-call parity on common trajectories is an identity of the sums (to the rounding of 1e8 additions)
    lhs = c.sum - p.sum
    # mean S_T from two strikes: (call - put)(K) = mean S_T - K for every K
    res2, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=252, base_seed=42), [(110.0, 1.0), (110.0, -1.0)], D)
    assert abs((lhs / n + 100.0) - ((res2[0].sum - res2[1].sum) / n + 110.0)) < 1e-9
    assert abs(res2[0].sum - c110.sum) <= 1e-13 * c110.sum     # another payoff slot of another launch: the order of additions only
    assert c90.price > c.price > c110.price and c90.price - c.price < 10.0 * D          # monotone, slope in [-D, 0]
    assert c90.price - 2 * c.price + c110.price > 0                                     # convex in the strike
    # shard additivity: four shards of 2.5e7 trajectories (global index offsets) reproduce the sums of the whole job
    s = [cuda.mc_european(m, SimSpec(n_paths=n // 4, path_offset=k * (n // 4), n_steps=252, base_seed=42), pay[:1], D)[0][0] for k in range(4)]
    assert sum(x.n for x in s) == n
    assert abs(sum(x.sum for x in s) - c.sum) <= 1e-12 * c.sum
    assert abs(sum(x.sumsq for x in s) - c.sumsq) <= 1e-12 * c.sumsq
    # the Float32 fast mode and the opt-in 64-bit stream agree with it statistically
    f32, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=252, base_seed=42, precision=abi.HH_PREC_F32), pay[:1], D)
    p64, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=252, base_seed=42, rng_mode=abi.HH_RNG_PHILOX_64), pay[:1], D)
    for other in (f32[0], p64[0]):
        assert abs(other.price - c.price) < 4 * math.hypot(other.std_error, c.std_error)


def test_c3_lsm_1e7_x_50(cuda):
    g, n, steps = gbm_model(), 10_000_000, 50
    Dstep = math.exp(-0.05 / steps)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=12345)
    out, *_ = cuda.lsm_american(g, sim, (100.0, -1.0), 3, Dstep)
    crr, berm = ANCHORS["c3_crr_american_put_1000"], ANCHORS["c3_crr_bermudan_put_50_dates"]
    assert abs(out.price - crr) < 0.02 * crr                                   # the reference's own bar (american_options.jl:49)
    assert abs(out.price - berm) < 4 * out.std_error + 0.002 * berm            # the 50-date Bermudan, cubic-basis bias included
    # American >= European on the same trajectories' law
    eur, _ = cuda.mc_european(g, SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=12345), [(100.0, -1.0)],
                              math.exp(-0.05))
    assert out.price > eur[0].price + 5 * (out.std_error + eur[0].std_error)
    again, *_ = cuda.lsm_american(g, sim, (100.0, -1.0), 3, Dstep)
    assert (again.sum, again.sumsq, again.n_dates_skipped) == (out.sum, out.sumsq, out.n_dates_skipped)
    deg5, *_ = cuda.lsm_american(g, sim, (100.0, -1.0), 5, Dstep)
    assert abs(deg5.price - out.price) < 0.003 * out.price                     # a richer basis moves the price by basis points


def test_c4_broadie_kaya_1e7_x_12(cuda):
    m, D, n = heston_model(), math.exp(-0.03), 10_000_000
    sim = SimSpec(n_paths=n, n_steps=12, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=42)
    res, _ = cuda.mc_european(m, sim, [(100.0, 1.0), (100.0, -1.0)], D)
    c, p = res
    st = cuda.bk_last_stats()
    assert st["transitions"] == 12 * n and st["n_fallback"] == 0 and c.n_nonfinite == 0
    assert abs(c.price - ANCHORS["c2_carr_madan_call"]) < 3.5 * c.std_error   # exact simulation: no time-stepping bias
    assert abs((c.sum - p.sum) / n * D + 100.0 * D - 100.0) < 4 * 100.0 * 0.2 / math.sqrt(n) * 1.2   # D E[S_T] = S0
    halves = [cuda.mc_european(m, SimSpec(n_paths=n // 2, path_offset=k * (n // 2), n_steps=12, scheme=abi.HH_SCHEME_HESTON_BK,
                                          base_seed=42), [(100.0, 1.0)], D)[0][0] for k in range(2)]
    assert abs(halves[0].sum + halves[1].sum - c.sum) <= 1e-12 * c.sum


def test_c5_greeks_1e7_x_252_on_64_strikes(cuda):
    payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
    market = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    strikes = np.linspace(60.0, 140.0, 64)
    lenses = [hh.SpotLens(), hh.optic("market_inputs.V0"), hh.ZeroRateSpineLens(1), hh.optic("market_inputs.kappa"),
              hh.optic("market_inputs.theta"), hh.optic("market_inputs.sigma"), hh.optic("market_inputs.rho")]
    m = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(10_000_000, steps=252, base_seed=42), ensemble=False)
    prices, g, se, sec = hh.strike_grid_greeks(hh.PricingProblem(payoff, market), strikes, lenses, m, engine=cuda, gamma_bump=0.5)
    a = ANCHORS["c5"]
    sel = slice(8, 56)                                                     # strikes 70 .. 130
    assert np.all(np.diff(prices) < 0)                                      # calls fall with the strike
    assert np.all(np.diff(prices, 2) > -1e-9)                               # and are convex in it (common trajectories: exact)
    assert np.all((g[:, 0] > 0) & (g[:, 0] < 1)) and np.all(np.diff(g[:, 0]) < 1e-12)   # delta in (0, 1), falling with the strike
    assert np.all(sec["fd"][sel] > 0)                                       # gamma
    # against finite differences of Carr-Madan: the reference's own Monte Carlo bars (greeks_agreement.jl:207-236), tightened
    assert np.max(np.abs(prices[sel] / np.array(a["price"])[sel] - 1)) < 0.02
    assert np.max(np.abs(g[sel, 0] / np.array(a["d_S0"])[sel] - 1)) < 0.02
    assert np.max(np.abs(g[sel, 2] / np.array(a["d_r"])[sel] - 1)) < 0.02
    assert np.max(np.abs(g[sel, 1] / np.array(a["d_V0"])[sel] - 1)) < 0.05
    mid = slice(24, 40)
    assert np.max(np.abs(sec["pathwise"][mid] / np.array(a["d2_S0"])[mid] - 1)) < 0.1
