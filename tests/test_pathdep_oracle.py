"""CPU tier for the path-dependent payoffs (roadmap Phase 5): the C restatement against closed forms
(discrete geometric Asian, Black-Scholes digitals, Reiner-Rubinstein barrier with the Broadie-Glasserman-Kou shift),
against a numpy restatement on shared normals, and the identities the payoffs must satisfy on common trajectories."""
import datetime as dt
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err
from oracle import anchors as A

REF, EXP = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
T = 366 / 365

ALL_KINDS = [(abi.HH_PD_VANILLA, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_ASIAN_ARITH, 100.0, 1.0, 0.0, 0.0),
             (abi.HH_PD_ASIAN_GEOM, 95.0, -1.0, 0.0, 0.0), (abi.HH_PD_UP_OUT, 100.0, 1.0, 120.0, 1.5),
             (abi.HH_PD_UP_IN, 100.0, 1.0, 120.0, 0.0), (abi.HH_PD_DOWN_OUT, 100.0, -1.0, 85.0, 0.0),
             (abi.HH_PD_DOWN_IN, 100.0, -1.0, 85.0, 0.5), (abi.HH_PD_DIGITAL_CASH, 105.0, 1.0, 0.0, 10.0),
             (abi.HH_PD_DIGITAL_ASSET, 105.0, -1.0, 0.0, 0.0), (abi.HH_PD_ASIAN_ARITH_MINUS_GEOM, 100.0, 1.0, 0.0, 0.0)]


def numpy_stats(m, z, every, heston):
    """Independent restatement in numpy (vectorised over trajectories): spots on the grid, then the five statistics."""
    n, M = z.shape[0], z.shape[1]
    dtt = m.T / M
    if not heston:
        x = math.log(m.S0) + np.cumsum((m.r - 0.5 * m.sigma ** 2) * dtt + m.sigma * math.sqrt(dtt) * z, axis=1)
    else:
        x = np.empty((n, M))
        xc, v = np.full(n, math.log(m.S0)), np.full(n, m.V0)
        sq = math.sqrt(dtt)
        for k in range(M):
            dW1 = sq * (m.m11 * z[:, k, 0] + m.m12 * z[:, k, 1])
            dW2 = sq * (m.m21 * z[:, k, 0] + m.m22 * z[:, k, 1])
            vp = np.maximum(v, 0.0)
            K1 = xc + dtt * (m.r - 0.5 * vp)
            K2 = v + dtt * m.kappa * (m.theta - vp)
            s = np.sqrt(np.maximum(K2 if (m.flags & abi.HH_FLAG_SPLIT_STEP) else v, 0.0))
            xc, v = K1 + s * dW1, K2 + m.xi * s * dW2
            x[:, k] = xc
    S = np.exp(x)[:, every - 1::every]
    return np.stack([np.exp(x[:, -1]), S.mean(axis=1), np.exp(np.log(S).mean(axis=1)), S.max(axis=1), S.min(axis=1)])


def numpy_payoff(c, st):
    kind, K, cp, B, amt = c
    ST, Am, G, mx, mn = st
    van = np.maximum(cp * (ST - K), 0.0)
    return {abi.HH_PD_VANILLA: van, abi.HH_PD_ASIAN_ARITH: np.maximum(cp * (Am - K), 0.0),
            abi.HH_PD_ASIAN_GEOM: np.maximum(cp * (G - K), 0.0), abi.HH_PD_UP_OUT: np.where(mx >= B, amt, van),
            abi.HH_PD_UP_IN: np.where(mx >= B, van, amt), abi.HH_PD_DOWN_OUT: np.where(mn <= B, amt, van),
            abi.HH_PD_DOWN_IN: np.where(mn <= B, van, amt), abi.HH_PD_DIGITAL_CASH: np.where(cp * (ST - K) > 0, amt, 0.0),
            abi.HH_PD_DIGITAL_ASSET: np.where(cp * (ST - K) > 0, ST, 0.0),
            abi.HH_PD_ASIAN_ARITH_MINUS_GEOM: np.maximum(cp * (Am - K), 0.0) - np.maximum(cp * (G - K), 0.0)}[kind]


@pytest.mark.parametrize("model", ["gbm", "heston"])
@pytest.mark.parametrize("every", [1, 4])
def test_restatement_matches_numpy_on_shared_normals(oracle, model, every):
    n, M = 3000, 24
    heston = model == "heston"
    m = heston_model(xi=0.6) if heston else gbm_model()
    z = np.random.default_rng(4).standard_normal((n, M, 2) if heston else (n, M))
    sim = SimSpec(n_paths=n, n_steps=M, scheme=abi.HH_SCHEME_EM, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    res, st = oracle.mc_path_dependent(m, sim, ALL_KINDS, 0.97, every, want_stats=True)
    ref = numpy_stats(m, z, every, heston)
    assert rel_err(st, ref) < 1e-12
    for c, r in zip(ALL_KINDS, res):
        pay = numpy_payoff(c, ref)
        assert r.sum == pytest.approx(pay.sum(), rel=1e-12)
        assert r.price == pytest.approx(0.97 * pay.mean(), rel=1e-12)
        assert r.std_error == pytest.approx(0.97 * pay.std(ddof=1) / math.sqrt(n), rel=1e-9)


def test_antithetic_pairs_and_exact_steps_form(oracle):
    """Antithetic: the minus side is the trajectory of -Z, payoffs averaged per pair (montecarlo.jl:430-432). The
    BlackScholesExact increments give the same spots as the log-space scheme on the same normals."""
    n, M = 2000, 12
    m = gbm_model()
    z = np.random.default_rng(5).standard_normal((n, M))
    out = {}
    for scheme in (abi.HH_SCHEME_EM, abi.HH_SCHEME_EXACT_STEPS):
        sim = SimSpec(n_paths=n, n_steps=M, scheme=scheme, vr=abi.HH_VR_ANTITHETIC, rng_mode=abi.HH_RNG_NORMALS, normals=z)
        out[scheme] = oracle.mc_path_dependent(m, sim, ALL_KINDS, 1.0, 1, want_stats=True)
    res, st = out[abi.HH_SCHEME_EM]
    plus, minus = numpy_stats(m, z, 1, False), numpy_stats(m, -z, 1, False)
    assert rel_err(st[:, :n], plus) < 1e-12 and rel_err(st[:, n:], minus) < 1e-12
    for c, r in zip(ALL_KINDS, res):
        assert r.sum == pytest.approx((0.5 * (numpy_payoff(c, plus) + numpy_payoff(c, minus))).sum(), rel=1e-12)
    assert rel_err(out[abi.HH_SCHEME_EXACT_STEPS][1], st) < 1e-12


def test_closed_forms_under_black_scholes(oracle):
    mk = hh.BlackScholesInputs(REF, 0.05, 100.0, 0.2)
    cfg = hh.SimulationConfig(200_000, steps=50, base_seed=3, variance_reduction=hh.Antithetic())
    mc = hh.MonteCarlo(hh.LognormalDynamics(), hh.EulerMaruyama(), cfg)
    ps = [hh.AsianOption(100.0, EXP, hh.Call(), hh.GeometricAverage()), hh.AsianOption(105.0, EXP, hh.Put(), hh.GeometricAverage()),
          hh.DigitalOption(100.0, EXP, hh.Put(), hh.CashOrNothing(2.0)), hh.DigitalOption(100.0, EXP, hh.Call(), hh.AssetOrNothing()),
          hh.BarrierOption(100.0, 125.0, EXP, hh.Call(), hh.Up(), hh.KnockIn()),
          hh.BarrierOption(100.0, 125.0, EXP, hh.Call(), hh.Up(), hh.KnockOut()),
          hh.VanillaOption(100.0, EXP, hh.European(), hh.Call(), hh.Spot()), hh.AsianOption(100.0, EXP, hh.Call())]
    sols = hh.solve(hh.BasketPricingProblem(ps, mk), mc, engine=oracle)
    Hs = A.discrete_barrier_shift(125.0, 0.2, T, 50)
    ui = A.up_and_in_call_price(100.0, 100.0, Hs, 0.05, 0.2, T)
    bs = A.bs_price(100.0, 100.0, 0.05, 0.2, T)
    exact = [A.geometric_asian_price(100.0, 100.0, 0.05, 0.2, T, 50), A.geometric_asian_price(100.0, 105.0, 0.05, 0.2, T, 50, cp=-1.0),
             A.digital_price(100.0, 100.0, 0.05, 0.2, T, -1.0, 2.0), A.digital_price(100.0, 100.0, 0.05, 0.2, T, 1.0), ui, bs - ui, bs]
    slack = [0, 0, 0, 0, 1e-2 * ui, 1e-2 * ui, 0]   # the continuity correction is itself an approximation (~0.5 % of the knock-in)
    for s, e, sl in zip(sols, exact, slack):
        assert abs(s.price - e) < 3.5 * s.std_error + sl, (s.price, e, s.std_error)
    # in + out = vanilla on common trajectories, to rounding; arithmetic >= geometric average (AM-GM)
    assert sols[4].price + sols[5].price == pytest.approx(sols[6].price, rel=1e-12)
    assert sols[7].price > sols[0].price
    one = hh.solve(hh.PricingProblem(ps[0], mk), mc, engine=oracle)
    assert one.price == pytest.approx(sols[0].price, rel=1e-12)


def test_monitoring_edge_cases(oracle):
    m = gbm_model()
    z = np.random.default_rng(6).standard_normal((500, 8))
    sim = SimSpec(n_paths=500, n_steps=8, scheme=abi.HH_SCHEME_EM, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    # one monitoring date at expiry: both averages are S_T and the barrier is a terminal condition
    res, st = oracle.mc_path_dependent(m, sim, ALL_KINDS, 1.0, 8, want_stats=True)
    assert rel_err(st[1], st[0]) < 1e-15 and rel_err(st[2], st[0]) < 1e-14 and np.all(st[3] == st[0]) and np.all(st[4] == st[0])
    assert res[1].sum == pytest.approx(res[0].sum, rel=1e-14)
    with pytest.raises(ValueError):
        oracle.mc_path_dependent(m, sim, ALL_KINDS, 1.0, 3)   # 8 is not a multiple of 3
    with pytest.raises(NotImplementedError):                  # the terminal-law sampler saves no dates
        oracle.mc_path_dependent(m, SimSpec(n_paths=10, n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL), ALL_KINDS, 1.0, 1)
    with pytest.raises(ValueError):
        hh.solve(hh.BasketPricingProblem([hh.AsianOption(100.0, EXP, hh.Call()), hh.AsianOption(100.0, EXP, hh.Call(), monitoring=hh.Monitoring(2))],
                                         hh.BlackScholesInputs(REF, 0.05, 100.0, 0.2)),
                 hh.MonteCarlo(hh.LognormalDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(100, steps=4)), engine=oracle)


def test_geometric_control_variate_for_the_arithmetic_asian(oracle):
    """Roadmap "control variates" (SURVEY N3): arithmetic - geometric on common trajectories plus the closed-form geometric
    price. Same expectation as the plain estimator (3.5 sigma), standard error more than 10x smaller; the closed form on
    the product side equals the oracle's anchor."""
    from hedgehog_jl_b200.pathdep import geometric_asian_closed_form
    mk = hh.BlackScholesInputs(REF, 0.05, 100.0, 0.2)
    mc = hh.MonteCarlo(hh.LognormalDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(100_000, steps=50, base_seed=3))
    ps = [hh.AsianOption(100.0, EXP, hh.Call()), hh.AsianOption(100.0, EXP, hh.Call(), control_variate=hh.GeometricControlVariate()),
          hh.AsianOption(95.0, EXP, hh.Put()), hh.AsianOption(95.0, EXP, hh.Put(), control_variate=hh.GeometricControlVariate())]
    plain_c, cv_c, plain_p, cv_p = hh.solve(hh.BasketPricingProblem(ps, mk), mc, engine=oracle)
    for plain, cv in ((plain_c, cv_c), (plain_p, cv_p)):
        assert abs(plain.price - cv.price) < 3.5 * plain.std_error
        assert cv.std_error < 0.1 * plain.std_error
    assert geometric_asian_closed_form(100.0, 95.0, 0.05, 0.2, T, 50, -1.0) == pytest.approx(
        A.geometric_asian_price(100.0, 95.0, 0.05, 0.2, T, 50, cp=-1.0), rel=1e-13)
    with pytest.raises(TypeError):   # no closed form under Heston
        hh.solve(hh.PricingProblem(ps[1], hh.HestonInputs(REF, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)),
                 hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(1000, steps=10)), engine=oracle)
    with pytest.raises(ValueError):
        hh.AsianOption(100.0, EXP, hh.Call(), hh.GeometricAverage(), control_variate=hh.GeometricControlVariate())


def test_black_scholes_control_variate_for_heston(oracle):
    """Roadmap "Control variates using Black-Scholes" (SURVEY N3): the control's sample mean matches its closed form, the
    controlled estimator has the plain estimator's expectation and a much smaller standard error, and it agrees with
    Carr-Madan up to the scheme's O(dt) bias."""
    from hedgehog_jl_b200.pathdep import bs_control_sigma, black_scholes_closed_form
    mk = hh.HestonInputs(REF, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    prob = hh.PricingProblem(hh.VanillaOption(100.0, EXP, hh.European(), hh.Call(), hh.Spot()), mk)
    cfg = hh.SimulationConfig(100_000, steps=100, base_seed=9)
    plain = hh.solve(prob, hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), cfg, ensemble=False), engine=oracle)
    cv = hh.solve(prob, hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), cfg, ensemble=False,
                                      control_variate=hh.BlackScholesControlVariate()), engine=oracle)
    assert cv.std_error < 0.5 * plain.std_error and 0.5 < cv.stats["beta"] < 1.0
    assert abs(cv.price - plain.price) < 3.5 * plain.std_error
    cm = A.heston_price(100.0, 100.0, 0.03, T, 0.04, 2.0, 0.04, 0.3, -0.7)
    assert abs(cv.price - cm) < 3.5 * cv.std_error + 0.02   # + the Euler bias at 100 steps (profiles/r1_i_euler_bias_c2.json)
    # the control alone: a log-GBM trajectory with sigma_cv, whose mean is the Black-Scholes price
    sig = bs_control_sigma(0.09, 1.5, 0.04, T)
    assert sig == pytest.approx(math.sqrt(0.04 + 0.05 * (1 - math.exp(-1.5 * T)) / (1.5 * T)), rel=1e-14)
    m = heston_model(r=0.03, T=T, V0=0.09, kappa=1.5, theta=0.04, xi=0.4, rho=-0.5)
    sim = SimSpec(n_paths=200_000, n_steps=50, scheme=abi.HH_SCHEME_EM, vr=abi.HH_VR_ANTITHETIC, base_seed=2)
    (ctrl,), _ = oracle.mc_path_dependent(m, sim, [(abi.HH_PD_BS_CONTROL, 105.0, -1.0, 0.0, 0.0)], math.exp(-0.03 * T), 1)
    assert abs(ctrl.price - black_scholes_closed_form(100.0, 105.0, 0.03, sig, T, -1.0)) < 3.5 * ctrl.std_error
    assert black_scholes_closed_form(100.0, 105.0, 0.03, sig, T, -1.0) == pytest.approx(A.bs_price(100.0, 105.0, 0.03, sig, T, cp=-1.0), rel=1e-12)
    with pytest.raises(ValueError):   # the control needs HestonDynamics + EulerMaruyama
        oracle.mc_path_dependent(gbm_model(), SimSpec(n_paths=10, n_steps=5, scheme=abi.HH_SCHEME_EM), [(abi.HH_PD_BS_CONTROL, 100.0, 1.0, 0.0, 0.0)], 1.0, 1)
