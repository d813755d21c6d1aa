"""The reference's own statistical agreement tests (★ rows), run through the host layer's solve() with
the CPU oracle injected as the engine. This pins the oracle's Monte Carlo arithmetic (and the host-side scalar
extraction) with the reference's parameters, path counts and tolerances; the GPU tests then compare CUDA to it."""
import datetime as dt
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from oracle import anchors as A

REF = dt.date(2020, 1, 1)
EXP = dt.date(2021, 1, 1)  # reference_date + Year(1) = 366 days


def _seeds(seed, n):
    return np.random.Generator(np.random.Philox(seed)).integers(0, 2**64, size=n, dtype=np.uint64)


@pytest.mark.parametrize("strategy", ["exact", "em"])
@pytest.mark.parametrize("anti", [False, True])
def test_gbm_scenarios_vs_black_scholes(oracle, strategy, anti):
    """test/agreement/montecarlo_black_scholes.jl:8-169 — 10 000 paths, steps=1, 5 trials, mean vs BS rtol 0.02."""
    prob = hh.PricingProblem(hh.VanillaOption(100.0, EXP, hh.European(), hh.Call(), hh.Spot()),
                             hh.BlackScholesInputs(REF, 0.05, 100.0, 0.20))
    bs = A.bs_price(100.0, 100.0, 0.05, 0.20, 366 / 365)
    prices, variances = [], []
    for trial in range(1, 6):
        cfg = hh.SimulationConfig(10_000, steps=1, seeds=_seeds(42 + trial, 10_000),
                                  variance_reduction=hh.Antithetic() if anti else hh.NoVarianceReduction())
        mc = hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact() if strategy == "exact" else hh.EulerMaruyama(), cfg)
        sol = hh.solve(prob, mc, engine=oracle)
        prices.append(sol.price)
        ens = sol.ensemble
        pay = prob.payoff(ens) if not anti else 0.5 * (prob.payoff(ens[0]) + prob.payoff(ens[1]))
        variances.append(pay.var())
        assert len(pay) == 10_000
    assert np.mean(prices) == pytest.approx(bs, rel=0.02)  # :130
    if anti:  # variance(antithetic) < variance(plain)  :141,151
        cfg = hh.SimulationConfig(10_000, steps=1, seeds=_seeds(43, 10_000))
        plain = hh.solve(prob, hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact() if strategy == "exact"
                                             else hh.EulerMaruyama(), cfg), engine=oracle)
        assert variances[0] < prob.payoff(plain.ensemble).var()


def test_heston_em_vs_carr_madan(oracle):
    """test/agreement/montecarlo_heston.jl:8-144 — S=K=100, r=.03, V0=.04, κ=2, θ=.04, σ=.3, ρ=-.7; steps=1 (Q9); rtol 0.05."""
    prob = hh.PricingProblem(hh.VanillaOption(100.0, EXP, hh.European(), hh.Call(), hh.Spot()),
                             hh.HestonInputs(REF, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7))
    cm = A.heston_price(100.0, 100.0, 0.03, 366 / 365, 0.04, 2.0, 0.04, 0.3, -0.7)
    plain = hh.solve(prob, hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(),
                                         hh.SimulationConfig(5000, steps=1, seeds=_seeds(1, 5000))), engine=oracle)
    anti = hh.solve(prob, hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(),
                                        hh.SimulationConfig(2500, steps=1, seeds=_seeds(1, 2500),
                                                            variance_reduction=hh.Antithetic())), engine=oracle)
    assert plain.price == pytest.approx(cm, rel=0.05)  # :116
    assert anti.price == pytest.approx(cm, rel=0.05)
    v_plain = prob.payoff(plain.ensemble).var()
    v_anti = (0.5 * (prob.payoff(anti.ensemble[0]) + prob.payoff(anti.ensemble[1]))).var()
    assert v_plain / v_anti > 1  # :126


def test_heston_em_q8_parameters_vs_carr_madan(oracle):
    """montecarlo_heston.jl:151-206 — positional mix-up (Q8): V0=1.5, κ=0.04, θ=0.3, σ=-0.6, ρ=0.04; EM antithetic
    50 000 x 200 steps vs Carr-Madan rtol 2e-2 (:205)."""
    expiry = REF + dt.timedelta(days=364)
    prob = hh.PricingProblem(hh.VanillaOption(100.0, expiry, hh.European(), hh.Call(), hh.Spot()),
                             hh.HestonInputs(REF, 0.05, 100.0, 1.5, 0.04, 0.3, -0.6, 0.04))
    cm = A.heston_price(100.0, 100.0, 0.05, 364 / 365, 1.5, 0.04, 0.3, -0.6, 0.04)
    sol = hh.solve(prob, hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(),
                                       hh.SimulationConfig(50_000, steps=200, base_seed=7, variance_reduction=hh.Antithetic()),
                                       ensemble=False), engine=oracle)
    assert sol.price == pytest.approx(cm, rel=2e-2)


def test_lsm_american_put_vs_crr(oracle):
    """test/agreement/american_options.jl:9-52 — 50 000 x 100, antithetic, degree 5 vs CRR(1000), rtol 0.02."""
    prob = hh.PricingProblem(hh.VanillaOption(100.0, EXP, hh.American(), hh.Put(), hh.Spot()),
                             hh.BlackScholesInputs(REF, 0.05, 100.0, 0.2))
    crr = A.crr_price(100.0, 100.0, 0.05, 0.2, 366 / 365, 1000, cp=-1, american=True)
    cfg = hh.SimulationConfig(50_000, steps=100, seeds=_seeds(12345, 50_000), variance_reduction=hh.Antithetic())
    sol = hh.solve(prob, hh.LSM(hh.LognormalDynamics(), hh.BlackScholesExact(), cfg, 5), engine=oracle, spot_paths=True)
    assert sol.price == pytest.approx(crr, rel=0.02)
    assert sol.spot_paths.shape == (101, 100_000)  # (steps+1) x 2N columns, [normal | antithetic]  lsm.jl:70-85
    assert len(sol.stopping_info) == 100_000
    taus = np.array([t for t, _ in sol.stopping_info])
    assert taus.min() >= 1 and taus.max() == 100  # no exercise at t = 0 (lsm.jl:114)
    # American >= European (american_options.jl:148-202)
    euro = A.bs_price(100.0, 100.0, 0.05, 0.2, 366 / 365, cp=-1)
    assert sol.price > euro


@pytest.mark.parametrize("strike", [80.0, 90.0, 100.0, 110.0, 120.0])
def test_lsm_multiple_strikes_vs_crr(oracle, strike):
    """american_options.jl:98-146 — 6M puts, sigma=.25, 20 000 x 50, degree 4; rtol .05 (OTM) / .03."""
    expiry = dt.date(2020, 7, 1)
    T = hh.yearfrac(REF, expiry)
    prob = hh.PricingProblem(hh.VanillaOption(strike, expiry, hh.American(), hh.Put(), hh.Spot()),
                             hh.BlackScholesInputs(REF, 0.05, 100.0, 0.25))
    crr = A.crr_price(100.0, strike, 0.05, 0.25, T, 500, cp=-1, american=True)
    cfg = hh.SimulationConfig(20_000, steps=50, seeds=_seeds(int(strike) * 1000, 20_000), variance_reduction=hh.Antithetic())
    sol = hh.solve(prob, hh.LSM(hh.LognormalDynamics(), hh.BlackScholesExact(), cfg, 4), engine=oracle, stopping_info=False)
    assert sol.price == pytest.approx(crr, rel=0.05 if strike < 100.0 else 0.03)


def test_lsm_american_call_high_rate_vs_crr(oracle):
    """american_options.jl:54-96 — call, r=.15, S=120, sigma=.3 vs CRR(800), rtol .03."""
    prob = hh.PricingProblem(hh.VanillaOption(100.0, EXP, hh.American(), hh.Call(), hh.Spot()),
                             hh.BlackScholesInputs(REF, 0.15, 120.0, 0.3))
    crr = A.crr_price(120.0, 100.0, 0.15, 0.3, 366 / 365, 800, cp=+1, american=True)
    cfg = hh.SimulationConfig(30_000, steps=50, seeds=_seeds(54321, 30_000), variance_reduction=hh.Antithetic())
    sol = hh.solve(prob, hh.LSM(hh.LognormalDynamics(), hh.BlackScholesExact(), cfg, 4), engine=oracle, stopping_info=False)
    assert sol.price == pytest.approx(crr, rel=0.03)


def test_mc_greeks_vs_analytic(oracle):
    """test/agreement/greeks_agreement.jl:170-241 — GBM exact, 100 000 paths, S=K=1, σ=1, r=.03."""
    prob = hh.PricingProblem(hh.VanillaOption(1.0, EXP, hh.European(), hh.Call(), hh.Spot()),
                             hh.BlackScholesInputs(REF, 0.03, 1.0, 1.0))
    mc = hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(100_000, seeds=_seeds(42, 100_000)))
    an = A.bs_greeks(1.0, 1.0, 0.03, 1.0, 366 / 365)
    assert hh.solve(prob, mc, engine=oracle).price == pytest.approx(A.bs_price(1.0, 1.0, 0.03, 1.0, 366 / 365), rel=3e-2)
    g = lambda lens, m=hh.ForwardAD(): hh.solve(hh.GreekProblem(prob, lens), m, mc, engine=oracle).greek
    assert g(hh.SpotLens()) == pytest.approx(an["delta"], rel=3e-2)
    assert g(hh.VolLens(1, 1)) == pytest.approx(an["vega"], rel=1e-1)
    assert g(hh.ZeroRateSpineLens(1)) == pytest.approx(an["rho"], rel=3e-2)
    gamma = hh.solve(hh.SecondOrderGreekProblem(prob, hh.SpotLens(), hh.SpotLens()), hh.FiniteDifference(1e-1), mc, engine=oracle).greek
    assert gamma == pytest.approx(an["gamma"], rel=2e-1)
    # FD (relative bump, common random numbers) agrees with the pathwise tangent
    fd = g(hh.SpotLens(), hh.FiniteDifference(1e-4))
    assert fd == pytest.approx(g(hh.SpotLens()), rel=2e-3)


def test_q1_sqrt_alpha_mean_switch(oracle):
    """Q1 (montecarlo.jl:302): reference-compat puts sqrt(alpha) in the mean; the two modes agree iff alpha = 1."""
    mk = lambda exp, q1: hh.solve(
        hh.PricingProblem(hh.VanillaOption(100.0, exp, hh.European(), hh.Call(), hh.Spot()),
                          hh.BlackScholesInputs(REF, 0.05, 100.0, 0.2)),
        hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(2000, base_seed=3), q1_compat=q1),
        engine=oracle).price
    one_year = REF + dt.timedelta(days=365)
    assert mk(one_year, True) == mk(one_year, False)
    two_years = REF + dt.timedelta(days=730)
    assert mk(two_years, True) != mk(two_years, False)


def test_f32_restatement_agrees_with_f64_statistically(oracle):
    """The binary32 restatement of Heston EM (the checker of the GPU's f32 fast mode) prices within 3 sigma of the
    binary64 one and of Carr-Madan."""
    from hedgehog_jl_b200 import _abi as abi
    from hedgehog_jl_b200.engine import SimSpec
    from helpers import heston_model
    from oracle import anchors
    m = heston_model()
    D = math.exp(-m.r * m.T)
    n = 200_000
    r32, _ = oracle.mc_european(m, SimSpec(n_paths=n, n_steps=100, precision=abi.HH_PREC_F32, base_seed=8), [(100.0, 1.0)], D)
    r64, _ = oracle.mc_european(m, SimSpec(n_paths=n, n_steps=100, base_seed=8), [(100.0, 1.0)], D)
    cm = anchors.heston_price(100.0, 100.0, m.r, m.T, m.V0, m.kappa, m.theta, m.xi, m.rho)
    assert abs(r32[0].price - r64[0].price) < 3 * math.hypot(r32[0].std_error, r64[0].std_error)
    assert abs(r32[0].price - cm) < 3 * r32[0].std_error + 0.02


# ---- LSM on the log-space schemes (extraction S = exp(x)) ---------------------------------------------

def test_lsm_logspace_extraction_is_consistent_across_schemes(oracle):
    """The same normals through BlackScholesExact (S-space), log-GBM Euler-Maruyama and a degenerate log-Heston
    (xi = 0, V0 = theta = sigma^2) are one law: equal spot grids (to rounding) and equal LSM prices."""
    from hedgehog_jl_b200 import _abi as abi
    from hedgehog_jl_b200.engine import SimSpec
    from helpers import gbm_model, heston_model, rel_err
    n, steps = 5000, 20
    z2 = np.random.default_rng(3).standard_normal((n, steps, 2))
    z1 = np.ascontiguousarray(z2[:, :, 0])
    mg, mh = gbm_model(r=0.03, sigma=0.2), heston_model(r=0.03, V0=0.04, theta=0.04, xi=0.0, rho=0.0)
    D = math.exp(-0.03 / steps)
    runs = []
    for m, scheme, z in ((mg, abi.HH_SCHEME_EXACT_STEPS, z1), (mg, abi.HH_SCHEME_EM, z1), (mh, abi.HH_SCHEME_EM, z2)):
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=abi.HH_VR_ANTITHETIC, rng_mode=abi.HH_RNG_NORMALS, normals=z)
        runs.append(oracle.lsm_american(m, sim, (100.0, -1.0), 3, D, want_stopping=True, want_paths=True))
    base = runs[0]
    for r in runs[1:]:
        assert rel_err(r[3], base[3]) < 1e-12
        assert np.mean(r[1] != base[1]) < 1e-3
        assert r[0].price == pytest.approx(base[0].price, rel=1e-4)
    assert np.all(base[3][:, 0] == 100.0)


def test_lsm_heston_american_put_premium(oracle):
    """LSM(HestonDynamics, EulerMaruyama): American >= European (Carr-Madan by put-call parity), with a premium of
    the Black-Scholes order at the same volatility level."""
    prob = hh.PricingProblem(hh.VanillaOption(100.0, EXP, hh.American(), hh.Put(), hh.Spot()),
                             hh.HestonInputs(REF, 0.05, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7))
    cfg = hh.SimulationConfig(40_000, steps=50, base_seed=11, variance_reduction=hh.Antithetic())
    sol = hh.solve(prob, hh.LSM(hh.HestonDynamics(), hh.EulerMaruyama(), cfg, 3), engine=oracle, stopping_info=False)
    T = 366 / 365
    euro_put = A.heston_price(100.0, 100.0, 0.05, T, 0.04, 2.0, 0.04, 0.3, -0.7) - 100.0 + 100.0 * math.exp(-0.05 * T)
    bs_premium = A.crr_price(100.0, 100.0, 0.05, 0.2, T, 1000, cp=-1, american=True) - A.bs_price(100.0, 100.0, 0.05, 0.2, T, cp=-1)
    assert sol.price > euro_put
    assert 0.4 * bs_premium < sol.price - euro_put < 2.0 * bs_premium


def test_lsm_rejects_schemes_without_saved_dates(oracle):
    from hedgehog_jl_b200 import _abi as abi
    from hedgehog_jl_b200.engine import SimSpec
    from helpers import gbm_model
    with pytest.raises(NotImplementedError):
        oracle.lsm_american(gbm_model(), SimSpec(n_paths=10, n_steps=5, scheme=abi.HH_SCHEME_EXACT_TERMINAL), (100.0, -1.0), 2, 0.99)
