"""GPU pathwise Greeks (in-kernel dual numbers) against the oracle's hand-derived tangent recursions,
against finite differences on common random numbers, and against analytic Black-Scholes Greeks."""
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model

pytestmark = pytest.mark.gpu


def _tan(**kw):
    t = abi.hh_tangent()
    for k, v in kw.items():
        setattr(t, k, v)
    return t


def heston_dirs(m, corr="cholesky"):
    _, dM = hh.corr_factor(m.rho, corr)
    return [_tan(dS0=1.0), _tan(dV0=1.0), _tan(dr=1.0), _tan(dkappa=1.0), _tan(dtheta=1.0), _tan(dxi=1.0),
            _tan(dm11=dM[0], dm12=dM[1], dm21=dM[2], dm22=dM[3])]


@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("ntan", [1, 2, 3, 7])
def test_heston_tangent_sums_match_oracle(cuda, oracle, anti, ntan):
    n, steps = 6000, 50
    m = heston_model()
    sim = SimSpec(n_paths=n, n_steps=steps, vr=int(anti), base_seed=77)
    pay = [(k, 1.0) for k in (80.0, 100.0, 120.0)] + [(100.0, -1.0)]
    tans = heston_dirs(m)[:ntan]
    sg, _ = cuda.tangent_sums(m, tans, sim, pay)
    so, _ = oracle.tangent_sums(m, tans, sim, pay)
    assert np.allclose(sg, so, rtol=1e-10, atol=1e-9 * np.abs(so).max())


@pytest.mark.parametrize("scheme,steps", [(abi.HH_SCHEME_EM, 20), (abi.HH_SCHEME_EXACT_STEPS, 20), (abi.HH_SCHEME_EXACT_TERMINAL, 1)])
@pytest.mark.parametrize("anti", [False, True])
def test_gbm_tangent_sums_match_oracle(cuda, oracle, scheme, steps, anti):
    m = gbm_model(S0=1.0, r=0.03, sigma=1.0, T=366 / 365)
    sim = SimSpec(n_paths=5000, n_steps=steps, scheme=scheme, vr=int(anti), base_seed=42)
    tans = [_tan(dS0=1.0), _tan(dsigma=1.0), _tan(dr=1.0)]
    sg, _ = cuda.tangent_sums(m, tans, sim, [(1.0, 1.0), (0.8, -1.0)])
    so, _ = oracle.tangent_sums(m, tans, sim, [(1.0, 1.0), (0.8, -1.0)])
    assert np.allclose(sg, so, rtol=1e-10, atol=1e-9 * np.abs(so).max())


@pytest.mark.parametrize("xi,tol", [(0.1, 5e-4), (0.3, 1e-1)])
def test_heston_tangent_vs_finite_difference_crn(cuda, xi, tol):
    """Central differences of the GPU price on common random numbers converge to the in-kernel tangent.
    With xi = 0.1 the variance never reaches the max(v, 0) kink and FD and AD agree to 5e-4 (the payoff kink at the strike is
    all that is left of the FD error); with the reference's
    xi = 0.3 about 1% of 40-step paths cross it, where the pathwise estimator (ForwardDiff's too) drops the kink term."""
    n, steps = 200_000, 40
    base = dict(S0=100.0, r=0.03, T=1.0, V0=0.04, kappa=2.0, theta=0.04, xi=xi, rho=-0.7)
    m = heston_model(**base)
    sim = SimSpec(n_paths=n, n_steps=steps, base_seed=9)
    pay = [(100.0, 1.0)]
    sums, _ = cuda.tangent_sums(m, heston_dirs(m), sim, pay)
    ad = sums[0, 2:2 + 7] / n  # d mean(payoff) / d param, discounting left out on both sides
    names = ["S0", "V0", "r", "kappa", "theta", "xi", "rho"]
    for i, name in enumerate(names):
        h = 1e-4 * max(abs(base[name]), 1.0)
        up, dn = dict(base), dict(base)
        up[name] += h
        dn[name] -= h
        pu, _ = cuda.mc_european(heston_model(**up), sim, pay, 1.0)
        pd, _ = cuda.mc_european(heston_model(**dn), sim, pay, 1.0)
        fd = (pu[0].price - pd[0].price) / (2 * h)
        assert abs(fd - ad[i]) <= tol * max(abs(ad[i]), 1e-2), (name, fd, ad[i])


def test_reference_mc_greeks_test(cuda):
    """test/agreement/greeks_agreement.jl:170-241: GBM exact, 100 000 paths, S=K=1, sigma=1, r=.03 vs analytic."""
    import datetime as dt
    from oracle import anchors as A
    payoff = hh.VanillaOption(1.0, dt.date(2021, 1, 1), hh.European(), hh.Call(), hh.Spot())
    market = hh.BlackScholesInputs(dt.date(2020, 1, 1), 0.03, 1.0, 1.0)
    prob = hh.PricingProblem(payoff, market)
    seeds = np.random.Generator(np.random.Philox(42)).integers(1, 10**9, size=100_000, dtype=np.uint64)
    mc = hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(100_000, seeds=seeds))
    T = 366 / 365
    an = A.bs_greeks(1.0, 1.0, 0.03, 1.0, T)
    price = hh.solve(prob, mc, engine=cuda).price
    # Q1: with T = 366/365 the reference's sqrt(alpha) mean shifts the price by ~1e-4 relative; inside its 3e-2
    assert price == pytest.approx(A.bs_price(1.0, 1.0, 0.03, 1.0, T), rel=3e-2)
    delta = hh.solve(hh.GreekProblem(prob, hh.SpotLens()), hh.ForwardAD(), mc, engine=cuda).greek
    assert delta == pytest.approx(an["delta"], rel=3e-2)
    gamma = hh.solve(hh.SecondOrderGreekProblem(prob, hh.SpotLens(), hh.SpotLens()), hh.FiniteDifference(1e-1), mc, engine=cuda).greek
    assert gamma == pytest.approx(an["gamma"], rel=2e-1)
    vega = hh.solve(hh.GreekProblem(prob, hh.VolLens(1, 1)), hh.ForwardAD(), mc, engine=cuda).greek
    assert vega == pytest.approx(an["vega"], rel=1e-1)
    rho = hh.solve(hh.GreekProblem(prob, hh.ZeroRateSpineLens(1)), hh.ForwardAD(), mc, engine=cuda).greek
    assert rho == pytest.approx(an["rho"], rel=3e-2)
    batch = hh.solve(hh.BatchGreekProblem(prob, (hh.SpotLens(), hh.VolLens(1, 1), hh.ZeroRateSpineLens(1))), hh.ForwardAD(), mc, engine=cuda)
    assert batch[hh.SpotLens()] == pytest.approx(delta, rel=1e-12)
    assert batch[hh.VolLens(1, 1)] == pytest.approx(vega, rel=1e-12)


@pytest.mark.parametrize("corr", ["cholesky", "sym_sqrt"])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("anti", [False, True])
def test_specialised_heston_tangent_kernel_equals_generic_dual_kernel(cuda, oracle, anti, split, corr):
    """Native-RNG Heston tangents run the written-out recursion (heston_tangent_kernel); the same normals fed in parity mode
    run the generic Dual<P> template. General directions (linear combinations, a factor with d m11 != 0) exercise every term,
    including directions the host classifies as trivial (only dS0 / dr)."""
    n, steps = 3000, 40
    m = heston_model(corr=corr, split=split, xi=0.6)
    _, dM = hh.corr_factor(m.rho, corr)
    tans = [_tan(dS0=1.0, dr=0.3), _tan(dV0=1.0, dkappa=-0.7, dtheta=0.2), _tan(dr=1.0),
            _tan(dxi=1.0, dm11=dM[0], dm12=dM[1], dm21=dM[2], dm22=dM[3]), _tan(dS0=2.0, dxi=0.1, dtheta=1.0)]
    pay = [(90.0, 1.0), (100.0, 1.0), (110.0, -1.0)]
    sim = SimSpec(n_paths=n, n_steps=steps, vr=int(anti), base_seed=5)
    z = oracle.fill_normals(m, sim)
    sim_par = SimSpec(n_paths=n, n_steps=steps, vr=int(anti), rng_mode=abi.HH_RNG_NORMALS, normals=z)
    s_special, _ = cuda.tangent_sums(m, tans, sim, pay)
    s_generic, _ = cuda.tangent_sums(m, tans, sim_par, pay)
    s_oracle, _ = oracle.tangent_sums(m, tans, sim, pay)
    scale = np.abs(s_oracle).max()
    assert np.allclose(s_special, s_generic, rtol=1e-10, atol=1e-10 * scale)
    assert np.allclose(s_special, s_oracle, rtol=1e-10, atol=1e-9 * scale)


# ---- second order in the spot (gamma): SecondOrderGreekProblem(spot, spot), greeks_problem.jl:395-412 ----------------------
@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("kind", ["heston", "heston_parity", "gbm_em", "gbm_steps", "gbm_terminal"])
def test_second_order_sums_equal_three_solves_of_the_oracle(cuda, oracle, kind, anti):
    """The kernel evaluates the bumped payoffs on the trajectories it has (every scheme is linear in S0); the oracle
    re-simulates every trajectory at S0 - eps, S0, S0 + eps like the reference's FiniteDifference does. Same sums."""
    n, eps = 6000, 0.35
    pay = [(k, 1.0) for k in (80.0, 100.0, 120.0)] + [(100.0, -1.0)]
    if kind.startswith("heston"):
        m, steps, scheme = heston_model(), 40, abi.HH_SCHEME_EM
        tans = heston_dirs(m)[:3]
    else:
        m = gbm_model(T=366 / 365)
        steps = 1 if kind == "gbm_terminal" else 20
        scheme = {"gbm_em": abi.HH_SCHEME_EM, "gbm_steps": abi.HH_SCHEME_EXACT_STEPS, "gbm_terminal": abi.HH_SCHEME_EXACT_TERMINAL}[kind]
        tans = [_tan(dS0=1.0), _tan(dsigma=1.0)]
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=int(anti), base_seed=31)
    if kind == "heston_parity":   # the generic Dual<P> template instead of the specialised tangent kernel
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=int(anti), rng_mode=abi.HH_RNG_NORMALS,
                      normals=oracle.fill_normals(m, sim))
    sg, _, g2 = cuda.tangent_sums(m, tans, sim, pay, spot_bump=eps)
    so, _, o2 = oracle.tangent_sums(m, tans, sim, pay, spot_bump=eps)
    assert np.allclose(sg, so, rtol=1e-10, atol=1e-9 * np.abs(so).max())
    # sd is a difference of O(10) payoffs: absolute agreement at the rounding level of the payoffs
    assert np.allclose(g2[:, 0], o2[:, 0], rtol=1e-9, atol=1e-9)
    assert np.allclose(g2[:, 1], o2[:, 1], rtol=1e-8, atol=1e-9)
    assert np.allclose(g2[:, 2:], o2[:, 2:], rtol=1e-9, atol=1e-10)
    # first-order sums do not change when the second-order ones are requested
    s0, _ = cuda.tangent_sums(m, tans, sim, pay)
    assert np.array_equal(s0, sg)
    with pytest.raises(ValueError):
        cuda.tangent_sums(m, tans, sim, pay, spot_bump=m.S0 + 1.0)


def test_heston_delta_gamma_vega_against_carr_madan(cuda):
    """Heston delta / gamma / vega (dV0) of the C5 problem on a strike grid, from ONE launch, within 3 standard errors (+ the
    Euler-Maruyama bias allowance) of finite differences of the Carr-Madan price (tests/golden/config_anchors.json)."""
    import datetime as dt
    import json
    import os
    anchors = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "config_anchors.json")))["c5"]
    strikes = np.array(anchors["strikes"])
    payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
    market = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    prob = hh.PricingProblem(payoff, market)
    mc = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(4_000_000, steps=252, base_seed=17), ensemble=False)
    lenses = [hh.SpotLens(), hh.optic("market_inputs.V0")]
    prices, g, se, sec = hh.strike_grid_greeks(prob, strikes, lenses, mc, engine=cuda, gamma_bump=0.5)
    sel = slice(8, 56)  # strikes 70..130: away from the wings where the estimators have few in-the-money paths
    for name, est, err, ref, bias in (("delta", g[:, 0], se[:, 0], np.array(anchors["d_S0"]), 2e-3),
                                      ("vega", g[:, 1], se[:, 1], np.array(anchors["d_V0"]), 0.6),
                                      ("gamma_fd", sec["fd"], sec["fd_stderr"], np.array(anchors["d2_S0"]), 4e-4),
                                      ("gamma_pw", sec["pathwise"], sec["pathwise_stderr"], np.array(anchors["d2_S0"]), 4e-4)):
        z = np.abs(est[sel] - ref[sel]) / (3 * err[sel] + bias)
        assert z.max() < 1.0, (name, float(z.max()), int(z.argmax()))
    assert np.all(np.abs(prices[sel] - np.array(anchors["price"])[sel]) < 3 * 0.008 + 0.012)
    # both gamma estimators agree with each other and the delta-difference form has the smaller standard error
    assert np.all(np.abs(sec["fd"][sel] - sec["pathwise"][sel]) < 4 * np.hypot(sec["fd_stderr"][sel], sec["pathwise_stderr"][sel]))
    # SecondOrderGreekProblem through solve takes the same launch
    gam = hh.solve(hh.SecondOrderGreekProblem(hh.PricingProblem(hh.VanillaOption(float(strikes[32]), dt.date(2020, 12, 31), hh.European(),
                                                                                hh.Call(), hh.Spot()), market),
                                              hh.SpotLens(), hh.SpotLens()), hh.FiniteDifference(0.5), mc, engine=cuda)
    assert gam.greek == pytest.approx(sec["fd"][32], rel=1e-12)
