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
