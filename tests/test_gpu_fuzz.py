"""Randomised differential test of the specialised kernels against the oracle: 60 random models (including harsh ones:
Feller condition violated, vol of vol up to 2, correlation near +-1, few large steps, tiny and huge spots, negative rates).
The specialised kernels replace libm by tables, integer clamps and seeded square roots, so this is where a range or
sign assumption would show."""
import math
import os

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err

pytestmark = pytest.mark.gpu
# HH_FUZZ_SCALE=8 python -m pytest tests/test_gpu_fuzz.py: eight times as many random cases of every family (a soak run; the
# default keeps the suite at a few seconds)
SCALE = int(os.environ.get("HH_FUZZ_SCALE", "1"))


def _log_bounds(run, m, ref, fin):
    """Per-path bounds (0.99 quantile, maximum) on |log gpu - log oracle| that follow the CONDITIONING of the case: the oracle is
    run again with kappa (or sigma) one ulp larger, and what that moves is what no implementation can be held to. Harsh
    models (kappa dt > 1: the Euler step overshoots into the truncation on every step) move by 4e-9 / 2e-6 under that
    change; benign ones by < 1e-13, and the floor 1e-10 / 3e-6 applies."""
    name = "kappa" if getattr(m, "kind", 0) == abi.HH_MODEL_HESTON else "sigma"
    old = getattr(m, name)
    setattr(m, name, float(np.nextafter(old, np.inf)))
    try:
        alt = run(m)
    finally:
        setattr(m, name, old)
    ok = fin & np.isfinite(alt) & (alt > 0)
    d = np.abs(np.log(alt[ok]) - np.log(ref[ok])) if ok.any() else np.zeros(1)
    # x64: the two implementations differ by a rounding in EVERY step, not by one ulp of one input
    return max(1e-10, 64.0 * float(np.quantile(d, 0.99))), max(3e-6, 64.0 * float(d.max()))


def _lsm_flips(oracle, m, sim, K, cp, deg, D, tg, to, price_o, allowance, G):
    """Stopping decisions that differ from the oracle's. The kernel fits by NORMAL equations in a Chebyshev variable (2 deg + 1
    streamed moment sums), the reference by QR on the design matrix, whose condition number is only the square root of the
    Gram matrix's. How well the Gram matrix is conditioned depends on how closely the interval mapped to [-1, 1] follows the
    data (hh_lsm_american: uab): with the reach taken from the sample size every one of 19 920 soak cases is inside the tie
    allowance. Should a case ever leave it, it is excused HERE only when the Gram matrix of the kernel's own basis is measured to
    be beyond 1e13 on some date — the corner where 2 deg + 1 moment sums in binary64 cannot carry the fit."""
    flips = int(np.sum(tg != to))
    if flips <= allowance:
        return flips
    from scipy.stats import norm
    n_dates = G.shape[1] - 1
    worst = 0.0
    zn = min(max(float(norm.isf(1.0 / G.shape[0])), 3.0), 6.0) + (0.3 if m.kind == abi.HH_MODEL_HESTON else 0.0)   # hh_lsm_american: uab
    for t in range(1, n_dates):
        ty = m.T * t / n_dates
        if m.kind == abi.HH_MODEL_HESTON:
            w = -math.expm1(-m.kappa * ty) / m.kappa
            var_t = max(m.theta * ty + (m.V0 - m.theta) * w, 0.0)
        else:
            var_t = m.sigma ** 2 * ty
        med, reach = m.S0 * math.exp(m.r * ty - 0.5 * var_t), math.exp(zn * math.sqrt(var_t))
        lo, hi = (min(med / reach, 0.9 * K), K) if cp < 0 else (K, max(med * reach, 1.1 * K))
        s = G[:, t]
        s = s[cp * (s - K) > 0]
        if s.size <= deg + 1:
            continue
        T = np.polynomial.chebyshev.chebvander(2.0 * (s - lo) / (hi - lo) - 1.0, deg)
        worst = max(worst, float(np.linalg.cond(T.T @ T)))
    assert worst > 1e13, ("decisions differ although the Gram matrix is well conditioned", flips, worst)
    return 0


def _random_heston(rng):
    corr = rng.choice(["cholesky", "sym_sqrt", "svd"])
    return heston_model(S0=float(10 ** rng.uniform(-3, 5)), r=float(rng.uniform(-0.05, 0.2)), T=float(rng.uniform(0.02, 10.0)),
                        V0=float(10 ** rng.uniform(-4, 0)), kappa=float(10 ** rng.uniform(-2, 1.2)), theta=float(10 ** rng.uniform(-4, 0)),
                        xi=float(rng.uniform(0.0, 2.0)), rho=float(rng.uniform(-0.999, 0.999)), corr=corr,
                        split=bool(rng.integers(0, 2)))


@pytest.mark.parametrize("seed", range(30 * SCALE))
def test_heston_fast_kernels_on_random_models(cuda, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    m = _random_heston(rng)
    steps = int(rng.integers(1, 60))
    anti = int(rng.integers(0, 2))
    n = int(rng.integers(1, 3000))
    pay = [(m.S0 * float(k), float(cp)) for k, cp in ((0.8, 1.0), (1.0, -1.0), (1.3, 1.0))]
    D = math.exp(-m.r * m.T)
    sim = SimSpec(n_paths=n, n_steps=steps, vr=anti, base_seed=int(rng.integers(0, 2 ** 62)), path_offset=int(rng.integers(0, 2 ** 40)))
    rg, tg = cuda.mc_european(m, sim, pay, D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, pay, D, want_terminal=True)
    fin = np.isfinite(to)
    assert np.array_equal(np.isfinite(tg), fin)
    # log-space comparison: x = log S carries the 1e-16-per-step differences. A path whose variance lands within rounding
    # of the truncation kink sees them amplified by d sqrt(K2)/dK2 -> infinity (ill-conditioned in the scheme itself, for
    # the oracle as for the GPU), hence a quantile bound plus a looser bound on the worst path.
    dlog = np.abs(np.log(tg[fin]) - np.log(to[fin]))
    q99, worst = _log_bounds(lambda mm: oracle.mc_european(mm, sim, pay, D, want_terminal=True)[1], m, to, fin)
    assert np.quantile(dlog, 0.99) < q99, (np.quantile(dlog, 0.99), q99)
    assert dlog.max() < worst, (dlog.max(), worst)
    # f32 fast mode stays finite and close on the same model (loose: MUFU approximations, binary32 state)
    sim32 = SimSpec(n_paths=n, n_steps=steps, vr=anti, precision=abi.HH_PREC_F32, base_seed=sim.base_seed)
    r32, t32 = cuda.mc_european(m, sim32, pay, D, want_terminal=True)
    o32, u32 = oracle.mc_european(m, sim32, pay, D, want_terminal=True)
    ok = np.isfinite(u32) & (u32 > 0)
    assert np.all(np.isfinite(t32[ok]))
    assert np.median(np.abs(np.log(t32[ok]) - np.log(u32[ok]))) < 1e-3


@pytest.mark.parametrize("seed", range(15 * SCALE))
def test_gbm_fast_kernels_on_random_models(cuda, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    m = gbm_model(S0=float(10 ** rng.uniform(-3, 5)), r=float(rng.uniform(-0.05, 0.3)), sigma=float(rng.uniform(0.0, 3.0)),
                  T=float(rng.uniform(0.02, 10.0)))
    # the reference's exact step S += S (exp(y) - 1) cancels catastrophically once exp(y) << 1; keep sigma sqrt(dt) <= 0.6 so
    # that the comparison measures the kernels, not that formula (|y| > 1/2 still occurs: the libm path is exercised)
    steps = max(int(rng.integers(1, 40)), int(math.ceil(m.T * (m.sigma / 0.6) ** 2)))
    anti = int(rng.integers(0, 2))
    n = int(rng.integers(1, 3000))
    pay = [(m.S0, 1.0), (m.S0 * 1.2, -1.0)]
    for scheme in (abi.HH_SCHEME_EM, abi.HH_SCHEME_EXACT_STEPS):  # sigma sqrt(dt) up to 3: the exp(y) - 1 slow path too
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=anti, base_seed=int(rng.integers(0, 2 ** 62)))
        rg, tg = cuda.mc_european(m, sim, pay, 0.9, want_terminal=True)
        ro, to = oracle.mc_european(m, sim, pay, 0.9, want_terminal=True)
        good = np.isfinite(to) & (to > 0)
        dlog = np.abs(np.log(tg[good]) - np.log(to[good]))
        assert np.quantile(dlog, 0.99) < 1e-10 and dlog.max() < 1e-7, (np.quantile(dlog, 0.99), dlog.max())


@pytest.mark.parametrize("seed", range(15 * SCALE))
def test_heston_tangent_kernel_on_random_models(cuda, oracle, seed):
    rng = np.random.default_rng(3000 + seed)
    m = _random_heston(rng)
    m.T = float(rng.uniform(0.05, 2.0))
    m.xi = float(rng.uniform(0.05, 0.8))
    steps = int(rng.integers(2, 40))
    anti = int(rng.integers(0, 2))
    n = 1500
    _, dM = hh.corr_factor(m.rho, "cholesky")
    (m.m11, m.m12, m.m21, m.m22), _ = hh.corr_factor(m.rho, "cholesky")
    tans = [abi.hh_tangent(dS0=1.0), abi.hh_tangent(dV0=1.0), abi.hh_tangent(dr=1.0), abi.hh_tangent(dkappa=1.0),
            abi.hh_tangent(dtheta=1.0), abi.hh_tangent(dxi=1.0), abi.hh_tangent(dm11=dM[0], dm12=dM[1], dm21=dM[2], dm22=dM[3])]
    nt = int(rng.integers(1, 8))
    pay = [(m.S0 * 0.9, 1.0), (m.S0 * 1.1, -1.0)]
    sim = SimSpec(n_paths=n, n_steps=steps, vr=anti, base_seed=int(rng.integers(0, 2 ** 62)))
    sg, _ = cuda.tangent_sums(m, tans[:nt], sim, pay)
    so, _ = oracle.tangent_sums(m, tans[:nt], sim, pay)
    # sums of squares of tangents can be huge when the variance sits at zero (d sqrt -> infinity); compare the first moments
    cols = [0, 1] + [2 + q for q in range(nt)]
    scale = np.maximum(np.abs(so[:, cols]), 1e-8 * np.abs(so[:, cols]).max() + 1e-300)
    assert np.max(np.abs(sg[:, cols] - so[:, cols]) / scale) < 3e-6   # (3 of 1800 soak cases sit at 1.1e-6 - 1.2e-6)


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_lsm_on_random_contracts(cuda, oracle, seed):
    """Random American contracts: stored paths 1e-12, stopping decisions equal to the QR oracle's except for a handful of
    ties, price within 1e-6 (persistent kernel, Chebyshev normal equations, time-0 money vs the oracle's literal form)."""
    rng = np.random.default_rng(4000 + seed)
    m = gbm_model(S0=float(rng.uniform(20, 200)), r=float(rng.uniform(0.0, 0.12)), sigma=float(rng.uniform(0.05, 0.6)),
                  T=float(rng.uniform(0.1, 3.0)))
    steps = int(rng.integers(2, 30)) if rng.random() < 0.75 else int(rng.integers(30, 150))
    deg = int(rng.integers(1, 7))
    anti = int(rng.integers(0, 2))
    cp = float(rng.choice([-1.0, -1.0, 1.0]))
    K = m.S0 * float(rng.uniform(0.85, 1.15))
    n = int(rng.integers(5000, 30000))
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=anti, base_seed=int(rng.integers(0, 2 ** 62)))
    D = math.exp(-m.r * m.T / steps)
    og, tg, vg, pg = cuda.lsm_american(m, sim, (K, cp), deg, D, want_stopping=True, want_paths=True)
    oo, to, vo, po = oracle.lsm_american(m, sim, (K, cp), deg, D, want_stopping=True, want_paths=True)
    assert rel_err(pg, po) < 1e-12
    # ties: a column whose exercise value equals the fitted continuation value to rounding, on any of its dates
    allowance = max(3, 3e-4 * len(to) * max(1.0, steps / 30))
    excused = int(np.sum(tg != to)) > allowance
    flips = _lsm_flips(oracle, m, sim, K, cp, deg, D, tg, to, oo.price, allowance, po)
    assert flips <= allowance, (flips, len(to), steps)
    if excused:
        return
    price_o = oo.price
    # a flipped decision replaces one column's cash flow by another realisation: O(price) / columns each
    # (deep out of the money the price is a few cash flows: the bound is per column, not relative to the price)
    tol = 1e-9 * max(abs(price_o), 1e-3) if flips == 0 else flips * (0.2 * K / len(to) + 1e-4 * abs(price_o))
    assert abs(og.price - price_o) <= tol, (og.price, price_o, flips)
    assert og.n_dates_skipped == oo.n_dates_skipped


@pytest.mark.parametrize("seed", range(20 * SCALE))
def test_path_dependent_kernels_on_random_models(cuda, oracle, seed):
    """The specialised path-dependent Heston kernel (folded step, table-driven exp) and the generic one (log-GBM) on
    harsh random models: per-column statistics agree in log space (quantile bound + looser worst-path bound, as above),
    non-finite columns coincide, and the sums of the continuous payoffs agree."""
    rng = np.random.default_rng(3000 + seed)
    heston = bool(seed % 2 == 0)
    m = _random_heston(rng) if heston else gbm_model(S0=float(10 ** rng.uniform(-3, 5)), r=float(rng.uniform(-0.05, 0.3)),
                                                     sigma=float(rng.uniform(0.0, 2.0)), T=float(rng.uniform(0.02, 10.0)))
    every = int(rng.choice([1, 2, 5]))
    steps = every * int(rng.integers(1, 20))
    anti = int(rng.integers(0, 2))
    n = int(rng.integers(1, 3000))
    kw = dict(base_seed=int(rng.integers(0, 2 ** 62)), path_offset=int(rng.integers(0, 2 ** 40)))
    if seed % 5 == 0:
        kw = dict(seeds=rng.integers(0, 2 ** 63, size=n, dtype=np.uint64))
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=anti, **kw)
    pays = [(abi.HH_PD_ASIAN_ARITH, m.S0, 1.0, 0.0, 0.0), (abi.HH_PD_ASIAN_GEOM, m.S0 * 1.1, -1.0, 0.0, 0.0),
            (abi.HH_PD_UP_OUT, m.S0, 1.0, m.S0 * 1.5, 0.0), (abi.HH_PD_DOWN_IN, m.S0, -1.0, m.S0 * 0.7, 0.0)]
    rg, sg = cuda.mc_path_dependent(m, sim, pays, 1.0, every, want_stats=True)
    ro, so = oracle.mc_path_dependent(m, sim, pays, 1.0, every, want_stats=True)
    fin = np.isfinite(so) & (so > 0)
    assert np.array_equal(np.isfinite(sg) & (sg > 0), fin)
    dlog = np.abs(np.log(sg[fin]) - np.log(so[fin]))
    q99, worst = _log_bounds(lambda mm: oracle.mc_path_dependent(mm, sim, pays, 1.0, every, want_stats=True)[1], m, so, fin)
    assert np.quantile(dlog, 0.99) < q99, (np.quantile(dlog, 0.99), q99)
    assert dlog.max() < worst, (dlog.max(), worst)
    if fin.all():
        for g, o in zip(rg[:2], ro[:2]):   # continuous payoffs: no decision to flip
            assert g.sum == pytest.approx(o.sum, rel=1e-6, abs=1e-6 * m.S0)
        assert rg[0].n_nonfinite == 0


WIDE = os.environ.get("HH_FUZZ_WIDE", "0") == "1"   # soak option: vol of vol down to 0.03 (orders ~1e3), up to 52 dates


def _random_bk_case(seed):
    rng = np.random.default_rng(5000 + seed)
    kappa = float(10 ** rng.uniform(-1, 1.3 if WIDE else 1))
    theta = float(10 ** rng.uniform(-2.5, -0.5))
    xi = float(10 ** rng.uniform(-1.5 if WIDE else -1.2, 0.2))
    pars = dict(kappa=kappa, theta=theta, xi=xi, rho=float(rng.uniform(-0.95, 0.5)), V0=theta * float(rng.uniform(0.3, 2.5)),
                r=float(rng.uniform(-0.01, 0.08)))
    return (pars, int(rng.integers(20, 730)), int(rng.choice([1, 4, 12, 52] if WIDE else [1, 1, 4])),
            100.0 * float(rng.uniform(0.9, 1.1)))


@pytest.mark.parametrize("seed", range(24 * SCALE))
def test_broadie_kaya_on_random_models(cuda, seed):
    """The exact sampler over the parameter space (Bessel orders nu = 2 kappa theta / xi^2 - 1 from -0.99 — degrees of freedom
    0.02, the variance sits at zero most of the time — to 157, horizons from three weeks to two years, one and four
    dates): no non-finite trajectory, next to no inversion fallbacks, and the price within 4 standard errors of Carr-Madan
    (whose own truncation error, visible for orders near -1, is added to the bound).

    The Fourier grid is h = pi / (mean + 12 sd) here instead of the reference's default 5 (sample_from_cf.jl:37): with 5
    the periodised CDF cuts the right tail of the integrated variance, which for degrees of freedom ~0.02 and vol of vol ~1
    biases the price by -0.3 % (7 standard errors at 2e6 trajectories; tools/bk_bias_probe.py,
    profiles/r2_l_bk_bias_probe.txt) — a property of the algorithm's defaults, not of an implementation, and this test is
    about the implementation."""
    from oracle import anchors as A
    pars, days, steps, K = _random_bk_case(seed)
    n = 200_000
    T = days / 365
    m = heston_model(S0=100.0, T=T, **pars)
    cfg = abi.hh_bk_config()
    cuda.lib.hh_default_bk_config(cfg)
    cfg.n_std = 12
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=77 + seed, bk=cfg)
    res, ens = cuda.mc_european(m, sim, [(K, 1.0)], math.exp(-pars["r"] * T), want_terminal=True)
    st = cuda.bk_last_stats()
    args = (100.0, K, pars["r"], T, pars["V0"], pars["kappa"], pars["theta"], pars["xi"], pars["rho"])
    cm, cm2 = A.heston_price(*args, bound=600.0), A.heston_price(*args, bound=200.0)
    assert res[0].n_nonfinite == 0
    assert st["n_fallback"] <= 2e-3 * n * steps, st
    assert np.all(np.isfinite(ens)) and np.all(ens > 0)
    price, se = res[0].price, res[0].std_error
    # Positive correlation with a large vol of vol: E[S_T^2] is infinite from some horizon on (Andersen & Piterbarg 2007:
    # kappa - 2 rho xi < 0 or (kappa - 2 rho xi)^2 < 2 xi^2), the payoff has no variance, and "4 standard errors" means
    # nothing (soak case: 16.8 +- 0.36 against 13.2 from one huge trajectory). Finite and positive is all that is asked there.
    b = pars["kappa"] - 2.0 * pars["rho"] * pars["xi"]
    if b < 0.0 or b * b < 2.0 * pars["xi"] ** 2:
        return
    assert abs(price - cm) < 4.0 * se + 3.0 * abs(cm - cm2) + 2e-4 * max(cm, 0.05), (price, cm, se, pars)


@pytest.mark.parametrize("seed", range(16 * SCALE))
def test_philox64_streams_on_random_models(cuda, oracle, seed):
    """The opt-in HH_RNG_PHILOX_64 stream on harsh random models: the Heston Euler-Maruyama kernel (one block per two
    steps) per path in log space, and the exact GBM generator of the LSM (one block per four steps) through the stored
    paths and the stopping decisions."""
    rng = np.random.default_rng(6000 + seed)
    m = _random_heston(rng)
    steps = int(rng.integers(1, 60))
    anti = int(rng.integers(0, 2))
    n = int(rng.integers(1, 3000))
    sim = SimSpec(n_paths=n, n_steps=steps, vr=anti, rng_mode=abi.HH_RNG_PHILOX_64, base_seed=int(rng.integers(0, 2 ** 62)),
                  path_offset=int(rng.integers(0, 2 ** 40)))
    pay = [(m.S0, 1.0), (m.S0 * 0.9, -1.0)]
    rg, tg = cuda.mc_european(m, sim, pay, 0.97, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, pay, 0.97, want_terminal=True)
    fin = np.isfinite(to)
    assert np.array_equal(np.isfinite(tg), fin)
    dlog = np.abs(np.log(tg[fin]) - np.log(to[fin]))
    q99, worst = _log_bounds(lambda mm: oracle.mc_european(mm, sim, pay, 0.97, want_terminal=True)[1], m, to, fin)
    assert np.quantile(dlog, 0.99) < q99 and dlog.max() < worst, (np.quantile(dlog, 0.99), q99, dlog.max(), worst)
    g = gbm_model(S0=float(rng.uniform(20, 200)), r=float(rng.uniform(0.0, 0.12)), sigma=float(rng.uniform(0.05, 0.9)),
                  T=float(rng.uniform(0.1, 3.0)))
    steps = int(rng.integers(2, 30))
    sim = SimSpec(n_paths=int(rng.integers(5000, 20000)), n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=anti,
                  rng_mode=abi.HH_RNG_PHILOX_64, base_seed=int(rng.integers(0, 2 ** 62)))
    D = math.exp(-g.r * g.T / steps)
    K = g.S0 * float(rng.uniform(0.9, 1.1))
    og, tg, vg, pg = cuda.lsm_american(g, sim, (K, -1.0), 3, D, want_stopping=True, want_paths=True)
    oo, to, vo, po = oracle.lsm_american(g, sim, (K, -1.0), 3, D, want_stopping=True, want_paths=True)
    assert rel_err(pg, po) < 1e-12
    flips = int(np.sum(tg != to))
    assert flips <= max(3, 3e-4 * len(to)), (flips, len(to))
    assert abs(og.price - oo.price) <= (1e-9 if flips == 0 else 2e-5) * max(abs(oo.price), 1e-3)


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_lsm_under_heston_on_random_models(cuda, oracle, seed):
    """American puts and calls under random Heston models through the log-space generator: stored spots
    1e-10, decisions equal to the oracle's up to counted ties."""
    rng = np.random.default_rng(7000 + seed)
    m = heston_model(S0=float(rng.uniform(20, 200)), r=float(rng.uniform(0.0, 0.1)), T=float(rng.uniform(0.1, 2.0)),
                     V0=float(10 ** rng.uniform(-2.5, -0.5)), kappa=float(10 ** rng.uniform(-1, 1)), theta=float(10 ** rng.uniform(-2.5, -0.5)),
                     xi=float(rng.uniform(0.05, 1.2)), rho=float(rng.uniform(-0.95, 0.5)), split=bool(rng.integers(0, 2)))
    steps = int(rng.integers(2, 40))
    anti = int(rng.integers(0, 2))
    cp = float(rng.choice([-1.0, -1.0, 1.0]))
    K = m.S0 * float(rng.uniform(0.9, 1.1))
    deg = int(rng.integers(1, 7))
    sim = SimSpec(n_paths=int(rng.integers(5000, 30000)), n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=anti,
                  base_seed=int(rng.integers(0, 2 ** 62)), path_offset=int(rng.integers(0, 2 ** 40)))
    D = math.exp(-m.r * m.T / steps)
    og, tg, vg, pg = cuda.lsm_american(m, sim, (K, cp), deg, D, want_stopping=True, want_paths=True)
    oo, to, vo, po = oracle.lsm_american(m, sim, (K, cp), deg, D, want_stopping=True, want_paths=True)
    fin = np.isfinite(po) & (po > 0)
    assert np.array_equal(np.isfinite(pg) & (pg > 0), fin)
    dlog = np.abs(np.log(pg[fin]) - np.log(po[fin]))
    assert np.quantile(dlog, 0.999) < 1e-10 and dlog.max() < 1e-6, (np.quantile(dlog, 0.999), dlog.max())
    allowance = max(3, 5e-4 * len(to))
    excused = int(np.sum(tg != to)) > allowance
    flips = _lsm_flips(oracle, m, sim, K, cp, deg, D, tg, to, oo.price, allowance, po)
    assert flips <= allowance, (flips, len(to))
    if excused:
        return
    price_o = oo.price
    # a flipped decision replaces one column's cash flow by another realisation: O(price) / columns each
    tol = 1e-9 * max(abs(price_o), 1e-3) if flips == 0 else flips * (0.2 * K / len(to) + 1e-4 * abs(price_o))
    assert abs(og.price - price_o) <= tol, (og.price, price_o, flips)


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_broadie_kaya_monitoring_dates_on_random_models(cuda, seed):
    """Path statistics on exact Broadie-Kaya dates over the parameter space: S_T is the European path's to the last bit,
    min <= geometric <= arithmetic <= max, everything finite, and — the scheme being exact — the discounted spot is a
    martingale on EVERY date: mean of the arithmetic average = S0 mean_k e^{r t_k} within 4 standard errors."""
    pars, days, _, _ = _random_bk_case(100 + seed)
    rng = np.random.default_rng(8000 + seed)
    dates = int(rng.choice([2, 5, 13, 52]))
    T = days / 365
    m = heston_model(S0=100.0, T=T, **pars)
    n = 100_000
    cfg = abi.hh_bk_config()
    cuda.lib.hh_default_bk_config(cfg)
    cfg.n_std = 12
    sim = SimSpec(n_paths=n, n_steps=dates, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=500 + seed, bk=cfg)
    pays = [(abi.HH_PD_VANILLA, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_ASIAN_ARITH, 0.0, 1.0, 0.0, 0.0), (abi.HH_PD_ASIAN_GEOM, 100.0, -1.0, 0.0, 0.0)]
    res, st = cuda.mc_path_dependent(m, sim, pays, 1.0, 1, want_stats=True)
    eur, term = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    assert np.array_equal(st[0], term)
    assert np.all(np.isfinite(st)) and np.all(st > 0)
    assert np.all(st[4] <= st[2] * (1 + 1e-13)) and np.all(st[2] <= st[1] * (1 + 1e-13)) and np.all(st[1] <= st[3] * (1 + 1e-13))
    assert all(r.n_nonfinite == 0 for r in res)
    # strike 0 call on the arithmetic average = its mean
    tk = T * np.arange(1, dates + 1) / dates
    want = 100.0 * np.mean(np.exp(pars["r"] * tk))
    assert abs(res[1].price - want) < 4.0 * res[1].std_error + 1e-4 * want, (res[1].price, want, res[1].std_error, pars, dates)


@pytest.mark.parametrize("seed", range(10 * SCALE))
def test_second_order_sums_and_control_variates_on_random_models(cuda, oracle, seed):
    """Gamma by bumped payoffs on the kernel's own trajectories (second_sums, greeks_problem.jl:395-412) and the
    Black-Scholes control-variate contracts, on random Heston models of moderate harshness: the sums the host layer
    combines agree with the oracle's (three re-simulations / a log-GBM trajectory on the same increments)."""
    rng = np.random.default_rng(9000 + seed)
    m = _random_heston(rng)
    m.S0 = float(rng.uniform(20, 300))
    m.T = float(rng.uniform(0.05, 2.0))
    m.xi = float(rng.uniform(0.05, 0.8))
    m.kappa = float(rng.uniform(0.2, 6.0))
    m.V0, m.theta = float(rng.uniform(0.01, 0.2)), float(rng.uniform(0.01, 0.2))
    (m.m11, m.m12, m.m21, m.m22), dM = hh.corr_factor(m.rho, "cholesky")
    steps = int(rng.integers(2, 40))
    anti = int(rng.integers(0, 2))
    n = 2000
    pay = [(m.S0 * 0.9, 1.0), (m.S0, 1.0), (m.S0 * 1.1, -1.0)]
    tans = [abi.hh_tangent(dS0=1.0), abi.hh_tangent(dV0=1.0)]
    sim = SimSpec(n_paths=n, n_steps=steps, vr=anti, base_seed=int(rng.integers(0, 2 ** 62)))
    eps = m.S0 * float(rng.uniform(1e-3, 2e-2))
    sg, _, g2 = cuda.tangent_sums(m, tans, sim, pay, spot_bump=eps)
    so, _, o2 = oracle.tangent_sums(m, tans, sim, pay, spot_bump=eps)
    scale = np.abs(o2).max(axis=0, keepdims=True) + 1e-300
    assert np.max(np.abs(g2 - o2) / scale) < 1e-7, np.max(np.abs(g2 - o2) / scale)
    cols = [0, 1, 2, 3]
    sc = np.maximum(np.abs(so[:, cols]), 1e-8 * np.abs(so[:, cols]).max() + 1e-300)
    assert np.max(np.abs(sg[:, cols] - so[:, cols]) / sc) < 1e-6
    # control variates (Heston Euler-Maruyama only): the control, and the payoff minus beta times the control
    pays = [(abi.HH_PD_VANILLA, m.S0, 1.0, 0.0, 0.0), (abi.HH_PD_BS_CONTROL, m.S0, 1.0, 0.0, 0.0),
            (abi.HH_PD_VANILLA_MINUS_BS, m.S0 * 0.95, -1.0, 0.0, float(rng.uniform(0.3, 1.2)))]
    rg, _ = cuda.mc_path_dependent(m, sim, pays, 0.98, 1)
    ro, _ = oracle.mc_path_dependent(m, sim, pays, 0.98, 1)
    for g, o in zip(rg, ro):
        assert g.sum == pytest.approx(o.sum, rel=1e-7, abs=1e-7 * m.S0)
        assert g.sumsq == pytest.approx(o.sumsq, rel=1e-6, abs=1e-7 * m.S0 ** 2)
        assert g.n_nonfinite == 0
