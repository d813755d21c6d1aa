"""GPU parity tests for the Broadie-Kaya exact Heston sampler (heston.jl:125-300, sample_from_cf.jl) through the C ABI.

Protocol: the deterministic pieces — log I_nu(z), the characteristic function of the integrated variance,
the finite-difference moments, the Fourier-series CDF — are compared with the scipy/AMOS oracle (oracle/bk_ref.py) at
tight tolerance; the inversion must return a root that the reference's own acceptance test (|F(x) - u| <= 1e-4,
sample_from_cf.jl:110,119) accepts; the samplers are checked in distribution (KS against scipy's noncentral chi-square)
and the prices within 3 standard errors of Carr-Madan (the reference's own agreement tests, montecarlo_heston.jl:151-253)."""
import datetime as dt
import math

import numpy as np
import pytest
from scipy import stats
from scipy.special import ive

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import heston_model
from oracle import anchors as A
from oracle import bk_ref as B

pytestmark = pytest.mark.gpu

C2 = dict(kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, V0=0.04, r=0.03)
Q8 = dict(kappa=0.04, theta=0.3, xi=-0.6, rho=0.04, V0=1.5, r=0.05)       # montecarlo_heston.jl:161-170 (Q8)
BK1 = dict(kappa=6.21, theta=0.019, xi=0.61, rho=-0.7, V0=0.010201, r=0.0319)  # Broadie-Kaya (2006) case 1
# low vol of vol = large Bessel order nu = 2 kappa theta / xi^2 - 1 (15, 31.7, 89): Debye's uniform expansion between the
# ascending series and the Hankel expansion
LOWXI = dict(kappa=2.0, theta=0.04, xi=0.1, rho=-0.7, V0=0.04, r=0.03)
LOWXI2 = dict(kappa=2.0, theta=0.04, xi=0.07, rho=-0.5, V0=0.05, r=0.03)
HIGHNU = dict(kappa=5.0, theta=0.09, xi=0.1, rho=-0.3, V0=0.07, r=0.01)


def test_elementary_functions_against_libm(cuda):
    """The table-driven exp / sincos / log / atan2 of csrc/hh_bessel.cuh (they replace the CUDA math library in the
    characteristic function) against numpy in binary64: a few ulp, over the ranges the kernels use and past them."""
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(-690, 690, 20000), rng.uniform(-2, 2, 20000), [0.0, -0.0, 1e-300, 709.0, -745.0, 800.0, -800.0]])
    got = cuda.bk_elementary("exp", x)
    ref = np.exp(x)
    ok = (ref > 1e-300) & np.isfinite(ref)
    assert np.max(np.abs(got[ok] / ref[ok] - 1.0)) < 4e-16
    assert got[-2] == np.inf and got[-1] == 0.0 and np.all(got[ref <= 1e-300] <= 1.1e-300)
    x = np.concatenate([rng.uniform(-1e3, 1e3, 30000), rng.uniform(-7, 7, 20000), rng.uniform(-1e6, 1e6, 5000), [0.0, 1e7, -3e9]])
    s, c = cuda.bk_elementary("sincos", x)
    assert np.max(np.abs(s - np.sin(x))) < 3e-16 and np.max(np.abs(c - np.cos(x))) < 3e-16
    x = np.concatenate([np.exp(rng.uniform(-600, 600, 30000)), rng.uniform(0.5, 2.0, 20000), [1.0, 2.0, 0.5, 1e-310]])
    got = cuda.bk_elementary("log", x)
    ref = np.log(x)
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 3e-16
    y = np.concatenate([rng.normal(size=30000) * np.exp(rng.uniform(-20, 20, 30000)), [0.0, 0.0, 1.0, -1.0, 0.0, -0.0, 1e-320]])
    xx = np.concatenate([rng.normal(size=30000) * np.exp(rng.uniform(-20, 20, 30000)), [1.0, -1.0, 0.0, 0.0, 0.0, -1.0, 1.0]])
    got = cuda.bk_elementary("atan2", xx, y)
    assert np.max(np.abs(got - np.arctan2(y, xx))) < 5e-16


@pytest.mark.parametrize("nu", [-0.933, -0.5, -0.366, 0.0, 0.778, 2.5, 12.0])
def test_log_besseli_matches_amos(cuda, nu):
    rng = np.random.default_rng(3)
    r = np.concatenate([rng.uniform(1e-3, 6, 3000), rng.uniform(4, 40, 3000), rng.uniform(30, 2000, 2000)])
    th = rng.uniform(-np.pi, np.pi, r.size)
    th[::50] = 0.0  # the real axis (z_kappa)
    z = r * np.exp(1j * th)
    z = z[np.abs(np.abs(th) - np.pi / 2) > 0.02]  # I_nu has its zeros on the imaginary axis: ill-conditioned there
    got = cuda.bk_log_besseli(nu, z)
    ref = np.log(ive(nu, z)) + np.abs(z.real)
    assert np.max(np.abs(np.exp(got - ref) - 1.0)) < 2e-12


@pytest.mark.parametrize("nu", [11.6, 12.2, 12.7, 15.0, 31.7, 89.0, 200.0])
def test_log_besseli_large_orders(cuda, nu):
    """Orders from 12 up: |z| between the reach of the tabulated series (~90 + nu/3) and the Hankel expansion
    (20 + nu^2/2) takes Debye's expansion near the real axis and the continued fractions elsewhere. Relative to
    max(1, |log I|): the value itself is only defined to an ulp of its magnitude."""
    rng = np.random.default_rng(7)
    chunks = []
    for lo, hi in [(0.1, 6), (4, 40), (30, 300), (300, 5000), (5000, 50000)]:
        r = rng.uniform(lo, hi, 3000)
        th = rng.uniform(-1.5, 1.5, r.size)
        th[::4] = 0.0          # the real axis (z_kappa)
        th[1::4] *= 0.2        # and its neighbourhood, where the characteristic function is evaluated
        chunks.append(r * np.exp(1j * th))
    z = np.concatenate(chunks)
    with np.errstate(all="ignore"):
        ref = np.log(ive(nu, z).astype(complex)) + np.abs(z.real)
    ok = np.isfinite(ref) & (np.abs(np.abs(np.angle(z)) - np.pi / 2) > 0.05)
    got = cuda.bk_log_besseli(nu, z[ok])
    d = got - ref[ok]
    d = d.real + 1j * ((d.imag + np.pi) % (2 * np.pi) - np.pi)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(d) / np.maximum(1.0, np.abs(ref[ok]))) < 5e-12


@pytest.mark.parametrize("pars,tau", [(C2, 1.0), (C2, 1 / 12), (Q8, 364 / 365), (BK1, 1.0), (BK1, 0.25),
                                      (LOWXI, 1.0), (LOWXI, 0.25), (LOWXI, 1 / 12), (LOWXI2, 0.25), (HIGHNU, 0.25), (HIGHNU, 1 / 12)])
def test_characteristic_function_matches_oracle(cuda, pars, tau):
    m = heston_model(S0=100.0, T=tau, **pars)
    rng = np.random.default_rng(5)
    n, na = 24, 120
    V0 = pars["V0"] * rng.uniform(0.3, 2.0, n)
    VT = pars["V0"] * rng.uniform(0.1, 3.0, n)
    a = np.empty((n, na))
    ref = np.empty((n, na), dtype=complex)
    for i in range(n):
        cf = B.HestonCF(pars["kappa"], pars["theta"], pars["xi"], V0[i], VT[i], tau)
        mean, var = B.moments_from_cf(cf)
        h = math.pi / (mean + 5 * math.sqrt(max(var, 1e-12)))
        a[i] = h * np.arange(1, na + 1)
        th = math.nan
        for j in range(na):
            ref[i, j], th = cf.evaluate(a[i, j], th)
    got = cuda.bk_chf(m, tau, V0, VT, a)
    # |phi| <= 1: absolute == relative to the series' scale. For large arguments (short horizons, low vol of vol) phi is
    # exp of a difference of two log I ~ |z|, each defined to an ulp of |z|: the bound scales with the argument.
    zmax = 4 * pars["kappa"] * np.sqrt(V0 * VT).max() / (pars["xi"] ** 2 * -math.expm1(-pars["kappa"] * tau))
    assert np.max(np.abs(got - ref)) < max(1e-12, 4e-16 * zmax * 8)


@pytest.mark.parametrize("pars,tau", [(C2, 1.0), (C2, 1 / 12), (Q8, 364 / 365), (BK1, 1.0), (LOWXI, 0.25), (HIGHNU, 0.25)])
def test_integral_inversion_against_reference_algorithm(cuda, pars, tau):
    m = heston_model(S0=100.0, T=tau, **pars)
    rng = np.random.default_rng(11)
    n = 96
    d, lam_s, c = B.vt_params(pars["kappa"], pars["theta"], pars["xi"], 1.0, tau)
    V0 = pars["V0"] * rng.uniform(0.5, 1.5, n)
    VT = np.array([c * stats.ncx2.rvs(d, lam_s * v, random_state=rng) for v in V0])
    U = rng.uniform(0.001, 0.999, n)
    U[:4] = [1e-9, 0.5, 0.999999, 0.9999999999]
    g = cuda.bk_integral(m, tau, V0, VT, U)
    for i in range(n):
        cf = B.HestonCF(pars["kappa"], pars["theta"], pars["xi"], V0[i], VT[i], tau)
        o = B.sample_integral_V(cf, U[i])
        # moments_from_cf differences Phi at +-1e-2: the SECOND difference cancels to ~1e-16 / h0^2 = 1e-12 absolute per
        # rounding, so the reference's own variance (and the h derived from it) is only defined to about that noise
        # (and, phi being exp of a difference of log I ~ z_kappa, the noise grows with the argument: low vol of vol)
        zk = 4 * pars["kappa"] * math.sqrt(V0[i] * VT[i]) * math.exp(-0.5 * pars["kappa"] * tau) / (pars["xi"] ** 2 * -math.expm1(-pars["kappa"] * tau))
        assert g["mean"][i] == pytest.approx(o["mean"], rel=1e-7, abs=1e-14 * zk)
        assert g["var"][i] == pytest.approx(o["var"], rel=1e-6, abs=max(5e-10, 5e-11 * zk))   # ~ eps |log I| / h0^2, a few roundings, on both sides
        assert g["h"][i] == pytest.approx(o["h"], rel=5e-3)
        assert abs(int(g["J"][i]) - o["J"]) <= 1
        # with IDENTICAL h the series (truncation rule included) and its CDF must agree tightly
        hg = g["h"][i]
        phis = B.cf_series(cf, hg)
        assert int(g["J"][i]) == len(phis)
        F_same_h = lambda x: B.cdf_from_series(phis, x, hg)
        if g["status"][i] == 2:  # fell back to max_guess (u ~ 1: the truncated series never reaches u), like the reference
            assert F_same_h(g["x"][i]) < U[i]
            assert g["x"][i] == pytest.approx(o["max_guess"], rel=1e-3)
            continue
        assert g["x"][i] >= 0.0
        assert abs(F_same_h(g["x"][i]) - U[i]) <= 1e-10
        assert abs(g["resid"][i]) <= 1e-12
        # the GPU root passes the reference's acceptance test (sample_from_cf.jl:119) on the oracle's own CDF (own h)
        assert abs(o["cdf"](g["x"][i]) - U[i]) <= 1e-4

@pytest.mark.parametrize("pars,tau,v0", [(C2, 1.0, 0.04), (C2, 1 / 252, 0.09), (Q8, 364 / 365, 1.5), (BK1, 1.0, 0.010201),
                                         (dict(C2, xi=0.05), 1 / 12, 0.04)])
def test_variance_sampler_is_noncentral_chisquare(cuda, pars, tau, v0):
    """d > 1 (chi2 + shifted normal), d < 1 (Poisson mixture, small and large means) against scipy's ncx2 law."""
    m = heston_model(S0=100.0, T=tau, **pars)
    n = 200_000
    VT = cuda.bk_variance(m, tau, np.full(n, v0), seed=123)
    d, lam, c = B.vt_params(pars["kappa"], pars["theta"], pars["xi"], v0, tau)
    assert np.all(VT >= 0) and np.all(np.isfinite(VT))
    ks = stats.kstest(VT / c, stats.ncx2(d, lam).cdf)
    assert ks.pvalue > 1e-3, ks
    assert VT.mean() == pytest.approx(c * (d + lam), rel=5 * math.sqrt(2 * (d + 2 * lam)) / (d + lam) / math.sqrt(n) + 1e-9)


def _bk_problem(pars, T_days, n, steps=1, strike=100.0, seed=2024):
    ref = dt.date(2020, 1, 1)
    prob = hh.PricingProblem(hh.VanillaOption(strike, ref + dt.timedelta(days=T_days), hh.European(), hh.Call(), hh.Spot()),
                             hh.HestonInputs(ref, pars["r"], 100.0, pars["V0"], pars["kappa"], pars["theta"], pars["xi"], pars["rho"]))
    mc = hh.MonteCarlo(hh.HestonDynamics(), hh.HestonBroadieKaya(), hh.SimulationConfig(n, steps=steps, base_seed=seed),
                       bk_steps_from_config=steps > 1)
    return prob, mc


@pytest.mark.parametrize("pars,days,tol", [(C2, 365, 2e-2), (Q8, 364, 5e-2), (BK1, 365, 2e-2), (LOWXI, 91, 2e-2), (LOWXI2, 30, 2e-2),
                                           (HIGHNU, 91, 2e-2)])
def test_bk_price_vs_carr_madan(cuda, pars, days, tol):
    """montecarlo_heston.jl:151-253: BK exact (NoVarianceReduction) vs Carr-Madan(1, 32), rtol 2e-2 / 5e-2 — and, being
    an exact scheme, within 3 standard errors (+ the 1e-4 Carr-Madan truncation)."""
    n = 400_000
    prob, mc = _bk_problem(pars, days, n)
    sol = hh.solve(prob, mc, engine=cuda)
    T = days / 365
    cm = A.heston_price(100.0, 100.0, pars["r"], T, pars["V0"], pars["kappa"], pars["theta"], pars["xi"], pars["rho"], bound=200.0)
    assert sol.price == pytest.approx(cm, rel=tol)
    assert abs(sol.price - cm) < 3.5 * sol.std_error + 2e-4 * cm, (sol.price, cm, sol.std_error)
    assert sol.stats["n_nonfinite"] == 0
    assert sol.stats["n_fallback"] <= 1e-4 * n
    assert len(sol.ensemble) == n and np.all(sol.ensemble > 0)


def test_bk_multi_date_is_consistent_with_one_transition(cuda):
    """Config C4: 12 exact transitions compose to the same law as one (heston.jl:82-91 restarts from (S, V) each date)."""
    n = 300_000
    p1, m1 = _bk_problem(C2, 365, n, steps=1, seed=5)
    p12, m12 = _bk_problem(C2, 365, n, steps=12, seed=6)
    s1 = hh.solve(p1, m1, engine=cuda)
    s12 = hh.solve(p12, m12, engine=cuda)
    cm = A.heston_price(100.0, 100.0, 0.03, 1.0, 0.04, 2.0, 0.04, 0.3, -0.7, bound=200.0)
    for s in (s1, s12):
        assert abs(s.price - cm) < 3.5 * s.std_error + 2e-4 * cm, (s.price, cm, s.std_error)
    assert stats.ks_2samp(s1.ensemble[:100_000], s12.ensemble[:100_000]).pvalue > 1e-3
    st = cuda.bk_last_stats()
    assert st["transitions"] == 12 * n and 3 < st["mean_series_terms"] < 500


def test_bk_errors_and_shards(cuda):
    m = heston_model()
    with pytest.raises(NotImplementedError):  # Q5: Antithetic + BK is a MethodError in the reference
        cuda.mc_european(m, SimSpec(n_paths=10, n_steps=1, scheme=abi.HH_SCHEME_HESTON_BK, vr=abi.HH_VR_ANTITHETIC), [(100.0, 1.0)], 1.0)
    # shard invariance: the stream is keyed by the global trajectory index
    sim = SimSpec(n_paths=4000, n_steps=2, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=9)
    _, full = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    parts = []
    for g in range(4):
        _, t = cuda.mc_european(m, SimSpec(n_paths=1000, path_offset=1000 * g, n_steps=2, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=9),
                                [(100.0, 1.0)], 1.0, want_terminal=True)
        parts.append(t)
    assert np.array_equal(np.concatenate(parts), full)


def test_bk_chunked_launch_is_bit_identical(cuda):
    """Jobs above 2^27 transitions are cut into chunks of trajectories (hh_bk.cu: bk_paths_launch). HH_BK_MAX_ITEMS forces
    small chunks in a second process; the terminal spots, the payoff sums and the inversion counters must not move."""
    import json
    import os
    import subprocess
    import sys
    code = (
        "import sys, json, hashlib, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import hedgehog_jl_b200 as hh\n"
        "from hedgehog_jl_b200 import _abi as abi\n"
        "from hedgehog_jl_b200.engine import SimSpec\n"
        "from helpers import heston_model\n"
        "eng = hh.default_engine(0)\n"
        "sim = SimSpec(n_paths=5000, n_steps=3, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=9)\n"
        "r, t = eng.mc_european(heston_model(), sim, [(100.0, 1.0), (90.0, -1.0)], 1.0, want_terminal=True)\n"
        "print(json.dumps({'sha': hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest(), 'sum': [x.sum for x in r],\n"
        "                  'stats': eng.bk_last_stats()}))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for extra in ({}, {"HH_BK_MAX_ITEMS": "3500"}):   # 3500 // 3 = 1166 trajectories per chunk: five chunks, the last one ragged
        env = dict(os.environ, **extra)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(json.loads(p.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1]
    assert outs[0]["stats"]["transitions"] == 15000


def _mp_cumulants(kappa, theta, xi, V0, VT, tau):
    """Mean and variance of int V dt given (V0, VT) from the characteristic function of heston.jl:184-212 in 50-digit
    arithmetic (derivatives of log Phi at 0): the ground truth the finite-difference moments approximate."""
    import mpmath as mp
    mp.mp.dps = 50
    k, s2, nu = mp.mpf(kappa), mp.mpf(xi) ** 2, 2 * mp.mpf(kappa) * theta / mp.mpf(xi) ** 2 - 1
    sv, vs, t = mp.sqrt(mp.mpf(V0) * VT), mp.mpf(V0) + VT, mp.mpf(tau)

    def parts(g):
        e = mp.exp(-g * t)
        return (1 - e) / g, g * (1 + e) / (1 - e), sv * 4 * g * mp.exp(-g * t / 2) / (s2 * (1 - e))

    zk, ek, wk = parts(k)

    def logphi(a):
        g = mp.sqrt(k * k - 2 * s2 * a * 1j)
        zg, eg, wg = parts(g)
        return -(g - k) * t / 2 + mp.log(zk / zg) + vs / s2 * (ek - eg) + mp.log(mp.besseli(nu, wg) / mp.besseli(nu, wk))

    d1 = mp.diff(logphi, 0, 1, h=mp.mpf(10) ** -12)
    d2 = mp.diff(logphi, 0, 2, h=mp.mpf(10) ** -12)
    return float(mp.im(d1)), float(-mp.re(d2))


@pytest.mark.parametrize("pars,tau", [(LOWXI, 1 / 52), (C2, 1 / 252), (C2, 1 / 52), (HIGHNU, 1 / 52), (C2, 1 / 12)])
def test_moments_where_the_finite_difference_is_noise(cuda, pars, tau):
    """sigma tau < ~0.01: the rounding of Phi (exponential of a difference of terms ~8 V / (sigma^2 tau)) divided by h0^2 = 1e-4
    exceeds the variance of the integral, and moments_from_cf (sample_from_cf.jl:50-64) returns noise — negative half of
    the time. The kernel then re-reads the variance from |Phi| at a step scaled to the law (hh_bk.cu: bk_sample_integral);
    checked against 50-digit cumulants. With h_fd < 0 (the reference's plain differences) the same call is off by the
    noise; at C4's monthly horizon the two modes are the same numbers to the last bit."""
    m = heston_model(S0=100.0, T=tau, **pars)
    rng = np.random.default_rng(21)
    n = 6
    d, lam_s, c = B.vt_params(pars["kappa"], pars["theta"], pars["xi"], 1.0, tau)
    V0 = pars["V0"] * rng.uniform(0.5, 1.5, n)
    VT = np.array([c * stats.ncx2.rvs(d, lam_s * v, random_state=rng) for v in V0])
    U = rng.uniform(0.05, 0.95, n)
    g = cuda.bk_integral(m, tau, V0, VT, U)
    plain = abi.hh_bk_config()
    cuda.lib.hh_default_bk_config(plain)
    plain.h_fd = -plain.h_fd
    gp = cuda.bk_integral(m, tau, V0, VT, U, cfg=plain)
    truth = np.array([_mp_cumulants(pars["kappa"], pars["theta"], pars["xi"], V0[i], VT[i], tau) for i in range(n)])
    assert np.max(np.abs(g["mean"] / truth[:, 0] - 1.0)) < 1e-5
    if tau >= 1 / 12:
        assert np.array_equal(g["var"], gp["var"]) and np.array_equal(g["x"], gp["x"])
        assert np.max(np.abs(g["var"] / truth[:, 1] - 1.0)) < 2e-3      # the finite difference's own truncation + noise
        return
    assert np.max(np.abs(g["var"] / truth[:, 1] - 1.0)) < 2e-2
    assert np.max(np.abs(gp["var"] / truth[:, 1] - 1.0)) > 0.05         # what the plain differences give there
    # and the sample is a root of the CDF on a grid that covers the law
    assert np.all(g["status"] == 0) and np.all(np.abs(g["x"] - truth[:, 0]) < 8 * np.sqrt(truth[:, 1]))


@pytest.mark.parametrize("pars,dates,xi_note", [(LOWXI, 52, "sigma tau = 0.002"), (C2, 252, "sigma tau = 0.0012")])
def test_bk_price_on_many_dates(cuda, pars, dates, xi_note):
    """Weekly and daily exact transitions (path-dependent payoffs and American exercise on Broadie-Kaya dates use them):
    within 3.5 standard errors of Carr-Madan. Before the variance was re-read in the noise-dominated regime the first case
    was 25 standard errors high."""
    n = 200_000
    prob, mc = _bk_problem(pars, 365, n, steps=dates)
    sol = hh.solve(prob, mc, engine=cuda)
    cm = A.heston_price(100.0, 100.0, pars["r"], 1.0, pars["V0"], pars["kappa"], pars["theta"], pars["xi"], pars["rho"], bound=200.0)
    assert abs(sol.price - cm) < 3.5 * sol.std_error + 2e-4 * cm, (sol.price, cm, sol.std_error)
    assert sol.stats["n_nonfinite"] == 0 and sol.stats["n_fallback"] <= 1e-4 * n * dates
