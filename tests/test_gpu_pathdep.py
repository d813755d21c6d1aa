"""GPU parity tests for the path-dependent payoffs (hh_mc_path_dependent) through the C ABI against the
CPU restatement (oracle/hh_oracle.c hho_mc_path_dependent, itself pinned by closed forms and numpy in
tests/test_pathdep_oracle.py).

Bars: per-column statistics rel 1e-11 (the kernel accumulates in log space, the oracle in S-space from saved spots);
sums rel 1e-10 for the continuous payoffs; for barriers and digitals a column whose statistic sits within rounding of the
barrier/strike may fall on the other side, so those are compared after excluding such ties (none at these sizes, which
the test asserts)."""
import datetime as dt
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err
from test_pathdep_oracle import ALL_KINDS, numpy_payoff

pytestmark = pytest.mark.gpu


HESTON_KINDS = ALL_KINDS + [(abi.HH_PD_BS_CONTROL, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_VANILLA_MINUS_BS, 95.0, -1.0, 0.0, 0.8)]


def _compare(cuda, oracle, m, sim, payoffs, every, stat_tol=1e-11):
    rg, sg = cuda.mc_path_dependent(m, sim, payoffs, 0.97, every, want_stats=True)
    ro, so = oracle.mc_path_dependent(m, sim, payoffs, 0.97, every, want_stats=True)
    assert rel_err(sg, so) < stat_tol
    n = sim.n_paths
    for c, g, o in zip(payoffs, rg, ro):
        if c[0] < abi.HH_PD_BS_CONTROL:  # (the control's terminal spot is not among the exported statistics)
            pg, po = numpy_payoff(c, sg), numpy_payoff(c, so)
            assert np.count_nonzero(np.abs(pg - po) > 1e-9 * (1.0 + np.abs(po))) == 0  # no decision flipped
        assert g.n == o.n == n
        assert g.sum == pytest.approx(o.sum, rel=1e-10, abs=1e-9)
        assert g.sumsq == pytest.approx(o.sumsq, rel=1e-10, abs=1e-9)
        assert g.price == pytest.approx(o.price, rel=1e-10, abs=1e-12)
        assert g.std_error == pytest.approx(o.std_error, rel=1e-8, abs=1e-12)
        assert g.n_nonfinite == 0
    # requesting the statistics must not change the sums
    rg2, none = cuda.mc_path_dependent(m, sim, payoffs, 0.97, every)
    assert none is None
    for a, b in zip(rg, rg2):
        assert a.sum == pytest.approx(b.sum, rel=1e-13)
    return rg, sg


@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("model", ["gbm", "gbm_steps", "heston", "heston_nosplit"])
def test_parity_mode_all_payoffs(cuda, oracle, model, anti):
    n, M = 5003, 24
    heston = model.startswith("heston")
    m = heston_model(xi=0.6, split=model == "heston") if heston else gbm_model()
    z = np.random.default_rng(11).standard_normal((n, M, 2) if heston else (n, M))
    scheme = abi.HH_SCHEME_EXACT_STEPS if model == "gbm_steps" else abi.HH_SCHEME_EM
    sim = SimSpec(n_paths=n, n_steps=M, scheme=scheme, vr=int(anti), rng_mode=abi.HH_RNG_NORMALS, normals=z)
    for every in (1, 6):
        _compare(cuda, oracle, m, sim, HESTON_KINDS if heston else ALL_KINDS, every)


@pytest.mark.parametrize("model", ["gbm", "heston"])
@pytest.mark.parametrize("seeded", [False, True])
def test_native_rng_all_payoffs(cuda, oracle, model, seeded):
    n, M = 40_001, 50
    m = heston_model() if model == "heston" else gbm_model(sigma=0.3)
    kw = dict(seeds=np.random.Generator(np.random.Philox(2)).integers(0, 2**64, size=n, dtype=np.uint64)) if seeded \
        else dict(base_seed=4242, path_offset=(1 << 33) + 5)
    sim = SimSpec(n_paths=n, n_steps=M, scheme=abi.HH_SCHEME_EM, vr=abi.HH_VR_ANTITHETIC, **kw)
    _compare(cuda, oracle, m, sim, HESTON_KINDS if model == "heston" else ALL_KINDS, 5, stat_tol=1e-10)
    if model == "heston":   # and without the antithetic side (the 1024-thread instantiation when the key is uniform)
        _compare(cuda, oracle, m, SimSpec(n_paths=200_003 if not seeded else n, n_steps=20, scheme=abi.HH_SCHEME_EM, **(kw if seeded else dict(base_seed=6))),
                 HESTON_KINDS, 4, stat_tol=1e-10)


def test_without_arithmetic_average_the_kernel_stays_in_log_space(cuda, oracle):
    """No arithmetic Asian requested and no statistics: the ARITH = false instantiation; same sums as the oracle."""
    m = heston_model()
    sim = SimSpec(n_paths=20_000, n_steps=30, scheme=abi.HH_SCHEME_EM, base_seed=9)
    pays = [c for c in ALL_KINDS if c[0] != abi.HH_PD_ASIAN_ARITH]
    rg, _ = cuda.mc_path_dependent(m, sim, pays, 1.0, 3)
    ro, _ = oracle.mc_path_dependent(m, sim, pays, 1.0, 3)
    for g, o in zip(rg, ro):
        assert g.sum == pytest.approx(o.sum, rel=1e-10)


def test_many_contracts_and_odd_shapes(cuda, oracle):
    m = gbm_model()
    rng = np.random.default_rng(3)
    pays = [(int(rng.integers(0, abi.HH_PD_BS_CONTROL)), float(rng.uniform(80, 120)), float(rng.choice([-1.0, 1.0])),
             float(rng.uniform(70, 140)), float(rng.uniform(0, 2))) for _ in range(256)]
    for n, M, every in ((1, 1, 1), (513, 7, 7), (1025, 9, 3)):
        sim = SimSpec(n_paths=n, n_steps=M, scheme=abi.HH_SCHEME_EM, base_seed=n)
        rg, _ = cuda.mc_path_dependent(m, sim, pays, 1.0, every)
        ro, _ = oracle.mc_path_dependent(m, sim, pays, 1.0, every)
        for g, o in zip(rg, ro):
            assert g.sum == pytest.approx(o.sum, rel=1e-10, abs=1e-9)


def test_argument_errors(cuda):
    m = gbm_model()
    sim = SimSpec(n_paths=100, n_steps=8, scheme=abi.HH_SCHEME_EM)
    with pytest.raises(ValueError):
        cuda.mc_path_dependent(m, sim, ALL_KINDS, 1.0, 3)                      # 8 % 3 != 0
    with pytest.raises(ValueError):
        cuda.mc_path_dependent(m, sim, [(99, 100.0, 1.0, 0.0, 0.0)], 1.0, 1)   # unknown kind
    with pytest.raises(ValueError):
        cuda.mc_path_dependent(m, sim, [(abi.HH_PD_UP_OUT, 100.0, 1.0, -5.0, 0.0)], 1.0, 1)  # barrier <= 0
    with pytest.raises(ValueError):
        cuda.mc_path_dependent(m, sim, [(abi.HH_PD_VANILLA, 100.0, 0.5, 0.0, 0.0)], 1.0, 1)  # cp not +-1
    with pytest.raises(ValueError):
        cuda.mc_path_dependent(m, sim, [ALL_KINDS[0]] * 257, 1.0, 1)
    with pytest.raises(NotImplementedError):
        cuda.mc_path_dependent(m, SimSpec(n_paths=100, n_steps=8, scheme=abi.HH_SCHEME_EXACT_TERMINAL), ALL_KINDS, 1.0, 1)
    with pytest.raises(ValueError):            # HestonBroadieKaya is not defined for LognormalDynamics
        cuda.mc_path_dependent(m, SimSpec(n_paths=100, n_steps=8, scheme=abi.HH_SCHEME_HESTON_BK), ALL_KINDS, 1.0, 1)
    with pytest.raises(NotImplementedError):
        cuda.mc_path_dependent(heston_model(), SimSpec(n_paths=100, n_steps=8, scheme=abi.HH_SCHEME_EM, precision=abi.HH_PREC_F32),
                               ALL_KINDS, 1.0, 1)


def test_closed_forms_through_solve(cuda):
    """solve(PricingProblem(AsianOption | BarrierOption | DigitalOption, BlackScholesInputs), MonteCarlo) at 4e6 antithetic
    pairs: 3.5 sigma of the closed forms; knock-in + knock-out = vanilla on common trajectories."""
    from oracle import anchors as A
    ref, exp = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
    T = 366 / 365
    mk = hh.BlackScholesInputs(ref, 0.05, 100.0, 0.2)
    cfg = hh.SimulationConfig(4_000_000, steps=100, base_seed=3, variance_reduction=hh.Antithetic())
    mc = hh.MonteCarlo(hh.LognormalDynamics(), hh.EulerMaruyama(), cfg)
    mon = hh.Monitoring(2)   # 50 monitoring dates on 100 steps
    ps = [hh.AsianOption(100.0, exp, hh.Call(), hh.GeometricAverage(), mon), hh.DigitalOption(100.0, exp, hh.Put(), hh.CashOrNothing(2.0), mon),
          hh.DigitalOption(100.0, exp, hh.Call(), hh.AssetOrNothing(), mon),
          hh.BarrierOption(100.0, 125.0, exp, hh.Call(), hh.Up(), hh.KnockIn(), monitoring=mon),
          hh.BarrierOption(100.0, 125.0, exp, hh.Call(), hh.Up(), hh.KnockOut(), monitoring=mon),
          hh.VanillaOption(100.0, exp, hh.European(), hh.Call(), hh.Spot())]
    sols = hh.solve(hh.BasketPricingProblem(ps, mk), mc, engine=cuda)
    ui = A.up_and_in_call_price(100.0, 100.0, A.discrete_barrier_shift(125.0, 0.2, T, 50), 0.05, 0.2, T)
    bs = A.bs_price(100.0, 100.0, 0.05, 0.2, T)
    exact = [A.geometric_asian_price(100.0, 100.0, 0.05, 0.2, T, 50), A.digital_price(100.0, 100.0, 0.05, 0.2, T, -1.0, 2.0),
             A.digital_price(100.0, 100.0, 0.05, 0.2, T, 1.0), ui, bs - ui, bs]
    slack = [0, 0, 0, 1e-2 * ui, 1e-2 * ui, 0]   # the continuity correction is itself an approximation (~0.5 % of the knock-in)
    for s, e, sl in zip(sols, exact, slack):
        assert abs(s.price - e) < 3.5 * s.std_error + sl, (s.price, e, s.std_error)
    assert sols[3].price + sols[4].price == pytest.approx(sols[5].price, rel=1e-12)


def test_heston_asian_put_call_parity_of_the_average(cuda):
    """E[A] under Heston is model-free: (S0/m) sum_i e^{r t_i}; so Asian call - put = D (E[A] - K) within 3.5 sigma."""
    ref, exp = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
    T = 366 / 365
    mk = hh.HestonInputs(ref, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    cfg = hh.SimulationConfig(2_000_000, steps=120, base_seed=8, variance_reduction=hh.Antithetic())
    mc = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), cfg)
    mon = hh.Monitoring(10)
    call, put = hh.solve(hh.BasketPricingProblem([hh.AsianOption(100.0, exp, hh.Call(), monitoring=mon),
                                                  hh.AsianOption(100.0, exp, hh.Put(), monitoring=mon)], mk), mc, engine=cuda)
    EA = 100.0 * np.mean([math.exp(0.03 * T * i / 12) for i in range(1, 13)])
    # Euler-Maruyama in log space carries a small martingale bias (E[S] is off by O(dt)); allow 2e-4 relative on E[A]
    assert abs((call.price - put.price) - math.exp(-0.03 * T) * (EA - 100.0)) < 3.5 * math.hypot(call.std_error, put.std_error) + 2e-2


# ---- exact Broadie-Kaya transitions between the monitoring dates -------------------------------------------------------

def test_broadie_kaya_statistics_match_the_european_path(cuda):
    """The statistics come out of the same path kernel as hh_mc_european under HestonBroadieKaya: S_T is bit-identical, a
    VANILLA contract reproduces the European price, knock-in + knock-out = vanilla, min <= G <= A <= max."""
    m = heston_model()
    n, dates = 20_000, 6
    sim = SimSpec(n_paths=n, n_steps=dates, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=31)
    pays = [(abi.HH_PD_VANILLA, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_UP_OUT, 100.0, 1.0, 125.0, 0.0), (abi.HH_PD_UP_IN, 100.0, 1.0, 125.0, 0.0),
            (abi.HH_PD_ASIAN_ARITH, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_ASIAN_GEOM, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_DOWN_IN, 100.0, -1.0, 85.0, 0.0)]
    res, st = cuda.mc_path_dependent(m, sim, pays, 0.97, 1, want_stats=True)
    eur, term = cuda.mc_european(m, sim, [(100.0, 1.0)], 0.97, want_terminal=True)
    assert np.array_equal(st[0], term)
    assert res[0].sum == pytest.approx(eur[0].sum, rel=1e-13)
    assert res[1].sum + res[2].sum == pytest.approx(res[0].sum, rel=1e-12)
    assert np.all(st[4] <= st[2] * (1 + 1e-14)) and np.all(st[2] <= st[1] * (1 + 1e-14)) and np.all(st[1] <= st[3] * (1 + 1e-14))
    for c, r in zip(pays, res):
        assert r.sum == pytest.approx(numpy_payoff(c, st).sum(), rel=1e-12)
        assert r.n_fallback == 0
    # monitoring every second date: the statistics of dates 2, 4, 6 only
    res2, st2 = cuda.mc_path_dependent(m, sim, pays, 0.97, 2, want_stats=True)
    assert np.array_equal(st2[0], st[0]) and np.all(st2[3] <= st[3]) and np.all(st2[4] >= st[4])
    with pytest.raises(NotImplementedError):   # Q5: Antithetic + HestonBroadieKaya
        cuda.mc_path_dependent(m, SimSpec(n_paths=n, n_steps=dates, scheme=abi.HH_SCHEME_HESTON_BK, vr=abi.HH_VR_ANTITHETIC), pays, 1.0, 1)


def test_broadie_kaya_asian_agrees_with_fine_euler_maruyama(cuda):
    """Monthly-monitored Asian and barrier options: exact transitions (12 dates) against Euler-Maruyama on 40 steps per
    month, through solve(); agreement within 4 combined standard errors plus the scheme's O(dt) bias."""
    ref, exp = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
    mk = hh.HestonInputs(ref, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    def basket(mon):
        return hh.BasketPricingProblem([hh.AsianOption(100.0, exp, hh.Call(), monitoring=mon),
                                        hh.AsianOption(100.0, exp, hh.Put(), hh.GeometricAverage(), mon),
                                        hh.BarrierOption(100.0, 120.0, exp, hh.Call(), hh.Up(), hh.KnockOut(), monitoring=mon)], mk)
    bk = hh.solve(basket(hh.Monitoring(1)), hh.MonteCarlo(hh.HestonDynamics(), hh.HestonBroadieKaya(),
                                                         hh.SimulationConfig(400_000, steps=12, base_seed=5), ensemble=False), engine=cuda)
    em = hh.solve(basket(hh.Monitoring(40)), hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(),
                                                          hh.SimulationConfig(2_000_000, steps=480, base_seed=6), ensemble=False), engine=cuda)
    for a, b in zip(bk, em):
        assert abs(a.price - b.price) < 4.0 * math.hypot(a.std_error, b.std_error) + 5e-3 * b.price, (a.price, b.price, a.std_error, b.std_error)


def test_results_are_reproducible_and_shard_additive(cuda):
    """Fixed-order reductions: the same call twice gives bit-identical sums; two shards of the global trajectory index
    (path_offset) add up to the whole job's sums to rounding, statistics concatenate exactly."""
    m = heston_model()
    pays = [c for c in ALL_KINDS]
    whole = SimSpec(n_paths=30_000, n_steps=20, scheme=abi.HH_SCHEME_EM, base_seed=77)
    r1, s1 = cuda.mc_path_dependent(m, whole, pays, 1.0, 4, want_stats=True)
    r2, s2 = cuda.mc_path_dependent(m, whole, pays, 1.0, 4, want_stats=True)
    assert all(a.sum == b.sum and a.sumsq == b.sumsq for a, b in zip(r1, r2)) and np.array_equal(s1, s2)
    lo = SimSpec(n_paths=12_345, n_steps=20, scheme=abi.HH_SCHEME_EM, base_seed=77, path_offset=0)
    hi = SimSpec(n_paths=30_000 - 12_345, n_steps=20, scheme=abi.HH_SCHEME_EM, base_seed=77, path_offset=12_345)
    ra, sa = cuda.mc_path_dependent(m, lo, pays, 1.0, 4, want_stats=True)
    rb, sb = cuda.mc_path_dependent(m, hi, pays, 1.0, 4, want_stats=True)
    assert np.array_equal(np.concatenate([sa, sb], axis=1), s1)
    for a, b, w in zip(ra, rb, r1):
        assert a.sum + b.sum == pytest.approx(w.sum, rel=1e-13)


def test_black_scholes_control_variate_through_solve(cuda):
    """MonteCarlo(HestonDynamics(), EulerMaruyama(), config, control_variate=BlackScholesControlVariate()): same
    expectation as the plain estimator, standard error more than halved (variance / 6 at these parameters), agreement
    with Carr-Madan up to the Euler bias; the control variate is rejected where its closed form does not apply."""
    from oracle import anchors as A
    ref, exp = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
    mk = hh.HestonInputs(ref, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    prob = hh.PricingProblem(hh.VanillaOption(100.0, exp, hh.European(), hh.Call(), hh.Spot()), mk)
    cfg = hh.SimulationConfig(4_000_000, steps=252, base_seed=9)
    plain = hh.solve(prob, hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), cfg, ensemble=False), engine=cuda)
    cv = hh.solve(prob, hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), cfg, ensemble=False,
                                      control_variate=hh.BlackScholesControlVariate()), engine=cuda)
    assert cv.std_error < 0.5 * plain.std_error and 0.5 < cv.stats["beta"] < 1.0
    assert abs(cv.price - plain.price) < 3.5 * plain.std_error
    cm = A.heston_price(100.0, 100.0, 0.03, 366 / 365, 0.04, 2.0, 0.04, 0.3, -0.7)
    assert abs(cv.price - cm) < 3.5 * cv.std_error + 0.008   # Euler bias at 252 steps: +0.0057
    with pytest.raises(TypeError):
        hh.solve(hh.PricingProblem(prob.payoff, hh.BlackScholesInputs(ref, 0.03, 100.0, 0.2)),
                 hh.MonteCarlo(hh.LognormalDynamics(), hh.EulerMaruyama(), cfg, control_variate=hh.BlackScholesControlVariate()), engine=cuda)
    with pytest.raises(ValueError):   # C ABI: the control kinds need HestonDynamics + EulerMaruyama
        cuda.mc_path_dependent(gbm_model(), SimSpec(n_paths=10, n_steps=5, scheme=abi.HH_SCHEME_EM),
                               [(abi.HH_PD_VANILLA_MINUS_BS, 100.0, 1.0, 0.0, 1.0)], 1.0, 1)
