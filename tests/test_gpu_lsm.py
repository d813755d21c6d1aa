"""GPU parity tests for Longstaff-Schwartz (solve(::PricingProblem{American}, ::LSM), least_squares_montecarlo.jl:99-136)
through the C ABI, against the CPU oracle (Householder QR on the raw Vandermonde, like Polynomials.fit) and CRR.

Bars: stored spot paths rel 1e-12 (same normals / same Philox stream); exercise decisions identical except where
payoff and fitted continuation tie within rounding (the oracle's QR on raw monomials of S ~ 100 is itself only good to
~1e-9 there), so flips are counted and bounded, and the price must agree to 1e-9 relative when no decision flipped."""
import datetime as dt
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err

pytestmark = pytest.mark.gpu


def _run_both(cuda, oracle, m, sim, payoff, degree, D):
    og, tg, vg, pg = cuda.lsm_american(m, sim, payoff, degree, D, want_stopping=True, want_paths=True)
    oo, to, vo, po = oracle.lsm_american(m, sim, payoff, degree, D, want_stopping=True, want_paths=True)
    return (og, tg, vg, pg), (oo, to, vo, po)


def _check(g, o, max_flip_frac=2e-4, path_tol=1e-12):
    (og, tg, vg, pg), (oo, to, vo, po) = g, o
    assert rel_err(pg, po) < path_tol
    assert og.n == oo.n
    flips = int(np.sum(tg != to))
    assert flips <= max(2, max_flip_frac * len(to)), flips
    same = tg == to
    assert rel_err(vg[same] + 1.0, vo[same] + 1.0) < path_tol
    if flips == 0:
        assert abs(og.price - oo.price) <= 1e-9 * abs(oo.price)
        assert abs(og.std_error - oo.std_error) <= 1e-7 * abs(oo.std_error)
    else:  # a flipped decision moves one column's cash flow by the (tiny) tie gap
        assert abs(og.price - oo.price) <= 1e-6 * abs(oo.price)
    assert og.n_dates_skipped == oo.n_dates_skipped
    return flips


@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("degree", [2, 3, 5])
def test_lsm_parity_mode(cuda, oracle, anti, degree):
    n, steps = 4097, 20  # odd column count when not antithetic
    m = gbm_model(S0=100.0, r=0.05, sigma=0.2, T=1.0)
    z = np.random.Generator(np.random.Philox(21)).standard_normal((n, steps, 1))
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=int(anti), rng_mode=abi.HH_RNG_NORMALS, normals=z)
    D = math.exp(-m.r * m.T / steps)
    _check(*_run_both(cuda, oracle, m, sim, (100.0, -1.0), degree, D))


@pytest.mark.parametrize("cp,strike", [(-1.0, 100.0), (-1.0, 110.0), (1.0, 100.0)])
def test_lsm_native_rng_matches_oracle(cuda, oracle, cp, strike):
    n, steps = 50_000, 50
    m = gbm_model(S0=100.0 if cp < 0 else 120.0, r=0.05 if cp < 0 else 0.15, sigma=0.25, T=0.5)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=abi.HH_VR_ANTITHETIC, base_seed=12345)
    D = math.exp(-m.r * m.T / steps)
    _check(*_run_both(cuda, oracle, m, sim, (strike, cp), 4, D))


@pytest.mark.parametrize("anti,steps,seeded", [(False, 50, False), (True, 49, False), (False, 7, True), (True, 2, False)])
def test_lsm_philox64_stream_matches_oracle(cuda, oracle, anti, steps, seeded):
    """HH_RNG_PHILOX_64 in the exact GBM generator: ONE Philox block per FOUR steps (32-bit radius + 32-bit angle per
    Box-Muller pair), restated in the oracle as hho_normal_pair64(key, idx, n >> 1) component n & 1. Step counts that are
    not multiples of four, antithetic columns, per-path keys; stored paths 1e-12, decisions equal up to counted ties."""
    n = 30_001
    m = gbm_model(S0=100.0, r=0.05, sigma=0.25, T=0.5)
    kw = dict(seeds=np.random.Generator(np.random.Philox(5)).integers(0, 2**64, size=n, dtype=np.uint64)) if seeded else dict(base_seed=777)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=int(anti), rng_mode=abi.HH_RNG_PHILOX_64, **kw)
    g, o = _run_both(cuda, oracle, m, sim, (100.0, -1.0), 3, math.exp(-m.r * m.T / steps))
    _check(g, o)
    # a different stream from HH_RNG_PHILOX (counter stream word 2): the same law, other numbers
    sim0 = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=int(anti), **kw)
    o0 = cuda.lsm_american(m, sim0, (100.0, -1.0), 3, math.exp(-m.r * m.T / steps))[0]
    assert g[0].price != o0.price and abs(g[0].price - o0.price) < 5 * (g[0].std_error + o0.std_error)


def test_lsm_philox64_is_rejected_where_it_is_not_defined(cuda):
    with pytest.raises(NotImplementedError):   # European pricing under LognormalDynamics
        cuda.mc_european(gbm_model(), SimSpec(n_paths=100, n_steps=4, scheme=abi.HH_SCHEME_EXACT_STEPS, rng_mode=abi.HH_RNG_PHILOX_64),
                         [(100.0, 1.0)], 1.0)
    with pytest.raises(NotImplementedError):   # LSM with the log-space generators
        cuda.lsm_american(gbm_model(), SimSpec(n_paths=100, n_steps=4, scheme=abi.HH_SCHEME_EM, rng_mode=abi.HH_RNG_PHILOX_64),
                          (100.0, -1.0), 2, 0.99)


def test_lsm_per_path_seeds(cuda, oracle):
    n, steps = 20_000, 30
    m = gbm_model()
    seeds = np.random.Generator(np.random.Philox(12345)).integers(0, 2**64, size=n, dtype=np.uint64)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, seeds=seeds)
    _check(*_run_both(cuda, oracle, m, sim, (100.0, -1.0), 3, math.exp(-m.r * m.T / steps)))


def test_reference_american_put_test_through_solve(cuda):
    """test/agreement/american_options.jl:9-52: 50 000 x 100, antithetic, degree 5 vs CRR(1000), rtol 0.02."""
    from oracle import anchors as A
    ref, exp = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
    prob = hh.PricingProblem(hh.VanillaOption(100.0, exp, hh.American(), hh.Put(), hh.Spot()),
                             hh.BlackScholesInputs(ref, 0.05, 100.0, 0.2))
    seeds = np.random.Generator(np.random.Philox(12345)).integers(0, 2**64, size=50_000, dtype=np.uint64)
    cfg = hh.SimulationConfig(50_000, steps=100, seeds=seeds, variance_reduction=hh.Antithetic())
    sol = hh.solve(prob, hh.LSM(hh.LognormalDynamics(), hh.BlackScholesExact(), cfg, 5), engine=cuda, spot_paths=True)
    crr = A.crr_price(100.0, 100.0, 0.05, 0.2, 366 / 365, 1000, cp=-1, american=True)
    assert sol.price == pytest.approx(crr, rel=0.02)
    assert sol.spot_paths.shape == (101, 100_000)
    assert np.all(sol.spot_paths[0] == 100.0)
    assert len(sol.stopping_info) == 100_000
    taus = np.array([t for t, _ in sol.stopping_info])
    assert taus.min() >= 1 and taus.max() == 100
    assert sol.price > A.bs_price(100.0, 100.0, 0.05, 0.2, 366 / 365, cp=-1)  # American >= European (:148-202)


def test_lsm_three_sigma_vs_crr_at_scale(cuda):
    """Config C3 parameters at 2e6 paths: within 3 standard errors + the known LSM low bias (<0.5%) of CRR."""
    from oracle import anchors as A
    m = gbm_model(S0=100.0, r=0.05, sigma=0.2, T=1.0)
    steps = 50
    sim = SimSpec(n_paths=2_000_000, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=12345)
    out, *_ = cuda.lsm_american(m, sim, (100.0, -1.0), 3, math.exp(-m.r * m.T / steps))
    crr = A.crr_price(100.0, 100.0, 0.05, 0.2, 1.0, 2000, cp=-1, american=True)
    assert abs(out.price - crr) < 3 * out.std_error + 5e-3 * crr, (out.price, crr, out.std_error)
    assert out.n == 2_000_000 and out.n_dates_skipped == 0


def test_lsm_edge_cases(cuda, oracle):
    m = gbm_model()
    # one exercise date: no regression at all, price = D * mean(payoff(S_T))
    sim = SimSpec(n_paths=1000, n_steps=1, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1)
    _check(*_run_both(cuda, oracle, m, sim, (100.0, -1.0), 3, math.exp(-m.r)))
    # deep out of the money: every date is skipped (least_squares_montecarlo.jl:122), price = 0
    sim = SimSpec(n_paths=999, n_steps=10, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1)
    g, o = _run_both(cuda, oracle, m, sim, (1.0, -1.0), 3, 0.99)
    _check(g, o)
    assert g[0].n_dates_skipped == 9 and g[0].price == 0.0
    # degree 0 and a single trajectory
    sim = SimSpec(n_paths=1, n_steps=5, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1)
    _check(*_run_both(cuda, oracle, m, sim, (100.0, -1.0), 0, 0.99))
    # argument errors
    with pytest.raises(ValueError):
        cuda.lsm_american(m, SimSpec(n_paths=10, n_steps=5, scheme=abi.HH_SCHEME_EXACT_STEPS), (100.0, -1.0), 99, 0.99)
    with pytest.raises(NotImplementedError):  # the terminal-law sampler saves no dates: nothing to regress on
        cuda.lsm_american(m, SimSpec(n_paths=10, n_steps=5, scheme=abi.HH_SCHEME_EXACT_TERMINAL), (100.0, -1.0), 3, 0.99)
    with pytest.raises(NotImplementedError):  # Q5: Antithetic + HestonBroadieKaya
        cuda.lsm_american(heston_model(), SimSpec(n_paths=10, n_steps=5, scheme=abi.HH_SCHEME_HESTON_BK, vr=abi.HH_VR_ANTITHETIC),
                          (100.0, -1.0), 3, 0.99)
    with pytest.raises(NotImplementedError):  # LSM grids are binary64
        cuda.lsm_american(heston_model(), SimSpec(n_paths=10, n_steps=5, scheme=abi.HH_SCHEME_EM, precision=abi.HH_PREC_F32),
                          (100.0, -1.0), 3, 0.99)


def test_peer_exchange_request_without_connection_is_an_argument_error(cuda):
    """hh_comm with no callback and world > 1 selects the in-kernel peer exchange, which needs hh_peer_connect first."""
    m = gbm_model()
    sim = SimSpec(n_paths=1000, n_steps=5, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1)
    comm = abi.hh_comm(abi.hh_allreduce_fn(), None, 0, 2)
    with pytest.raises(ValueError):
        cuda.lsm_american(m, sim, (100.0, -1.0), 2, 0.99, comm=comm)


def test_peer_mailbox_loopback(cuda):
    """A world of one: export, connect to itself, run (the exchange is skipped), disconnect."""
    h = cuda.peer_export()
    assert len(h) == abi.HH_IPC_HANDLE_BYTES and any(h)
    cuda.peer_connect(0, 1, [h])
    m = gbm_model()
    sim = SimSpec(n_paths=2000, n_steps=5, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1)
    a, *_ = cuda.lsm_american(m, sim, (100.0, -1.0), 2, 0.99, comm=abi.hh_comm(abi.hh_allreduce_fn(), None, 0, 1))
    b, *_ = cuda.lsm_american(m, sim, (100.0, -1.0), 2, 0.99)
    assert a.price == b.price
    cuda.peer_disconnect()


def test_persistent_and_per_date_kernels_agree(cuda):
    """The persistent cooperative kernel (default) and the one-launch-per-date form (HH_LSM_PERSISTENT=0, also the path the
    NCCL callback takes) run the same column arithmetic; block partitions differ, so sums agree to rounding and the
    stopping decisions exactly (up to ties)."""
    import json
    import os
    import subprocess
    import sys
    code = r'''
import json, math, sys
sys.path.insert(0, %r)
import numpy as np
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
sys.path.insert(0, %r)
from helpers import gbm_model
eng = hh.default_engine(0)
m = gbm_model()
out = {}
for anti in (0, 1):
    sim = SimSpec(n_paths=300_001, n_steps=25, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=anti, base_seed=9)
    o, tau, val, _ = eng.lsm_american(m, sim, (100.0, -1.0), 3, math.exp(-m.r * m.T / 25), want_stopping=True)
    out[str(anti)] = [o.price, o.std_error, int(tau.sum()), float(val.sum())]
print(json.dumps(out))
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("1", "0"):
        env = dict(os.environ, HH_LSM_PERSISTENT=mode)
        p = subprocess.run([sys.executable, "-c", code % (root, os.path.join(root, "tests"))], env=env, capture_output=True,
                           text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        res[mode] = json.loads(p.stdout.strip().splitlines()[-1])
    for anti in ("0", "1"):
        a, b = res["1"][anti], res["0"][anti]
        assert abs(a[0] - b[0]) <= 1e-9 * abs(b[0]), (a, b)
        assert abs(a[2] - b[2]) <= 25 * 3  # at most a few tie flips


# ---- log-space generators with the corrected S-space extraction ------------------------------------
# The oracle restates the same extension (oracle/hh_oracle.c simulate_one, grid stores exp(x)); the reference itself
# would regress on log S here (least_squares_montecarlo.jl:53), so there is no reference behaviour to match beyond
# the schemes themselves, which the European parity tests pin.

def _normals(n, steps, nc, seed=5):
    return np.random.default_rng(seed).standard_normal((n, steps, nc)) if nc > 1 else np.random.default_rng(seed).standard_normal((n, steps))


@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("model", ["gbm", "heston", "heston_nosplit"])
def test_lsm_logspace_parity_mode(cuda, oracle, model, anti):
    n, steps = 4099, 24
    m = gbm_model() if model == "gbm" else heston_model(split=model == "heston")
    z = np.ascontiguousarray(_normals(n, steps, 1 if model == "gbm" else 2))
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=int(anti), rng_mode=abi.HH_RNG_NORMALS, normals=z)
    D = math.exp(-m.r * m.T / steps)
    g, o = _run_both(cuda, oracle, m, sim, (100.0, -1.0), 3, D)
    _check(g, o, max_flip_frac=5e-4, path_tol=1e-11)
    assert np.all(g[3][:, 0] == m.S0) and np.all(g[3] > 0.0)  # spots, not log-spots


@pytest.mark.parametrize("model", ["gbm", "heston"])
@pytest.mark.parametrize("seeded", [False, True])
def test_lsm_logspace_native_rng(cuda, oracle, model, seeded):
    n, steps = 30_001, 40
    m = gbm_model(sigma=0.3) if model == "gbm" else heston_model(xi=0.5, rho=-0.5)
    kw = dict(seeds=np.random.Generator(np.random.Philox(3)).integers(0, 2**64, size=n, dtype=np.uint64)) if seeded \
        else dict(base_seed=99, path_offset=12345)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=abi.HH_VR_ANTITHETIC, **kw)
    D = math.exp(-m.r * m.T / steps)
    _check(*_run_both(cuda, oracle, m, sim, (105.0, -1.0), 4, D), max_flip_frac=5e-4, path_tol=1e-10)


def test_lsm_gbm_em_equals_exact_steps_on_the_same_normals(cuda):
    """x' = x + (r - s^2/2) dt + s sqrt(dt) Z and S' = S exp(the same) are one law: with the same normals the two
    generators give the same spots to rounding and hence the same price."""
    n, steps = 20_000, 30
    m = gbm_model()
    z = np.random.default_rng(8).standard_normal((n, steps))
    D = math.exp(-m.r * m.T / steps)
    res = {}
    for scheme in (abi.HH_SCHEME_EM, abi.HH_SCHEME_EXACT_STEPS):
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=abi.HH_VR_ANTITHETIC, rng_mode=abi.HH_RNG_NORMALS, normals=z)
        res[scheme] = cuda.lsm_american(m, sim, (100.0, -1.0), 3, D, want_stopping=True, want_paths=True)
    a, b = res[abi.HH_SCHEME_EM], res[abi.HH_SCHEME_EXACT_STEPS]
    assert rel_err(a[3], b[3]) < 1e-12
    assert np.mean(a[1] != b[1]) < 5e-4
    assert abs(a[0].price - b[0].price) < 1e-5 * b[0].price


def test_lsm_heston_degenerate_to_gbm(cuda):
    """xi = 0, V0 = theta: the variance stays at theta and the log-Heston step is the log-GBM step with sigma = sqrt(theta)."""
    n, steps = 20_000, 30
    z2 = np.random.default_rng(9).standard_normal((n, steps, 2))
    mg = gbm_model(r=0.03, sigma=0.2)
    mh = heston_model(r=0.03, V0=0.04, theta=0.04, xi=0.0, rho=0.0)
    D = math.exp(-0.03 / steps)
    sg = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, rng_mode=abi.HH_RNG_NORMALS, normals=np.ascontiguousarray(z2[:, :, 0]))
    sh = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, rng_mode=abi.HH_RNG_NORMALS, normals=z2)
    a = cuda.lsm_american(mg, sg, (100.0, -1.0), 3, D, want_stopping=True, want_paths=True)
    b = cuda.lsm_american(mh, sh, (100.0, -1.0), 3, D, want_stopping=True, want_paths=True)
    assert rel_err(a[3], b[3]) < 1e-12
    assert abs(a[0].price - b[0].price) < 1e-5 * b[0].price


def test_lsm_heston_american_put_through_solve(cuda):
    """solve(PricingProblem{American}, LSM(HestonDynamics, EulerMaruyama)): American >= European (Carr-Madan) and the
    early-exercise premium is of the Black-Scholes order of magnitude at the same volatility level."""
    from oracle import anchors as A
    ref, exp = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
    prob = hh.PricingProblem(hh.VanillaOption(100.0, exp, hh.American(), hh.Put(), hh.Spot()),
                             hh.HestonInputs(ref, 0.05, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7))
    cfg = hh.SimulationConfig(400_000, steps=100, base_seed=2024, variance_reduction=hh.Antithetic())
    sol = hh.solve(prob, hh.LSM(hh.HestonDynamics(), hh.EulerMaruyama(), cfg, 3), engine=cuda, stopping_info=False)
    T = 366 / 365
    call = A.heston_price(100.0, 100.0, 0.05, T, 0.04, 2.0, 0.04, 0.3, -0.7)
    euro_put = call - 100.0 + 100.0 * math.exp(-0.05 * T)
    assert sol.price > euro_put + 3 * sol.std_error
    bs_premium = A.crr_price(100.0, 100.0, 0.05, 0.2, T, 1000, cp=-1, american=True) - A.bs_price(100.0, 100.0, 0.05, 0.2, T, cp=-1)
    assert 0.4 * bs_premium < sol.price - euro_put < 2.0 * bs_premium


# ---- exercise dates simulated exactly (HestonBroadieKaya) --------------------------------------------------------------

def test_lsm_broadie_kaya_grid_is_the_european_path(cuda):
    """The grid comes out of the same path kernel as hh_mc_european under HestonBroadieKaya: date 0 = S0, the last date is
    bit-identical to the European terminal spots; tau in 1..dates."""
    m = heston_model()
    n, dates = 20_000, 6
    sim = SimSpec(n_paths=n, n_steps=dates, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=31)
    out, tau, val, paths = cuda.lsm_american(m, sim, (100.0, -1.0), 3, math.exp(-m.r * m.T / dates), want_stopping=True, want_paths=True)
    _, term = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    assert paths.shape == (n, dates + 1)
    assert np.all(paths[:, 0] == m.S0) and np.array_equal(paths[:, -1], term) and np.all(paths > 0)
    assert tau.min() >= 1 and tau.max() == dates and out.n == n
    with pytest.raises(NotImplementedError):  # Broadie-Kaya draws in-kernel only
        cuda.lsm_american(m, SimSpec(n_paths=10, n_steps=3, scheme=abi.HH_SCHEME_HESTON_BK, rng_mode=abi.HH_RNG_NORMALS,
                                     normals=np.zeros((10, 3, 2))), (100.0, -1.0), 2, 0.99)


def test_lsm_broadie_kaya_bermudan_put_through_solve(cuda):
    """Monthly-exercisable put under Heston: exact transitions between the 12 exercise dates against Euler-Maruyama with
    the same 12 dates (coarse steps, so only loosely equal), and above the European put (Carr-Madan)."""
    from oracle import anchors as A
    ref, exp = dt.date(2020, 1, 1), dt.date(2021, 1, 1)
    T = 366 / 365
    mk = hh.HestonInputs(ref, 0.05, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    put = hh.PricingProblem(hh.VanillaOption(100.0, exp, hh.American(), hh.Put(), hh.Spot()), mk)
    bk = hh.solve(put, hh.LSM(hh.HestonDynamics(), hh.HestonBroadieKaya(), hh.SimulationConfig(400_000, steps=12, base_seed=3), 3),
                  engine=cuda, stopping_info=False)
    em = hh.solve(put, hh.LSM(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(400_000, steps=12, base_seed=4), 3),
                  engine=cuda, stopping_info=False)
    euro_put = A.heston_price(100.0, 100.0, 0.05, T, 0.04, 2.0, 0.04, 0.3, -0.7) - 100.0 + 100.0 * math.exp(-0.05 * T)
    assert bk.price > euro_put + 3 * bk.std_error
    assert abs(bk.price - em.price) < 0.02 * bk.price
    assert bk.stats["n_cols_total"] == 400_000


@pytest.mark.parametrize("case", [
    dict(S0=79.52269401822332, r=0.014583956816892861, sigma=0.5471058154187945, T=1.811123699654538, steps=18, deg=4, anti=1, cp=1.0,
         K=84.6995352661219, n=6444),      # volatile call: 479 of 12888 decisions differed before the per-date Chebyshev interval
    dict(S0=44.27014586154759, r=0.04242860160990872, sigma=0.08016317337524555, T=1.568259129916532, steps=4, deg=4, anti=0, cp=-1.0,
         K=39.0076598804521, n=7957),      # low-volatility out-of-the-money put, degree 4
    dict(S0=23.0020217126248, r=0.0037183838025878343, sigma=0.3367368057225824, T=2.9504215302361443, steps=18, deg=4, anti=1, cp=1.0,
         K=24.128030401113264, n=23345),
])
def test_lsm_regression_is_well_conditioned_on_every_date(cuda, oracle, case):
    """The fit solves normal equations in a Chebyshev variable u = ua_t S + ub_t; the interval mapped to [-1, 1] follows the
    reach of the spot at EACH date. With one interval for all dates these cases (from HH_FUZZ_SCALE=8 tests/test_gpu_fuzz.py)
    had the data of the early dates in a few per cent of [-1, 1], a Gram matrix without rank in binary64, and up to 4 % of
    the decisions (1.4 % of the price) away from the reference's QR fit. Now: no decision differs."""
    m = gbm_model(S0=case["S0"], r=case["r"], sigma=case["sigma"], T=case["T"])
    sim = SimSpec(n_paths=case["n"], n_steps=case["steps"], scheme=abi.HH_SCHEME_EXACT_STEPS, vr=case["anti"], base_seed=2024)
    D = math.exp(-m.r * m.T / case["steps"])
    og, tg, vg, pg = cuda.lsm_american(m, sim, (case["K"], case["cp"]), case["deg"], D, want_stopping=True, want_paths=True)
    oo, to, vo, po = oracle.lsm_american(m, sim, (case["K"], case["cp"]), case["deg"], D, want_stopping=True, want_paths=True)
    assert rel_err(pg, po) < 1e-12
    assert int(np.sum(tg != to)) <= 1
    assert abs(og.price - oo.price) <= 1e-9 * abs(oo.price) + (5e-4 * abs(oo.price) if np.any(tg != to) else 0.0)


def test_american_call_without_dividends_is_the_european_call(cuda):
    """A known answer that needs no oracle: early exercise of a call on a non-dividend-paying asset is never optimal (r > 0),
    so LSM — a lower bound in expectation — must land just below Black-Scholes. The volatile set of the conditioning test above:
    with the single Chebyshev interval the early-date fits were noise and the price fell 1.4 % short."""
    from oracle import anchors as A
    S0, K, r, sigma, T = 79.52269401822332, 84.6995352661219, 0.014583956816892861, 0.5471058154187945, 1.811123699654538
    m = gbm_model(S0=S0, r=r, sigma=sigma, T=T)
    steps = 18
    sim = SimSpec(n_paths=400_000, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=abi.HH_VR_ANTITHETIC, base_seed=77)
    out = cuda.lsm_american(m, sim, (K, 1.0), 4, math.exp(-r * T / steps))[0]
    bs = A.bs_price(S0, K, r, sigma, T, 1.0)
    assert out.price < bs + 3.5 * out.std_error
    assert out.price > bs * (1 - 5e-3) - 3.5 * out.std_error, (out.price, bs, out.std_error)
