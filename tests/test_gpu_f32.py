"""Float32 fast mode of the Heston Euler-Maruyama kernel (config C2 "Float32 fast mode"), through the C ABI.

Bars: the f32 mode computes with MUFU approximations (lg2, sin, cos, sqrt), so per-path agreement with the oracle's
binary32 restatement (libm) is ~1e-5 typical, not bitwise; prices must fall within 3 standard errors of Carr-Madan
and of the Float64 kernel; payoff sums are accumulated in f64 and are bit-reproducible run to run."""
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model
from oracle import anchors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("anti", [0, 1])
@pytest.mark.parametrize("steps", [252, 13])  # odd step count: the last Philox block is half used
def test_f32_paths_track_the_binary32_oracle(cuda, oracle, anti, steps):
    m = heston_model()
    sim = SimSpec(n_paths=20_000, n_steps=steps, vr=anti, precision=abi.HH_PREC_F32, base_seed=99)
    D = math.exp(-m.r * m.T)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    rel = np.abs(tg - to) / to
    assert np.median(rel) < 2e-6, np.median(rel)
    assert np.mean(rel) < 5e-5, np.mean(rel)
    assert np.quantile(rel, 0.999) < 5e-3, np.quantile(rel, 0.999)
    assert abs(rg[0].price - ro[0].price) < 0.1 * ro[0].std_error
    assert rg[0].n_nonfinite == 0


def test_f32_price_within_3_sigma_of_carr_madan_and_f64(cuda):
    m = heston_model()
    D = math.exp(-m.r * m.T)
    cm = anchors.heston_price(100.0, 100.0, m.r, m.T, m.V0, m.kappa, m.theta, m.xi, m.rho)
    n = 4_000_000
    r32, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=252, precision=abi.HH_PREC_F32, base_seed=5), [(100.0, 1.0)], D)
    r64, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=252, base_seed=5), [(100.0, 1.0)], D)
    assert abs(r32[0].price - cm) < 3 * r32[0].std_error + 0.01  # + the Euler bias at 252 steps
    assert abs(r32[0].price - r64[0].price) < 3 * math.hypot(r32[0].std_error, r64[0].std_error)
    again, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=252, precision=abi.HH_PREC_F32, base_seed=5), [(100.0, 1.0)], D)
    assert (again[0].sum, again[0].sumsq) == (r32[0].sum, r32[0].sumsq)


def test_f32_through_solve_and_shards(cuda):
    import datetime as dt
    payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
    market = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    method = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(100_001, steps=50, base_seed=3),
                           precision="f32", ensemble=True)
    whole = hh.solve(hh.PricingProblem(payoff, market), method, engine=cuda, shard=(0, 1))
    parts = [hh.solve(hh.PricingProblem(payoff, market), method, engine=cuda, shard=(r, 2)) for r in range(2)]
    np.testing.assert_array_equal(np.concatenate([p.ensemble for p in parts]), whole.ensemble)  # shard-invariant streams


def test_f32_is_rejected_where_it_is_not_built(cuda):
    with pytest.raises(NotImplementedError):
        cuda.mc_european(gbm_model(), SimSpec(n_paths=1000, n_steps=10, precision=abi.HH_PREC_F32), [(100.0, 1.0)], 1.0)
