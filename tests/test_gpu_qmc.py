"""HH_VR_QUASI_RANDOM through the C ABI: randomised van der Corput points in the one-draw exact sampler (the roadmap's
"Stratified sampling / quasi-random", docs/src/derivatives_pricing_roadmap.md:164; LognormalDynamics + BlackScholesExact,
montecarlo.jl:293-303, 454-459). The oracle restates the points in numpy (oracle.quasi_random_normals) and consumes them in
parity mode; the kernel must reproduce the terminal spots per path."""
import datetime as dt
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu
BS_CALL = 10.450583572185565


def test_per_path_parity_with_the_restated_points(cuda, oracle):
    n, seed = 20_001, 77
    g = gbm_model(T=366 / 365)   # Q1 visible
    res, t = cuda.mc_european(g, SimSpec(n_paths=n, n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL, vr=abi.HH_VR_QUASI_RANDOM,
                                         base_seed=seed, path_offset=500), [(100.0, 1.0), (90.0, -1.0)], 0.95, want_terminal=True)
    z = O.quasi_random_normals(seed, 500, n)
    m = O.o_model()
    m.kind, m.flags = O.HH_MODEL_GBM, O.HH_FLAG_SPLIT_STEP | O.HH_FLAG_Q1_SQRT_MEAN
    m.S0, m.r, m.T, m.sigma = g.S0, g.r, g.T, g.sigma
    ores, ot = oracle.mc_european(m, O.OSim(n_paths=n, n_steps=1, scheme=O.HH_SCHEME_EXACT_TERMINAL, rng_mode=O.HH_RNG_NORMALS,
                                            normals=np.ascontiguousarray(z.reshape(-1, 1, 1))), [(100.0, 1.0), (90.0, -1.0)], 0.95,
                                  want_terminal=True)
    assert rel_err(t, ot) < 1e-12
    for a, b in zip(res, ores):
        assert a.price == pytest.approx(b.price, rel=1e-11) and a.n == b.n


def test_error_against_black_scholes_and_shard_invariance(cuda):
    n = 1 << 20
    g = gbm_model()
    D = math.exp(-0.05)
    full, t = cuda.mc_european(g, SimSpec(n_paths=n, n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL, vr=abi.HH_VR_QUASI_RANDOM,
                                          base_seed=3), [(100.0, 1.0)], D, want_terminal=True)
    plain, _ = cuda.mc_european(g, SimSpec(n_paths=n, n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL, base_seed=3), [(100.0, 1.0)], D)
    assert abs(full[0].price - BS_CALL) < 0.01 * plain[0].std_error          # 1e6 points: error ~1e-5 against ~1.4e-2
    assert abs(full[0].price - BS_CALL) < 0.05 * abs(plain[0].price - BS_CALL) + 1e-5
    parts = [cuda.mc_european(g, SimSpec(n_paths=n // 4, path_offset=k * (n // 4), n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL,
                                         vr=abi.HH_VR_QUASI_RANDOM, base_seed=3), [(100.0, 1.0)], D, want_terminal=True)[1] for k in range(4)]
    assert np.array_equal(np.concatenate(parts), t)


def test_only_defined_for_the_one_draw_sampler(cuda):
    with pytest.raises(NotImplementedError):
        cuda.mc_european(gbm_model(), SimSpec(n_paths=100, n_steps=4, scheme=abi.HH_SCHEME_EM, vr=abi.HH_VR_QUASI_RANDOM), [(100.0, 1.0)], 1.0)
    with pytest.raises(NotImplementedError):
        cuda.mc_european(heston_model(), SimSpec(n_paths=100, n_steps=4, vr=abi.HH_VR_QUASI_RANDOM), [(100.0, 1.0)], 1.0)


def test_through_solve_with_greeks(cuda):
    payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
    market = hh.BlackScholesInputs(dt.date(2020, 1, 1), 0.05, 100.0, 0.2)
    prob = hh.PricingProblem(payoff, market)
    m = hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(),
                      hh.SimulationConfig(1 << 18, base_seed=11, variance_reduction=hh.QuasiRandom()), ensemble=False)
    sol = hh.solve(prob, m, engine=cuda)
    assert sol.price == pytest.approx(BS_CALL, abs=2e-4)
    delta = hh.solve(hh.GreekProblem(prob, hh.SpotLens()), hh.ForwardAD(), m, engine=cuda).greek
    d1 = (0.05 + 0.5 * 0.04) / 0.2
    assert delta == pytest.approx(0.5 * (1 + math.erf(d1 / math.sqrt(2))), abs=2e-4)
