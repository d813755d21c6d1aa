"""CPU checks of the quasi-random restatement (oracle.quasi_random_normals = HH_VR_QUASI_RANDOM, include/hedgehog_mc.h):
the roadmap's "Stratified sampling / quasi-random" (docs/src/derivatives_pricing_roadmap.md:164) for the one-draw exact
sampler of LognormalDynamics + BlackScholesExact (montecarlo.jl:293-303, 454-459)."""
import math

import numpy as np

from oracle import oracle as O

BS_CALL = 10.450583572185565   # Black-Scholes call S = K = 100, r = 0.05, sigma = 0.2, T = 1 (tests/golden/config_anchors.json)


def gbm(T=1.0):
    m = O.o_model()
    m.kind, m.flags = O.HH_MODEL_GBM, O.HH_FLAG_SPLIT_STEP | O.HH_FLAG_Q1_SQRT_MEAN
    m.S0, m.r, m.T, m.sigma = 100.0, 0.05, T, 0.2
    return m


def price(oracle, z, strike=100.0, cp=1.0, T=1.0):
    sim = O.OSim(n_paths=z.shape[0], n_steps=1, scheme=O.HH_SCHEME_EXACT_TERMINAL, rng_mode=O.HH_RNG_NORMALS,
                 normals=np.ascontiguousarray(z.reshape(-1, 1, 1)))
    res, _ = oracle.mc_european(gbm(T), sim, [(strike, cp)], math.exp(-0.05 * T))
    return res[0].price, res[0].std_error


def test_points_are_shard_invariant_and_seed_dependent():
    full = O.quasi_random_normals(42, 0, 4096)
    parts = np.concatenate([O.quasi_random_normals(42, 1024 * k, 1024) for k in range(4)])
    assert np.array_equal(full, parts)
    assert not np.array_equal(full, O.quasi_random_normals(43, 0, 4096))
    assert np.all(np.isfinite(full))


def test_low_discrepancy_in_every_prefix():
    """Any N consecutive van der Corput points have star discrepancy O(log N / N), rotation or not: the empirical CDF of the
    uniforms stays within that of the diagonal, far inside the 1 / sqrt(N) of pseudo-random draws."""
    from scipy.special import ndtr
    for n in (1000, 4096, 50_000):
        u = np.sort(ndtr(O.quasi_random_normals(7, 12345, n)))
        d = np.max(np.abs(u - (np.arange(n) + 0.5) / n))
        assert d <= (math.log2(n) + 2) / n


def test_price_error_is_far_below_the_monte_carlo_standard_error(oracle):
    n = 1 << 16
    errs = []
    for seed in range(8):
        p, se = price(oracle, O.quasi_random_normals(seed, 0, n))
        errs.append(p - BS_CALL)
        assert abs(p - BS_CALL) < 0.02 * se    # a pseudo-random run of this size sits ~1 standard error away
    # the rotations are unbiased: the eight errors share no offset beyond their own spread
    assert abs(np.mean(errs)) <= 3 * np.std(errs) / math.sqrt(len(errs)) + 1e-6
