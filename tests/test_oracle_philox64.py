"""CPU tests of the opt-in HH_RNG_PHILOX_64 stream as the ORACLE restates it (oracle/hh_oracle.c: hho_normal_pair64):
one Philox4x32-10 block per two Heston steps, a 32-bit radius uniform and a 32-bit angle per step. The stream is this
build's own convention (the reference draws from Xoshiro inside third-party packages), so what has to hold is that
(i) the bits map to the documented uniforms, (ii) the normals are standard normal and independent across steps and
components, and (iii) prices agree with Carr-Madan within 3 standard errors (plus the scheme's own time-stepping bias)."""
import math

import numpy as np
import pytest
from scipy import stats

from oracle import anchors as A
from oracle import oracle as O


def _pairs(n_paths, n_steps, key=42):
    m = O.heston_model(100.0, 0.03, 1.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    sim = O.OSim(n_paths=n_paths, n_steps=n_steps, scheme=O.HH_SCHEME_EM, rng_mode=O.HH_RNG_PHILOX_64, base_seed=key)
    return O.OracleEngine().fill_normals(m, sim)  # [path, step, 2]


def test_bits_to_uniforms_are_the_documented_maps():
    for key, idx, step in [(1, 0, 0), (42, 123456789012, 1), (2**64 - 1, 7, 250), (9, 2**40 + 3, 251)]:
        w = O.philox([idx & 0xFFFFFFFF, idx >> 32, step >> 1, 2], [key & 0xFFFFFFFF, key >> 32])
        wa, wb = w[(step & 1) * 2], w[(step & 1) * 2 + 1]
        rotl = lambda x: ((x << 12) | (x >> 20)) & 0xFFFFFFFF
        u1 = 1.0 - (rotl(wb) + 0.5) * 2.0**-32            # midpoints of the 2^32 grid
        theta = 2 * math.pi * (rotl(wa) + (wa & 0xFFFFF) * 2.0**-20) * 2.0**-32
        r = math.sqrt(-2.0 * math.log(u1))
        z1, z2 = O.normal_pair64(key, idx, step)
        assert abs(z1 - r * math.cos(theta)) < 5e-15 * max(1.0, r)
        assert abs(z2 - r * math.sin(theta)) < 5e-15 * max(1.0, r)


def test_disjoint_from_the_default_stream():
    a = O.normal_pair(42, 5, 0)
    b = O.normal_pair64(42, 5, 0)
    assert abs(a[0] - b[0]) > 1e-6 and abs(a[1] - b[1]) > 1e-6


def test_normals_are_standard_and_independent():
    z = _pairs(40_000, 50)            # 4e6 normals
    flat = z.reshape(-1)
    n = flat.size
    assert abs(flat.mean()) < 4 / math.sqrt(n)
    assert abs(flat.var() - 1) < 4 * math.sqrt(2 / n)
    assert abs(stats.skew(flat)) < 4 * math.sqrt(6 / n)
    assert abs(stats.kurtosis(flat)) < 4 * math.sqrt(24 / n)
    assert np.abs(flat).max() < 6.78                     # u1 >= 2^-33
    ks = stats.kstest(flat[:1_000_000], "norm")
    assert ks.pvalue > 1e-3, ks
    # components of a step, consecutive steps sharing one Philox block, and steps of neighbouring blocks
    lim = 4 / math.sqrt(z.shape[0] * (z.shape[1] - 2))
    assert abs(np.mean(z[:, :, 0] * z[:, :, 1])) < lim
    for lag in (1, 2):
        for ca in (0, 1):
            for cb in (0, 1):
                assert abs(np.mean(z[:, :-lag, ca] * z[:, lag:, cb])) < lim, (lag, ca, cb)
    # squared radii of the two halves of a block are independent too
    r2 = (z ** 2).sum(axis=2)
    c = np.corrcoef(r2[:, 0::2].reshape(-1), r2[:, 1::2].reshape(-1))[0, 1]
    assert abs(c) < 4 / math.sqrt(r2.size / 2)
    # radius^2 / 2 is Exp(1), the angle is uniform
    assert stats.kstest(r2.reshape(-1)[:1_000_000] / 2, "expon").pvalue > 1e-3
    ang = (np.arctan2(z[:, :, 1], z[:, :, 0]).reshape(-1)[:1_000_000] / (2 * math.pi)) % 1.0
    assert stats.kstest(ang, "uniform").pvalue > 1e-3


@pytest.mark.parametrize("anti", [0, 1])
def test_price_within_three_sigma_of_carr_madan(anti):
    S0, K, r, T, V0, kappa, theta, xi, rho = 100.0, 100.0, 0.03, 1.0, 0.04, 2.0, 0.04, 0.3, -0.7
    m = O.heston_model(S0, r, T, V0, kappa, theta, xi, rho)
    ref = A.heston_price(S0, K, r, T, V0, kappa, theta, xi, rho)
    eng = O.OracleEngine()
    D = math.exp(-r * T)
    sim64 = O.OSim(n_paths=200_000, n_steps=100, scheme=O.HH_SCHEME_EM, vr=anti, rng_mode=O.HH_RNG_PHILOX_64, base_seed=7)
    sim52 = O.OSim(n_paths=200_000, n_steps=100, scheme=O.HH_SCHEME_EM, vr=anti, rng_mode=O.HH_RNG_PHILOX, base_seed=7)
    r64 = eng.mc_european(m, sim64, [(K, 1.0)], D)[0][0]
    r52 = eng.mc_european(m, sim52, [(K, 1.0)], D)[0][0]
    bias = 0.015  # Euler-Maruyama at 100 steps (profiles/r1_i_euler_bias_c2.json: +0.0112 at 126, +0.0217 at 63)
    assert abs(r64.price - ref - bias) < 3 * r64.std_error + 0.01
    # the two streams are independent estimates of the same expectation
    assert abs(r64.price - r52.price) < 3 * math.hypot(r64.std_error, r52.std_error)
    assert r64.n_nonfinite == 0


def test_only_heston_em_f64():
    eng = O.OracleEngine()
    g = O.o_model()
    g.kind, g.S0, g.r, g.T, g.sigma = O.HH_MODEL_GBM, 100.0, 0.05, 1.0, 0.2
    with pytest.raises(NotImplementedError):
        eng.mc_european(g, O.OSim(n_paths=10, n_steps=4, rng_mode=O.HH_RNG_PHILOX_64), [(100.0, 1.0)], 1.0)


def test_lsm_generator_stream_four_steps_per_block(oracle):
    """The exact GBM generator under HH_RNG_PHILOX_64 takes ONE normal per step: step n is component n & 1 of the 64-bit
    Box-Muller pair hho_normal_pair64(key, idx, n >> 1), i.e. one Philox block per four steps. Law and independence."""
    from scipy import stats
    m = O.o_model()
    m.kind, m.flags = O.HH_MODEL_GBM, O.HH_FLAG_SPLIT_STEP | O.HH_FLAG_Q1_SQRT_MEAN
    m.S0, m.r, m.T, m.sigma = 100.0, 0.05, 1.0, 0.2
    z = oracle.fill_normals(m, O.OSim(n_paths=100_000, n_steps=9, scheme=O.HH_SCHEME_EXACT_STEPS, rng_mode=O.HH_RNG_PHILOX_64,
                                      base_seed=3))[:, :, 0]
    for n in range(9):
        assert stats.kstest(z[:, n], "norm").pvalue > 1e-3
        a, b = O.normal_pair64(3, 17, n >> 1)
        assert z[17, n] == (b if n & 1 else a)
    c = np.corrcoef(z.T)
    assert np.max(np.abs(c - np.eye(9))) < 0.015
    z0 = oracle.fill_normals(m, O.OSim(n_paths=1000, n_steps=9, scheme=O.HH_SCHEME_EXACT_STEPS, base_seed=3))[:, :, 0]
    assert not np.allclose(z0, z[:1000])   # counter stream word 2: never the HH_RNG_PHILOX numbers
