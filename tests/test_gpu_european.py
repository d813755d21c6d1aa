"""GPU parity tests for the European Monte Carlo kernels, through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star):
  parity mode  — the GPU consumes pre-generated normals; per-path terminal values and prices must match the
                 oracle to rel 1e-12 in Float64;
  native RNG   — same Philox stream on both sides, so terminal values are also compared path by path (1e-11),
                 and prices must sit within 3 standard errors of the analytic / Carr-Madan value.
"""
import math

import numpy as np
import pytest

from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err

pytestmark = pytest.mark.gpu

PARITY_RTOL = 1e-12


def _normals(n, steps, ncomp, seed=7):
    return np.random.Generator(np.random.Philox(seed)).standard_normal((n, steps, ncomp))


@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("corr", ["cholesky", "sym_sqrt", "svd"])
def test_heston_em_parity_mode(cuda, oracle, anti, split, corr):
    n, steps = 4096, 252
    m = heston_model(corr=corr, split=split)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=int(anti), rng_mode=abi.HH_RNG_NORMALS,
                  normals=_normals(n, steps, 2))
    D = math.exp(-m.r * m.T)
    pay = [(100.0, 1.0)]
    rg, tg = cuda.mc_european(m, sim, pay, D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, pay, D, want_terminal=True)
    assert rel_err(tg, to) < PARITY_RTOL
    assert abs(rg[0].price - ro[0].price) <= PARITY_RTOL * abs(ro[0].price)
    assert abs(rg[0].sumsq - ro[0].sumsq) <= PARITY_RTOL * abs(ro[0].sumsq)
    assert rg[0].n == ro[0].n == n and rg[0].n_nonfinite == 0


def test_heston_em_parity_q8_parameters(cuda, oracle):
    """Q8: the reference's second Heston test effectively runs V0=1.5, kappa=0.04, theta=0.3, sigma=-0.6, rho=0.04."""
    n, steps = 2048, 200
    m = heston_model(S0=100.0, r=0.05, T=364 / 365, V0=1.5, kappa=0.04, theta=0.3, xi=-0.6, rho=0.04)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=abi.HH_VR_ANTITHETIC, rng_mode=abi.HH_RNG_NORMALS,
                  normals=_normals(n, steps, 2, seed=11))
    D = math.exp(-m.r * m.T)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    assert rel_err(tg, to) < PARITY_RTOL
    assert abs(rg[0].price - ro[0].price) <= PARITY_RTOL * abs(ro[0].price)


@pytest.mark.parametrize("scheme", [abi.HH_SCHEME_EM, abi.HH_SCHEME_EXACT_TERMINAL, abi.HH_SCHEME_EXACT_STEPS])
@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("steps", [1, 7, 50])
def test_gbm_parity_mode(cuda, oracle, scheme, anti, steps):
    n = 5000  # ragged: not a multiple of the 256-trajectory batch
    m = gbm_model(T=366 / 365)
    ns = 1 if scheme == abi.HH_SCHEME_EXACT_TERMINAL else steps
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=int(anti), rng_mode=abi.HH_RNG_NORMALS,
                  normals=_normals(n, ns, 1, seed=3))
    D = math.exp(-m.r * m.T)
    pay = [(100.0, 1.0), (95.0, -1.0)]
    rg, tg = cuda.mc_european(m, sim, pay, D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, pay, D, want_terminal=True)
    assert rel_err(tg, to) < PARITY_RTOL
    for a, b in zip(rg, ro):
        assert abs(a.price - b.price) <= PARITY_RTOL * abs(b.price)
        assert abs(a.std_error - b.std_error) <= 1e-9 * abs(b.std_error)


@pytest.mark.parametrize("seeds_mode", ["base", "per_path"])
@pytest.mark.parametrize("anti", [False, True])
def test_heston_em_native_rng_matches_oracle_stream(cuda, oracle, seeds_mode, anti):
    """Same Philox4x32-10 + Box-Muller on both sides: compared per path, not just statistically."""
    n, steps = 10_000, 64
    m = heston_model()
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=int(anti), base_seed=42, path_offset=123456789012)
    if seeds_mode == "per_path":
        sim.seeds = np.random.Generator(np.random.Philox(5)).integers(0, 2**64, size=n, dtype=np.uint64)
    D = math.exp(-m.r * m.T)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    assert rel_err(tg, to) < 1e-11
    assert abs(rg[0].price - ro[0].price) <= 1e-11 * abs(ro[0].price)


def test_gbm_native_rng_matches_oracle_stream(cuda, oracle):
    m = gbm_model()
    D = math.exp(-m.r * m.T)
    for scheme, steps in [(abi.HH_SCHEME_EM, 33), (abi.HH_SCHEME_EXACT_STEPS, 50), (abi.HH_SCHEME_EXACT_TERMINAL, 1)]:
        sim = SimSpec(n_paths=7777, n_steps=steps, scheme=scheme, vr=abi.HH_VR_ANTITHETIC, base_seed=99)
        rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
        ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
        assert rel_err(tg, to) < 1e-11, scheme
        assert abs(rg[0].price - ro[0].price) <= 1e-11 * abs(ro[0].price)


def test_strike_grid_on_common_paths(cuda, oracle):
    """64 strikes priced by ONE simulation equal 64 single-strike runs on the same stream (and the oracle)."""
    n, steps = 20_000, 32
    m = heston_model()
    strikes = np.linspace(60, 140, 64)
    pay = [(k, 1.0) for k in strikes] + [(k, -1.0) for k in strikes[:7]]  # 71 payoffs -> padded to 128
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, base_seed=1)
    D = math.exp(-m.r * m.T)
    rg, _ = cuda.mc_european(m, sim, pay, D)
    ro, _ = oracle.mc_european(m, sim, pay, D)
    for a, b in zip(rg, ro):
        assert abs(a.price - b.price) <= 1e-11 * max(abs(b.price), 1e-3)
    single, _ = cuda.mc_european(m, sim, [pay[10]], D)
    assert abs(single[0].price - rg[10].price) <= 1e-13 * abs(single[0].price)


def test_shard_invariance(cuda):
    """8 shards run one after the other on one GPU reproduce the 1-GPU sums (Philox keyed by global path index)."""
    n, steps = 80_000, 16
    m = heston_model()
    D = math.exp(-m.r * m.T)
    full, tfull = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=steps, base_seed=5), [(100.0, 1.0)], D, want_terminal=True)
    s = q = 0.0
    parts = []
    for g in range(8):
        lo, hi = n * g // 8, n * (g + 1) // 8
        r, t = cuda.mc_european(m, SimSpec(n_paths=hi - lo, path_offset=lo, n_steps=steps, base_seed=5), [(100.0, 1.0)], D,
                                want_terminal=True)
        s += r[0].sum
        q += r[0].sumsq
        parts.append(t)
    assert np.array_equal(np.concatenate(parts), tfull)  # bit-identical paths
    assert abs(s - full[0].sum) <= 1e-13 * abs(full[0].sum)
    assert abs(q - full[0].sumsq) <= 1e-13 * abs(full[0].sumsq)


def test_three_sigma_against_analytic_and_carr_madan(cuda):
    from oracle import anchors as A
    # GBM, reference set-up of test/agreement/montecarlo_black_scholes.jl: S=K=100, r=.05, sigma=.2
    m = gbm_model(T=1.0)
    D = math.exp(-m.r * m.T)
    bs = A.bs_price(100, 100, 0.05, 0.2, 1.0)
    for scheme, steps in [(abi.HH_SCHEME_EXACT_TERMINAL, 1), (abi.HH_SCHEME_EM, 10), (abi.HH_SCHEME_EXACT_STEPS, 10)]:
        r, _ = cuda.mc_european(m, SimSpec(n_paths=2_000_000, n_steps=steps, scheme=scheme, base_seed=42), [(100.0, 1.0)], D)
        assert abs(r[0].price - bs) < 3 * r[0].std_error, (scheme, r[0].price, bs, r[0].std_error)
    # Heston EM, reference set-up of test/agreement/montecarlo_heston.jl:13-22 vs Carr-Madan(1, 32)
    mh = heston_model()
    Dh = math.exp(-mh.r * mh.T)
    cm = A.heston_price(100, 100, 0.03, 1.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    r, _ = cuda.mc_european(mh, SimSpec(n_paths=4_000_000, n_steps=252, base_seed=42), [(100.0, 1.0)], Dh)
    # Euler bias at 252 steps is below the MC error here; 3 sigma + 0.2% discretisation allowance
    assert abs(r[0].price - cm) < 3 * r[0].std_error + 2e-3 * cm, (r[0].price, cm, r[0].std_error)


def test_edge_cases(cuda, oracle):
    m = heston_model()
    D = 1.0
    # one trajectory, one step
    sim = SimSpec(n_paths=1, n_steps=1, base_seed=3)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    assert rel_err(tg, to) < 1e-12 and rg[0].std_error == 0.0
    # exactly one batch, and one over
    for n in (256, 257):
        sim = SimSpec(n_paths=n, n_steps=3, base_seed=3)
        rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
        ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
        assert rel_err(tg, to) < 1e-12
    # argument errors mirror the reference's ArgumentError
    with pytest.raises(ValueError):
        cuda.mc_european(m, SimSpec(n_paths=0, n_steps=1), [(100.0, 1.0)], D)
    with pytest.raises(ValueError):
        cuda.mc_european(m, SimSpec(n_paths=10, n_steps=0), [(100.0, 1.0)], D)
    with pytest.raises(ValueError):  # seeds shorter than trajectories (montecarlo.jl:65-66)
        cuda.mc_european(m, SimSpec(n_paths=10, n_steps=1, seeds=np.arange(5, dtype=np.uint64)), [(100.0, 1.0)], D)
    with pytest.raises(NotImplementedError):  # Q5
        cuda.mc_european(m, SimSpec(n_paths=10, n_steps=1, scheme=abi.HH_SCHEME_HESTON_BK, vr=abi.HH_VR_ANTITHETIC),
                         [(100.0, 1.0)], D)
