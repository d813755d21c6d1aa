"""GPU parity of the opt-in HH_RNG_PHILOX_64 stream (one Philox block per two Heston steps): the CUDA kernel against the
oracle's restatement of the same stream, per path, through the C ABI; then the statistical bar (3 sigma vs Carr-Madan)."""
import math

import numpy as np
import pytest

from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seeds_mode", ["base", "per_path"])
@pytest.mark.parametrize("anti", [False, True])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("steps", [64, 37])   # odd: the last Philox block feeds one step only
def test_matches_oracle_stream_per_path(cuda, oracle, seeds_mode, anti, split, steps):
    n = 10_000
    m = heston_model(split=split)
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=int(anti), rng_mode=abi.HH_RNG_PHILOX_64,
                  base_seed=42, path_offset=123456789012)
    if seeds_mode == "per_path":
        sim.seeds = np.random.Generator(np.random.Philox(5)).integers(0, 2**64, size=n, dtype=np.uint64)
    D = math.exp(-m.r * m.T)
    pay = [(100.0, 1.0), (90.0, -1.0)]
    rg, tg = cuda.mc_european(m, sim, pay, D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, pay, D, want_terminal=True)
    assert rel_err(tg, to) < 1e-11
    for a, b in zip(rg, ro):
        assert abs(a.price - b.price) <= 1e-11 * abs(b.price)
        assert a.n_nonfinite == 0


def test_large_job_block_shape_matches_oracle(cuda, oracle):
    """Jobs of >= 4 x 1024 trajectories per SM take the 1024-thread instantiation (the one bench.py times)."""
    sm = cuda.device_info()["sm_count"]
    n, steps = sm * 1024 * 4 + 777, 16
    m = heston_model()
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, rng_mode=abi.HH_RNG_PHILOX_64, base_seed=3)
    D = math.exp(-m.r * m.T)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], D, want_terminal=True)
    assert rel_err(tg, to) < 1e-11
    assert abs(rg[0].sum - ro[0].sum) <= 1e-11 * abs(ro[0].sum)


def test_three_sigma_against_carr_madan_and_the_default_stream(cuda):
    from oracle import anchors as A
    m = heston_model()
    ref = A.heston_price(100.0, 100.0, m.r, m.T, m.V0, m.kappa, m.theta, m.xi, m.rho)
    D = math.exp(-m.r * m.T)
    n, steps = 8_000_000, 252
    r64, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=steps, rng_mode=abi.HH_RNG_PHILOX_64, base_seed=11), [(100.0, 1.0)], D)
    r52, _ = cuda.mc_european(m, SimSpec(n_paths=n, n_steps=steps, rng_mode=abi.HH_RNG_PHILOX, base_seed=11), [(100.0, 1.0)], D)
    bias = 0.005651  # Euler-Maruyama at 252 steps (profiles/r1_i_euler_bias_c2.json)
    assert abs(r64[0].price - ref - bias) < 3 * r64[0].std_error + 1.5e-3
    assert abs(r64[0].price - r52[0].price) < 3 * math.hypot(r64[0].std_error, r52[0].std_error)
    assert r64[0].n_nonfinite == 0


def test_shard_invariance(cuda):
    n, steps = 80_000, 17
    m = heston_model()
    D = math.exp(-m.r * m.T)
    mk = lambda k, off: SimSpec(n_paths=k, path_offset=off, n_steps=steps, rng_mode=abi.HH_RNG_PHILOX_64, base_seed=5)
    full, tfull = cuda.mc_european(m, mk(n, 0), [(100.0, 1.0)], D, want_terminal=True)
    parts = []
    for g in range(4):
        lo, hi = n * g // 4, n * (g + 1) // 4
        parts.append(cuda.mc_european(m, mk(hi - lo, lo), [(100.0, 1.0)], D, want_terminal=True)[1])
    assert np.array_equal(np.concatenate(parts), tfull)


def test_scope_is_heston_em_f64_pricing(cuda):
    D = 1.0
    with pytest.raises(NotImplementedError):
        cuda.mc_european(gbm_model(), SimSpec(n_paths=100, n_steps=4, rng_mode=abi.HH_RNG_PHILOX_64), [(100.0, 1.0)], D)
    with pytest.raises(NotImplementedError):
        cuda.mc_european(heston_model(), SimSpec(n_paths=100, n_steps=4, rng_mode=abi.HH_RNG_PHILOX_64,
                                                 precision=abi.HH_PREC_F32), [(100.0, 1.0)], D)
    t = abi.hh_tangent()
    t.dS0 = 1.0
    with pytest.raises(NotImplementedError):
        cuda.tangent_sums(heston_model(), [t], SimSpec(n_paths=100, n_steps=4, rng_mode=abi.HH_RNG_PHILOX_64), [(100.0, 1.0)])


def test_ablation_entry_point_runs(cuda):
    """hh_bench_heston_ablation: every part of both streams launches and reports a device time."""
    for rng in (abi.HH_RNG_PHILOX, abi.HH_RNG_PHILOX_64):
        ms = [cuda.heston_ablation(2_000_000, 64, rng, part) for part in (0, 1, 2)]
        assert all(x > 0 for x in ms)
        assert ms[1] < ms[0] and ms[2] < ms[0]      # each part alone is cheaper than the whole step
    with pytest.raises(ValueError):
        cuda.heston_ablation(1000, 4, abi.HH_RNG_NORMALS, 0)
