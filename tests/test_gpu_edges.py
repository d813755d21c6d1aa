"""Edge cases of the hot path through the C ABI: ragged sizes around the block and batch boundaries of every kernel,
the maximum strike count, degenerate models, and the error behaviour that mirrors the reference."""
import math

import numpy as np
import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 31, 255, 256, 257, 511, 513, 1023, 1025, 2049])
@pytest.mark.parametrize("prec", [abi.HH_PREC_F64, abi.HH_PREC_F32])
def test_ragged_path_counts_around_block_boundaries(cuda, oracle, n, prec):
    m = heston_model()
    sim = SimSpec(n_paths=n, n_steps=7, precision=prec, base_seed=11)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0), (95.0, -1.0)], 0.97, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0), (95.0, -1.0)], 0.97, want_terminal=True)
    assert tg.shape == to.shape == (n,)
    assert rel_err(tg, to) < (1e-11 if prec == abi.HH_PREC_F64 else 5e-3)
    for a, b in zip(rg, ro):
        assert a.n == b.n == n
        assert abs(a.price - b.price) <= (1e-11 if prec == abi.HH_PREC_F64 else 1e-3) * max(abs(b.price), 1.0)


def test_large_job_takes_the_1024_thread_blocks_and_matches_the_small_block_sums(cuda):
    """>= 4 x SMs x 1024 trajectories switch the headline kernel to 1024-thread blocks: same trajectories, the payoff sums
    differ only by summation order."""
    m = heston_model()
    n = 148 * 1024 * 4 + 77
    sim = SimSpec(n_paths=n, n_steps=10, base_seed=2)
    big, tb = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    halves = [cuda.mc_european(m, SimSpec(n_paths=h, n_steps=10, base_seed=2, path_offset=o), [(100.0, 1.0)], 1.0, want_terminal=True)
              for o, h in ((0, n // 2), (n // 2, n - n // 2))]
    np.testing.assert_array_equal(np.concatenate([t for _, t in halves]), tb)
    assert abs(sum(r[0].sum for r, _ in halves) - big[0].sum) <= 1e-12 * big[0].sum


def test_maximum_strike_grid(cuda, oracle):
    m = heston_model()
    strikes = np.linspace(40.0, 200.0, 256)
    pay = [(float(k), 1.0 if i % 2 else -1.0) for i, k in enumerate(strikes)]
    sim = SimSpec(n_paths=3001, n_steps=9, vr=abi.HH_VR_ANTITHETIC, base_seed=4)
    rg, _ = cuda.mc_european(m, sim, pay, 0.9)
    ro, _ = oracle.mc_european(m, sim, pay, 0.9)
    assert rel_err([r.sum + 1.0 for r in rg], [r.sum + 1.0 for r in ro]) < 1e-11
    with pytest.raises(ValueError):
        cuda.mc_european(m, sim, pay + [(100.0, 1.0)], 0.9)   # 257 payoffs
    with pytest.raises(ValueError):
        cuda.mc_european(m, sim, [], 0.9)                      # empty payoff list


def test_degenerate_models(cuda, oracle):
    # zero vol of vol and zero correlation: the variance is deterministic
    m = heston_model(xi=0.0, rho=0.0)
    sim = SimSpec(n_paths=500, n_steps=12, base_seed=1)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    assert rel_err(tg, to) < 1e-11
    # variance started at zero with a violated Feller condition: full truncation keeps everything finite
    m = heston_model(V0=0.0, kappa=0.5, theta=0.01, xi=1.0, rho=-0.9)
    sim = SimSpec(n_paths=4000, n_steps=50, base_seed=1)
    rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    assert np.all(np.isfinite(tg)) and rg[0].n_nonfinite == 0
    assert rel_err(tg, to) < 1e-10
    # zero volatility GBM: every path is the forward
    g = gbm_model(sigma=0.0)
    rg, tg = cuda.mc_european(g, SimSpec(n_paths=100, n_steps=4, scheme=abi.HH_SCHEME_EM, base_seed=1), [(100.0, 1.0)], 1.0,
                              want_terminal=True)
    assert rel_err(tg, np.full(100, 100.0 * math.exp(0.05))) < 1e-14


@pytest.mark.parametrize("n,steps,deg", [(2, 2, 1), (3, 1, 2), (513, 3, 0), (1000, 4, 8)])
def test_lsm_small_and_odd_shapes(cuda, oracle, n, steps, deg):
    m = gbm_model()
    z = np.random.Generator(np.random.Philox(n)).standard_normal((n, steps, 1))
    sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_STEPS, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    D = math.exp(-m.r * m.T / steps)
    og, tg, vg, pg = cuda.lsm_american(m, sim, (105.0, -1.0), deg, D, want_stopping=True, want_paths=True)
    oo, to, vo, po = oracle.lsm_american(m, sim, (105.0, -1.0), min(deg, 5), D, want_stopping=True, want_paths=True)
    assert rel_err(pg, po) < 1e-12
    assert og.n == n and np.all((tg >= 1) & (tg <= steps))
    if deg <= 2 and n > 100:
        assert abs(og.price - oo.price) < 1e-6 * oo.price
    with pytest.raises(ValueError):
        cuda.lsm_american(m, sim, (105.0, -1.0), 9, D)     # degree above the built maximum
    with pytest.raises(NotImplementedError):                # the terminal-law sampler saves no dates
        cuda.lsm_american(m, SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EXACT_TERMINAL), (105.0, -1.0), 2, D)


def test_global_trajectory_index_beyond_32_bits(cuda, oracle):
    """Shards deep into a huge job: the trajectory index fills both counter words of Philox (idx >> 32 != 0)."""
    m = heston_model()
    for prec, tol in ((abi.HH_PREC_F64, 1e-11), (abi.HH_PREC_F32, 5e-3)):
        sim = SimSpec(n_paths=700, n_steps=9, precision=prec, base_seed=77, path_offset=(1 << 33) + 12345)
        rg, tg = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
        ro, to = oracle.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
        assert rel_err(tg, to) < tol
        near = SimSpec(n_paths=700, n_steps=9, precision=prec, base_seed=77, path_offset=12345)
        _, t0 = cuda.mc_european(m, near, [(100.0, 1.0)], 1.0, want_terminal=True)
        assert not np.allclose(t0, tg)  # a different part of the stream
    g = gbm_model()
    sim = SimSpec(n_paths=300, n_steps=6, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=5, path_offset=(1 << 40) + 3)
    og, _, _, pg = cuda.lsm_american(g, sim, (100.0, -1.0), 2, 0.99, want_paths=True)
    oo, _, _, po = oracle.lsm_american(g, sim, (100.0, -1.0), 2, 0.99, want_paths=True)
    assert rel_err(pg, po) < 1e-12


@pytest.mark.parametrize("prec", [abi.HH_PREC_F64, abi.HH_PREC_F32])
def test_segmented_launch_for_the_ensemble_is_invisible(cuda, prec):
    """A large job that returns the terminal vector is launched in segments (the copy of one overlaps the simulation of the
    next): same trajectories as two explicit shards, sums equal up to summation order."""
    m = heston_model()
    n = (8 << 20) + 4321  # >= 2 x the segment minimum
    sim = SimSpec(n_paths=n, n_steps=5, precision=prec, base_seed=21)
    res, term = cuda.mc_european(m, sim, [(100.0, 1.0), (90.0, -1.0)], 1.0, want_terminal=True)
    res0, _ = cuda.mc_european(m, sim, [(100.0, 1.0), (90.0, -1.0)], 1.0)               # one launch, no terminal vector
    parts = [cuda.mc_european(m, SimSpec(n_paths=h, n_steps=5, precision=prec, base_seed=21, path_offset=o), [(100.0, 1.0)], 1.0,
                              want_terminal=True)[1] for o, h in ((0, 3_000_000), (3_000_000, n - 3_000_000))]
    np.testing.assert_array_equal(np.concatenate(parts), term)
    for a, b in zip(res, res0):
        assert a.n == b.n == n
        assert abs(a.sum - b.sum) <= 1e-12 * abs(b.sum) and abs(a.sumsq - b.sumsq) <= 1e-12 * abs(b.sumsq)
    pay = np.maximum(term - 100.0, 0.0)
    assert abs(res[0].sum - pay.sum()) <= 1e-10 * pay.sum()


def test_buffer_lengths_are_checked(cuda):
    """hh_sim carries the element counts behind `seeds` and `normals`; a short buffer is HH_ERR_ARG, never an
    out-of-bounds host read (wrong shapes are easy to produce: exact schemes force one step, Heston needs 2 components)."""
    m = heston_model()
    z = np.zeros((100, 8, 1))                      # Heston needs [path, step, 2]
    with pytest.raises(ValueError, match="normals buffer too short"):
        cuda.mc_european(m, SimSpec(n_paths=100, n_steps=8, rng_mode=abi.HH_RNG_NORMALS, normals=z), [(100.0, 1.0)], 1.0)
    with pytest.raises(ValueError, match="normals buffer too short"):
        cuda.lsm_american(gbm_model(), SimSpec(n_paths=100, n_steps=9, scheme=abi.HH_SCHEME_EXACT_STEPS,
                                               rng_mode=abi.HH_RNG_NORMALS, normals=z), (100.0, -1.0), 2, 0.99)
    # the C side checks seeds_len itself (the Python mirror also raises before the call: bypass it with a raw struct)
    import ctypes as C
    s, keep = SimSpec(n_paths=50, n_steps=4, seeds=np.arange(50, dtype=np.uint64)).to_c(cuda.lib)
    s.n_paths = 60
    res = (abi.hh_result * 1)()
    pa = (abi.hh_payoff * 1)()
    pa[0].strike, pa[0].cp = 100.0, 1.0
    rc = cuda.lib.hh_mc_european(cuda.h, C.byref(m), C.byref(s), pa, 1, 1.0, res, None, 0)
    assert rc == abi.HH_ERR_ARG and b"Number of seeds (50)" in cuda.lib.hh_last_error(cuda.h)


def test_two_contexts_and_two_threads_in_one_process(cuda, oracle):
    """(i) A second context (on a second GPU when the box has one) launches every kernel family that opts in to more than
    48 KB of dynamic shared memory: the opt-in is tracked per device, not per process. (ii) Two threads sharing ONE
    context: hh_mc_european holds the context mutex across launch and collect, so each thread gets its own results."""
    import threading
    import torch
    dev = 1 if torch.cuda.device_count() > 1 else 0
    other = hh.CudaEngine(dev)
    try:
        m, g = heston_model(), gbm_model()
        sim = SimSpec(n_paths=5000, n_steps=12, base_seed=9)
        ra, ta = cuda.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
        rb, tb = other.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
        np.testing.assert_array_equal(ta, tb)
        t = abi.hh_tangent()
        t.dV0 = 1.0
        sa, _ = cuda.tangent_sums(m, [t], sim, [(100.0, 1.0)])
        sb, _ = other.tangent_sums(m, [t], sim, [(100.0, 1.0)])
        np.testing.assert_array_equal(sa, sb)
        lsim = SimSpec(n_paths=4000, n_steps=10, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=3)
        la = cuda.lsm_american(g, lsim, (100.0, -1.0), 3, 0.995)[0]
        lb = other.lsm_american(g, lsim, (100.0, -1.0), 3, 0.995)[0]
        assert la.price == lb.price
        pd = [(abi.HH_PD_ASIAN_ARITH, 100.0, 1.0, 0.0, 0.0)]
        pa, _ = cuda.mc_path_dependent(m, sim, pd, 1.0)
        pb, _ = other.mc_path_dependent(m, sim, pd, 1.0)
        assert pa[0].sum == pb[0].sum
    finally:
        other.close()

    expect = {}
    for seed in range(8):
        expect[seed] = cuda.mc_european(heston_model(), SimSpec(n_paths=20_000, n_steps=20, base_seed=seed), [(100.0, 1.0)], 1.0)[0][0].sum
    got, errs = {}, []

    def work(seeds):
        try:
            for _ in range(5):
                for seed in seeds:
                    r = cuda.mc_european(heston_model(), SimSpec(n_paths=20_000, n_steps=20, base_seed=seed), [(100.0, 1.0)], 1.0)
                    if r[0][0].sum != expect[seed]:
                        errs.append((seed, r[0][0].sum))
                    got[seed] = r[0][0].sum
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=work, args=(range(0, 4),)), threading.Thread(target=work, args=(range(4, 8),))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs[:3]
    assert got == expect


def test_non_finite_parameters_are_argument_errors(cuda):
    """NaN does not always surface as a NaN price (the full-truncation max(v, 0) swallows it: a NaN kappa once priced a call
    at 2.62 with n_nonfinite = 0): non-finite model parameters are rejected at the boundary, on every entry point
    (tools/bad_input_probe.py runs the longer list: nothing crashes, nothing fails to return)."""
    nan, inf = float("nan"), float("inf")
    sim = SimSpec(n_paths=1000, n_steps=4, base_seed=1)
    for kw in (dict(kappa=nan), dict(xi=nan), dict(V0=inf), dict(theta=nan), dict(r=nan), dict(S0=inf), dict(T=inf)):
        m = heston_model(**kw)
        with pytest.raises(ValueError):
            cuda.mc_european(m, sim, [(100.0, 1.0)], 0.97)
        with pytest.raises(ValueError):
            cuda.lsm_american(m, sim, (100.0, -1.0), 2, 0.99)
        with pytest.raises(ValueError):
            cuda.mc_path_dependent(m, sim, [(abi.HH_PD_ASIAN_ARITH, 100.0, 1.0, 0.0, 0.0)], 0.97, 1)
        with pytest.raises(ValueError):
            cuda.tangent_sums(m, [abi.hh_tangent(dS0=1.0)], sim, [(100.0, 1.0)])
    for kw in (dict(sigma=nan), dict(sigma=inf), dict(r=inf)):
        with pytest.raises(ValueError):
            cuda.mc_european(gbm_model(**kw), SimSpec(n_paths=1000, n_steps=4, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1), [(100.0, 1.0)], 0.97)
