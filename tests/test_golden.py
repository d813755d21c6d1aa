"""Committed golden vectors (tests/golden/mc_golden.npz, made by tools/gen_golden.py from the CPU oracle — the Julia
package cannot run here, so these pin the oracle and the kernels against drift, not against Julia).

CPU tier: the oracle reproduces its own fixtures bit for bit (any change to the restated arithmetic shows up here).
GPU tier: the CUDA path, through the C ABI, matches the fixtures: parity mode rel 1e-12, native RNG rel 1e-11,
LSM stored paths 1e-12 and identical stopping decisions."""
import math
import os
import sys

import numpy as np
import pytest

from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_golden as G  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_golden.npz"))


def _run(engine, name, mode):
    m, scheme, n, steps, anti, payoffs, z = G.build(name)
    np.testing.assert_array_equal(z, GOLD[f"{name}/normals"])  # the generator itself is pinned
    D = math.exp(-m.r * m.T)
    if mode == "parity":
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=anti, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    else:
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=anti, base_seed=2024, path_offset=5)
    res, term = engine.mc_european(m, sim, payoffs, D, want_terminal=True)
    return term, np.array([r.price for r in res])


@pytest.mark.parametrize("mode", ["parity", "native"])
@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(oracle, name, mode):
    term, prices = _run(oracle, name, mode)
    assert rel_err(term, GOLD[f"{name}/{mode}_terminal"]) < 1e-14  # libm may differ in the last bit across hosts
    assert rel_err(prices, GOLD[f"{name}/{mode}_prices"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["parity", "native"])
@pytest.mark.parametrize("name", sorted(G.CASES))
def test_cuda_matches_golden(cuda, name, mode):
    term, prices = _run(cuda, name, mode)
    tol = 1e-12 if mode == "parity" else 1e-11
    assert rel_err(term, GOLD[f"{name}/{mode}_terminal"]) < tol
    assert rel_err(prices, GOLD[f"{name}/{mode}_prices"]) < tol


def _lsm(engine):
    m = gbm_model()
    z = GOLD["lsm/normals"]
    sim = SimSpec(n_paths=512, n_steps=20, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=1, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    return engine.lsm_american(m, sim, (100.0, -1.0), 3, math.exp(-m.r * m.T / 20), want_stopping=True, want_paths=True)


def test_oracle_lsm_reproduces_golden(oracle):
    o, tau, val, paths = _lsm(oracle)
    assert rel_err(paths, GOLD["lsm/paths"]) < 1e-14
    np.testing.assert_array_equal(tau, GOLD["lsm/tau"])
    assert abs(o.price - GOLD["lsm/price"][0]) < 1e-12 * o.price


@pytest.mark.gpu
def test_cuda_lsm_matches_golden(cuda):
    o, tau, val, paths = _lsm(cuda)
    assert rel_err(paths, GOLD["lsm/paths"]) < 1e-12
    flips = int(np.sum(tau != GOLD["lsm/tau"]))
    assert flips <= 1, flips
    if flips == 0:
        assert abs(o.price - GOLD["lsm/price"][0]) < 1e-9 * o.price
