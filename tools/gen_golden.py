#!/usr/bin/env python
"""Writes tests/golden/mc_golden.npz: small seeded inputs and the CPU oracle's outputs for every scheme of the path.

STATUS: these vectors were produced by oracle/ (the C restatement), NOT by the Julia package, which cannot run in this
image. They pin the oracle and the CUDA kernels against drift and let the GPU tests run without rebuilding the
inputs; the reference's own known answers (Black-Scholes, CRR, Philox KATs) are pinned separately in
tests/test_oracle_anchors.py. Regenerate with:  python tools/gen_golden.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hedgehog_jl_b200 import _abi as abi  # noqa: E402
from hedgehog_jl_b200.engine import SimSpec  # noqa: E402
from helpers import gbm_model, heston_model  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (model kwargs, scheme, n_paths, n_steps, antithetic, payoffs)
    "heston_em": ("heston", dict(), abi.HH_SCHEME_EM, 96, 24, 0, [(100.0, 1.0), (90.0, -1.0)]),
    "heston_em_anti_symsqrt": ("heston", dict(corr="sym_sqrt", rho=-0.5, xi=0.5), abi.HH_SCHEME_EM, 64, 12, 1, [(105.0, 1.0)]),
    "heston_em_nosplit": ("heston", dict(split=False), abi.HH_SCHEME_EM, 64, 12, 0, [(100.0, 1.0)]),
    "gbm_em": ("gbm", dict(), abi.HH_SCHEME_EM, 96, 10, 1, [(100.0, 1.0)]),
    "gbm_exact_terminal": ("gbm", dict(T=366.0 / 365.0), abi.HH_SCHEME_EXACT_TERMINAL, 128, 1, 1, [(100.0, 1.0), (100.0, -1.0)]),
    "gbm_exact_steps": ("gbm", dict(), abi.HH_SCHEME_EXACT_STEPS, 96, 10, 0, [(100.0, -1.0)]),
}


def build(name):
    kind, kw, scheme, n, steps, anti, payoffs = CASES[name]
    m = heston_model(**kw) if kind == "heston" else gbm_model(**kw)
    ncomp = 2 if kind == "heston" else 1
    nst = 1 if scheme == abi.HH_SCHEME_EXACT_TERMINAL else steps
    z = np.random.Generator(np.random.Philox(sum(map(ord, name)))).standard_normal((n, nst, ncomp))
    return m, scheme, n, steps, anti, payoffs, z


def main():
    eng = O.OracleEngine(threads=1)
    out = {}
    for name in CASES:
        m, scheme, n, steps, anti, payoffs, z = build(name)
        D = math.exp(-m.r * m.T)
        # parity mode: caller-supplied normals
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=anti, rng_mode=abi.HH_RNG_NORMALS, normals=z)
        res, term = eng.mc_european(m, sim, payoffs, D, want_terminal=True)
        out[f"{name}/normals"] = z
        out[f"{name}/parity_terminal"] = term
        out[f"{name}/parity_prices"] = np.array([r.price for r in res])
        # native RNG: Philox4x32-10 keyed by base_seed, counter = (trajectory, step)
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=anti, base_seed=2024, path_offset=5)
        res, term = eng.mc_european(m, sim, payoffs, D, want_terminal=True)
        out[f"{name}/native_terminal"] = term
        out[f"{name}/native_prices"] = np.array([r.price for r in res])
    # Longstaff-Schwartz on stored paths (American put, degree 3, 20 dates)
    m = gbm_model()
    z = np.random.Generator(np.random.Philox(77)).standard_normal((512, 20, 1))
    sim = SimSpec(n_paths=512, n_steps=20, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=1, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    o, tau, val, paths = eng.lsm_american(m, sim, (100.0, -1.0), 3, math.exp(-m.r * m.T / 20), want_stopping=True, want_paths=True)
    out["lsm/normals"], out["lsm/tau"], out["lsm/val"], out["lsm/paths"] = z, tau, val, paths
    out["lsm/price"] = np.array([o.price, o.std_error])
    path = os.path.join(ROOT, "tests", "golden", "mc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
