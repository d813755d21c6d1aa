#!/usr/bin/env python
"""Headline kernel under both Philox streams, with the ablation split (hh_bench_heston_ablation). One GPU.
    python tools/time_streams.py [paths] [steps]   -> JSON on stdout"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 252
eng = hh.default_engine(0)
m = abi.hh_model()
m.kind, m.flags = abi.HH_MODEL_HESTON, abi.HH_FLAG_SPLIT_STEP
m.S0, m.r, m.T, m.V0, m.kappa, m.theta, m.xi, m.rho = 100.0, 0.03, 1.0, 0.04, 2.0, 0.04, 0.3, -0.7
(m.m11, m.m12, m.m21, m.m22), _ = hh.corr_factor(m.rho, "cholesky")
D = math.exp(-0.03)
out = {"paths": n, "steps": steps, "fp64_peak_tflops": eng.fp64_peak()[0]}
clock_hz, smsp = 1.965e9, 148 * 4
for name, rng in (("philox_52bit", abi.HH_RNG_PHILOX), ("philox_64", abi.HH_RNG_PHILOX_64)):
    best, price = 1e30, None
    for rep in range(3):
        r, _ = eng.mc_european(m, SimSpec(n_paths=n, n_steps=steps, rng_mode=rng, base_seed=100 + rep), [(100.0, 1.0)], D)
        best, price = min(best, r[0].kernel_ms), r[0].price
    parts = {}
    for part, label in ((0, "full"), (1, "no_philox"), (2, "philox_only")):
        ms = min(eng.heston_ablation(n, steps, rng, part) for _ in range(2))
        parts[label] = {"ms": ms, "cycles_per_warp_step": ms * 1e-3 * clock_hz / (n * steps / 32 / smsp)}
    out[name] = {"ms": best, "path_steps_per_s": n * steps / (best * 1e-3), "price": price,
                 "cycles_per_warp_step": best * 1e-3 * clock_hz / (n * steps / 32 / smsp), "ablation": parts}
print(json.dumps(out, indent=1))
