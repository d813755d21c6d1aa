#!/usr/bin/env python
"""Throughput of the LognormalDynamics kernels (the generic european_kernel template): GBM Euler-Maruyama, exact steps,
exact terminal. usage: python tools/time_gbm.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec

eng = hh.default_engine(0)
m = abi.hh_model()
m.kind, m.flags = abi.HH_MODEL_GBM, abi.HH_FLAG_SPLIT_STEP | abi.HH_FLAG_Q1_SQRT_MEAN
m.S0, m.r, m.T, m.sigma = 100.0, 0.05, 1.0, 0.2
D = math.exp(-0.05)
for name, scheme, n, steps in (("GBM EM", abi.HH_SCHEME_EM, 20_000_000, 252), ("GBM exact steps", abi.HH_SCHEME_EXACT_STEPS, 20_000_000, 252),
                               ("GBM exact terminal", abi.HH_SCHEME_EXACT_TERMINAL, 100_000_000, 1)):
    for anti in (0, 1):
        best = 1e30
        for rep in range(3):
            res, _ = eng.mc_european(m, SimSpec(n_paths=n, n_steps=steps, scheme=scheme, vr=anti, base_seed=rep), [(100.0, 1.0)], D)
            best = min(best, res[0].kernel_ms)
        print(f"{name:20s} anti={anti} paths={n} steps={steps} best_ms={best:.3f} path_steps_per_s={n * steps * (1 + anti) / best * 1e3:.3e} "
              f"price={res[0].price:.4f} se={res[0].std_error:.4f}", flush=True)
