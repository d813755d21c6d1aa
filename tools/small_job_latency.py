#!/usr/bin/env python
"""Wall time of small jobs through the C ABI (warm context): the sizes the reference's own tests use (1e3 ... 1e5 trajectories).
   python tools/small_job_latency.py"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model

eng = hh.default_engine(0)
g, h = gbm_model(), heston_model()
jobs = {
    "GBM exact terminal, 1 step": lambda n: eng.mc_european(g, SimSpec(n_paths=n, n_steps=1, scheme=abi.HH_SCHEME_EXACT_TERMINAL, base_seed=1), [(100.0, 1.0)], 0.95),
    "Heston EM, 252 steps": lambda n: eng.mc_european(h, SimSpec(n_paths=n, n_steps=252, base_seed=1), [(100.0, 1.0)], 0.97),
    "Heston EM + terminal vector": lambda n: eng.mc_european(h, SimSpec(n_paths=n, n_steps=252, base_seed=1), [(100.0, 1.0)], 0.97, want_terminal=True),
    "Heston Broadie-Kaya, 1 date": lambda n: eng.mc_european(h, SimSpec(n_paths=n, n_steps=1, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=1), [(100.0, 1.0)], 0.97),
    "Heston tangents (7 directions)": lambda n: eng.tangent_sums(h, [abi.hh_tangent(dS0=1.0)] * 7, SimSpec(n_paths=n, n_steps=252, base_seed=1), [(100.0, 1.0)]),
    "LSM GBM 100 dates degree 5 (+ stopping info)": lambda n: eng.lsm_american(g, SimSpec(n_paths=n, n_steps=100, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=1, base_seed=1),
                                                                               (100.0, -1.0), 5, math.exp(-0.05 / 100), want_stopping=True),
    "path-dependent Heston, 5 contracts": lambda n: eng.mc_path_dependent(h, SimSpec(n_paths=n, n_steps=252, base_seed=1),
                                                                        [(k, 100.0, 1.0, 130.0, 0.0) for k in (abi.HH_PD_ASIAN_ARITH, abi.HH_PD_ASIAN_GEOM, abi.HH_PD_UP_OUT,
                                                                                                                abi.HH_PD_UP_IN, abi.HH_PD_VANILLA)], 0.97, 21),
}
print("%-46s %10s %10s %10s   (ms, best of 20, warm)" % ("job", "n=1e3", "n=1e4", "n=1e5"))
for name, f in jobs.items():
    row = []
    for n in (1000, 10000, 100000):
        f(n)
        best = 1e9
        for _ in range(20):
            t0 = time.perf_counter()
            f(n)
            best = min(best, time.perf_counter() - t0)
        row.append(best * 1e3)
    print("%-46s %10.3f %10.3f %10.3f" % (name, *row), flush=True)
