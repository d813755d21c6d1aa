#!/usr/bin/env python
"""Wall time of the American LSM solve with everything the reference's LSMSolution carries (stopping_info and the
(steps+1) x paths spot matrix: 4.2 GB at config C3) copied to the host. usage: python tools/time_lsm_full.py [paths]"""
import datetime as dt
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedgehog_jl_b200 as hh

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
eng = hh.default_engine(0)
put = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.American(), hh.Put(), hh.Spot())
bs = hh.BlackScholesInputs(dt.date(2020, 1, 1), 0.05, 100.0, 0.2)
for rep in range(3):
    lsm = hh.LSM(hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(n, steps=50, base_seed=rep)), 3)
    for label, kw in (("price only", dict(stopping_info=False)), ("+ stopping_info arrays", dict(stopping_info="arrays")),
                      ("+ spot_paths", dict(stopping_info="arrays", spot_paths=True))):
        t0 = time.perf_counter()
        sol = hh.solve(hh.PricingProblem(put, bs), lsm, engine=eng, **kw)
        t1 = time.perf_counter()
        print(f"rep {rep} {label:24s} wall {1e3 * (t1 - t0):8.1f} ms  kernels {sol.stats['kernel_ms']:.2f} ms  price {sol.price:.5f}", flush=True)
        del sol
