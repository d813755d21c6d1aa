#!/usr/bin/env python
"""Wall time of solve() with MonteCarloSolution.ensemble materialised on the host (config C2: 800 MB D2H)."""
import datetime as dt
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedgehog_jl_b200 as hh

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
eng = hh.default_engine(0)
payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
market = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
for rep in range(4):
    m = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(n, steps=252, base_seed=rep), ensemble=True)
    t0 = time.perf_counter()
    sol = hh.solve(hh.PricingProblem(payoff, market), m, engine=eng)
    t1 = time.perf_counter()
    print(f"rep {rep}: wall {1e3 * (t1 - t0):.1f} ms, kernel {sol.stats['kernel_ms']:.1f} ms, ensemble mean {sol.ensemble.mean():.4f}", flush=True)
    del sol
