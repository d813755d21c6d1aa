#!/usr/bin/env python
"""Launches the kernels whose ncu captures feed profiles/ncu_constants.json, at sizes that select the SAME instantiation
bench.py launches. One warm-up of each family, then one measured launch. usage: python tools/ncu_targets.py [which ...]
   which in {c2, c2_64, c2_f32, c3, c4, c5}; prints the work units of each measured launch as JSON (read by tools/ncu_constants.py)."""
import datetime as dt
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import hedgehog_jl_b200 as hh

which = sys.argv[1:] or ["c2", "c2_64", "c2_f32", "c3", "c4", "c5"]
eng = hh.default_engine(0)
call = lambda K=100.0, ex=None, cp=None: hh.VanillaOption(K, dt.date(2020, 12, 31), ex or hh.European(), cp or hh.Call(), hh.Spot())
bs = hh.BlackScholesInputs(dt.date(2020, 1, 1), 0.05, 100.0, 0.2)
heston = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
units = {}
for w in which:
    if w in ("c2", "c2_64"):
        n = 4_000_000  # >= 4 x 148 x 1024: the 1024-thread instantiation of the headline kernel
        m = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(n, steps=252, base_seed=42), ensemble=False,
                          rng="philox64" if w == "c2_64" else "philox")
        for _ in range(2):
            hh.solve(hh.PricingProblem(call(), heston), m, engine=eng)
        units[w] = {"path_steps": n * 252}
    elif w == "c2_f32":
        n = 4_000_000
        m = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(n, steps=252, base_seed=42), ensemble=False,
                          precision="f32")
        for _ in range(2):
            hh.solve(hh.PricingProblem(call(), heston), m, engine=eng)
        units[w] = {"path_steps": n * 252}
    elif w == "c3":
        n = 10_000_000
        lsm = hh.LSM(hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(n, steps=50, base_seed=12345)), 3)
        for _ in range(2):
            hh.solve(hh.PricingProblem(call(100.0, hh.American(), hh.Put()), bs), lsm, engine=eng, stopping_info=False)
        units[w] = {"path_dates": n * 50}
    elif w == "c4":
        n = 2_000_000
        m = hh.MonteCarlo(hh.HestonDynamics(), hh.HestonBroadieKaya(), hh.SimulationConfig(n, steps=12, base_seed=42), ensemble=False,
                          bk_steps_from_config=True)
        for _ in range(2):
            hh.solve(hh.PricingProblem(call(), heston), m, engine=eng)
        units[w] = {"transitions": n * 12}
    elif w == "c5":
        n = 2_000_000
        strikes = np.linspace(60.0, 140.0, 64)
        lenses = [hh.SpotLens(), hh.optic("market_inputs.V0"), hh.ZeroRateSpineLens(1), hh.optic("market_inputs.kappa"),
                  hh.optic("market_inputs.theta"), hh.optic("market_inputs.sigma"), hh.optic("market_inputs.rho")]
        m = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(n, steps=252, base_seed=42), ensemble=False)
        for _ in range(2):
            hh.strike_grid_greeks(hh.PricingProblem(call(), heston), strikes, lenses, m, engine=eng, gamma_bump=0.5)
        units[w] = {"path_steps": n * 252}
print(json.dumps(units))
