#!/usr/bin/env python
"""Times the five BASELINE.json configurations on one GPU through the public API and prints one JSON object
(kernel milliseconds from the library's CUDA events, wall milliseconds around solve()). usage: python tools/time_configs.py [scale]
scale < 1 shrinks every path count (smoke runs)."""
import datetime as dt
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import hedgehog_jl_b200 as hh

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
eng = hh.default_engine(0)
out = {}


def wall(f, reps=3):
    best, res = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = f()
        t = (time.perf_counter() - t0) * 1e3
        best = t if best is None or t < best else best
    return res, best


def want(name):
    return only is None or name in only


call = lambda K=100.0, ex=hh.European(), cp=hh.Call(): hh.VanillaOption(K, dt.date(2020, 12, 31), ex, cp, hh.Spot())
bs = hh.BlackScholesInputs(dt.date(2020, 1, 1), 0.05, 100.0, 0.2)
heston = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)

if want("C1"):
    n = int(1e6 * scale)
    m = hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(n, steps=1, base_seed=42), ensemble=False)
    sol, w = wall(lambda: hh.solve(hh.PricingProblem(call(), bs), m, engine=eng), 5)
    m2 = hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(n, steps=1, base_seed=42), ensemble=True)
    sol2, w2 = wall(lambda: hh.solve(hh.PricingProblem(call(), bs), m2, engine=eng), 5)
    out["C1"] = {"paths": n, "price": sol.price, "se": sol.std_error, "kernel_ms": sol.stats["kernel_ms"], "wall_ms": w,
                 "wall_ms_with_ensemble": w2}

if want("C2"):
    n = int(1e8 * scale)
    for prec in ("f64", "f32"):
        m = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(n, steps=252, base_seed=42), precision=prec,
                          ensemble=False)
        sol, w = wall(lambda: hh.solve(hh.PricingProblem(call(), heston), m, engine=eng), 2)
        out[f"C2_{prec}"] = {"paths": n, "price": sol.price, "se": sol.std_error, "kernel_ms": sol.stats["kernel_ms"], "wall_ms": w,
                             "path_steps_per_s": n * 252 / sol.stats["kernel_ms"] * 1e3}

if want("C3"):
    n = int(1e7 * scale)
    put = call(100.0, hh.American(), hh.Put())
    lsm = hh.LSM(hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(n, steps=50, base_seed=12345)), 3)
    sol, w = wall(lambda: hh.solve(hh.PricingProblem(put, bs), lsm, engine=eng, stopping_info=False), 3)
    out["C3"] = {"paths": n, "price": sol.price, "se": sol.std_error, "kernel_ms": sol.stats["kernel_ms"], "path_ms": sol.stats["path_ms"],
                 "regress_ms": sol.stats["regress_ms"], "wall_ms": w,
                 "algorithmic_GBps": n * 50 * 32 / sol.stats["kernel_ms"] / 1e6}

if want("C4"):
    n = int(1e7 * scale)
    m = hh.MonteCarlo(hh.HestonDynamics(), hh.HestonBroadieKaya(), hh.SimulationConfig(n, steps=12, base_seed=42), ensemble=False,
                      bk_steps_from_config=True)
    sol, w = wall(lambda: hh.solve(hh.PricingProblem(call(), heston), m, engine=eng), 2)
    st = eng.bk_last_stats()
    out["C4"] = {"paths": n, "dates": 12, "price": sol.price, "se": sol.std_error, "kernel_ms": sol.stats["kernel_ms"], "wall_ms": w,
                 "transitions_per_s": n * 12 / sol.stats["kernel_ms"] * 1e3, "bk_stats": st}

if want("C5"):
    n = int(1e7 * scale)
    strikes = np.linspace(60.0, 140.0, 64)
    lenses = [hh.SpotLens(), hh.optic("market_inputs.V0"), hh.ZeroRateSpineLens(1), hh.optic("market_inputs.kappa"),
              hh.optic("market_inputs.theta"), hh.optic("market_inputs.sigma"), hh.optic("market_inputs.rho")]
    m = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(n, steps=252, base_seed=42), ensemble=False)
    prob = hh.PricingProblem(call(), heston)
    (prices, g, se), w = wall(lambda: hh.strike_grid_greeks(prob, strikes, lenses, m, engine=eng), 2)
    k = 32  # strike 100.6
    out["C5"] = {"paths": n, "strikes": 64, "lenses": len(lenses), "wall_ms": w, "path_steps_per_s": n * 252 / w * 1e3,
                 "atm": {"strike": float(strikes[k]), "price": float(prices[k]), "delta": float(g[k, 0]), "dV0": float(g[k, 1]),
                         "rho_rate": float(g[k, 2]), "dkappa": float(g[k, 3]), "dtheta": float(g[k, 4]), "dxi": float(g[k, 5]),
                         "drho": float(g[k, 6]), "delta_se": float(se[k, 0])}}
print(json.dumps(out))
