#!/usr/bin/env python
"""Broadie-Kaya (config C4) on one GPU: timing, inversion statistics and a digest of the terminal spots.

    python tools/time_bk.py [scale] [--ensemble-digest] [--xi=0.1] [--dates=12]   -> JSON on stdout   (--xi: vol of vol, default C4's 0.3)
    HH_LIB_PATH=tools/_build/libhedgehog_mc_prev.so python tools/time_bk.py ...   # another build of the library (A/B)

`--ensemble-digest` also prices 200 000 trajectories with the terminal vector requested and prints its SHA-256, so that
two builds can be compared to the last bit from two processes.
"""
import hashlib
import json
import os
import sys
import time
import datetime as dt

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hedgehog_jl_b200 as hh

args = [a for a in sys.argv[1:] if not a.startswith("--")]
scale = float(args[0]) if args else 1.0
eng = hh.default_engine(0)
payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
xi = float(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("--xi=")), 0.3))
heston = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, xi, -0.7)
prob = hh.PricingProblem(payoff, heston)


DATES = int(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("--dates=")), 12))


def method(n, ensemble=False, dates=DATES):
    return hh.MonteCarlo(hh.HestonDynamics(), hh.HestonBroadieKaya(), hh.SimulationConfig(n, steps=dates, base_seed=42),
                         ensemble=ensemble, bk_steps_from_config=True)


out = {"lib": os.environ.get("HH_LIB_PATH", "in-tree"), "xi": xi, "bessel_order": 2 * 2.0 * 0.04 / xi ** 2 - 1}
n = max(int(1e7 * scale), 1000)
hh.solve(prob, method(n), engine=eng)
best, sol = 1e30, None
for _ in range(3):
    t0 = time.perf_counter()
    sol = hh.solve(prob, method(n), engine=eng)
    best = min(best, sol.stats["kernel_ms"])
    wall = (time.perf_counter() - t0) * 1e3
st = eng.bk_last_stats()
out.update({"paths": n, "dates": DATES, "price": sol.price, "std_error": sol.std_error, "kernel_ms": best, "wall_ms": wall,
            "transitions_per_s": n * DATES / best * 1e3, "cf_evaluations_per_s": n * DATES * (1 + st["mean_series_terms"]) / best * 1e3,
            "bk_stats": st})
if "--ensemble-digest" in sys.argv:
    s2 = hh.solve(prob, method(200_000, ensemble=True), engine=eng)
    ens = np.ascontiguousarray(s2.ensemble, dtype=np.float64)
    out["digest_200k"] = {"sha256": hashlib.sha256(ens.tobytes()).hexdigest(), "price": s2.price, "mean": float(ens.mean())}
    s3 = hh.solve(prob, method(50_000, ensemble=True, dates=1), engine=eng)
    ens = np.ascontiguousarray(s3.ensemble, dtype=np.float64)
    out["digest_50k_one_date"] = {"sha256": hashlib.sha256(ens.tobytes()).hexdigest(), "price": s3.price}
print(json.dumps(out, indent=1))
