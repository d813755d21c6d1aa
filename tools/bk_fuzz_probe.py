#!/usr/bin/env python
"""One random Broadie-Kaya case of tests/test_gpu_fuzz.py per process, with a time limit: finds the parameter sets on which
the sampler is slow.   python tools/bk_fuzz_probe.py [n_paths] [seed ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import sys, json, time, datetime as dt
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import hedgehog_jl_b200 as hh
from test_gpu_fuzz import _random_bk_case
seed, n = int(sys.argv[1]), int(sys.argv[2])
pars, days, steps, K = _random_bk_case(seed)
eng = hh.default_engine(0)
ref = dt.date(2020, 1, 1)
prob = hh.PricingProblem(hh.VanillaOption(K, ref + dt.timedelta(days=days), hh.European(), hh.Call(), hh.Spot()),
                         hh.HestonInputs(ref, pars["r"], 100.0, pars["V0"], pars["kappa"], pars["theta"], pars["xi"], pars["rho"]))
mc = hh.MonteCarlo(hh.HestonDynamics(), hh.HestonBroadieKaya(), hh.SimulationConfig(n, steps=steps, base_seed=77 + seed),
                   bk_steps_from_config=steps > 1)
t0 = time.perf_counter()
sol = hh.solve(prob, mc, engine=eng)
print(json.dumps({"seed": seed, "nu": 2 * pars["kappa"] * pars["theta"] / pars["xi"] ** 2 - 1, "days": days, "steps": steps,
                  "wall_s": time.perf_counter() - t0, "kernel_ms": sol.stats["kernel_ms"], "price": sol.price,
                  "stats": eng.bk_last_stats(), "pars": pars}))
""" % (ROOT, os.path.join(ROOT, "tests"))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seeds = [int(a) for a in sys.argv[2:]] or list(range(24))
for s in seeds:
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-c", CODE, str(s), str(n)], capture_output=True, text=True, timeout=40)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else "rc=%d %s" % (p.returncode, p.stderr[-300:])
    except subprocess.TimeoutExpired:
        line = json.dumps({"seed": s, "timeout_s": time.time() - t0})
    print(line, flush=True)
