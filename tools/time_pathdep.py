#!/usr/bin/env python
"""Times hh_mc_path_dependent: usage python tools/time_pathdep.py [paths] [steps] [every]
HH_PD_MODEL = heston (default) | gbm; HH_PD_SET = all | logspace (no arithmetic average: the kernel never calls exp) | vanilla | cv (Black-Scholes control)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 252
every = int(sys.argv[3]) if len(sys.argv) > 3 else 1
which = os.environ.get("HH_PD_MODEL", "heston")
eng = hh.default_engine(0)
m = abi.hh_model()
m.kind, m.flags = abi.HH_MODEL_GBM, abi.HH_FLAG_SPLIT_STEP
m.S0, m.r, m.T, m.sigma = 100.0, 0.03, 1.0, 0.2
if which == "heston":
    m.kind = abi.HH_MODEL_HESTON
    m.V0, m.kappa, m.theta, m.xi, m.rho = 0.04, 2.0, 0.04, 0.3, -0.7
    (m.m11, m.m12, m.m21, m.m22), _ = hh.corr_factor(m.rho, "cholesky")
pays = [(abi.HH_PD_ASIAN_GEOM, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_UP_OUT, 100.0, 1.0, 130.0, 0.0),
        (abi.HH_PD_DOWN_IN, 100.0, -1.0, 80.0, 0.0), (abi.HH_PD_DIGITAL_CASH, 100.0, 1.0, 0.0, 1.0)]
which_set = os.environ.get("HH_PD_SET", "all")
if which_set == "all":
    pays.append((abi.HH_PD_ASIAN_ARITH, 100.0, 1.0, 0.0, 0.0))
elif which_set == "vanilla":      # the European payoff through this kernel (baseline of the control-variate cost)
    pays = [(abi.HH_PD_VANILLA, 100.0, 1.0, 0.0, 0.0)]
elif which_set == "cv":           # Black-Scholes control variate (Heston only): the control trajectory is advanced too
    pays = [(abi.HH_PD_VANILLA_MINUS_BS, 100.0, 1.0, 0.0, 0.77)]
for anti in (0, 1):
    best = None
    for rep in range(4):
        sim = SimSpec(n_paths=n, n_steps=steps, scheme=abi.HH_SCHEME_EM, vr=anti, base_seed=100 + rep)
        res, _ = eng.mc_path_dependent(m, sim, pays, math.exp(-m.r * m.T), every)
        if best is None or res[0].kernel_ms < best[0].kernel_ms:
            best = res
    cols = n * (2 if anti else 1)
    print(f"model={which} anti={anti} paths={n} steps={steps} every={every} contracts={len(pays)} kernel_ms={best[0].kernel_ms:.3f} "
          f"column_steps_per_s={cols * steps / (best[0].kernel_ms * 1e-3):.3e} prices=" + " ".join(f"{r.price:.4f}" for r in best))
