#!/usr/bin/env python
"""Price bias of the Broadie-Kaya sampler against Carr-Madan as a function of the inversion settings (n_std: the grid
h = pi / (mean + n_std sd) of the Fourier CDF; cf_tol: the truncation of its series) on a harsh parameter set (degrees
of freedom 0.019, vol of vol 1.04). Separates the algorithm's own approximation (sample_from_cf.jl defaults n = 5,
cf_tol = 1e-3) from the implementation.   python tools/bk_bias_probe.py [n_paths]"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import heston_model

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
C4 = "--c4" in sys.argv   # the benchmark's own parameter set on 12 dates instead of the harsh one
K, r, T = 99.45516304072481, 0.06431298788868413, 412 / 365
pars = dict(V0=0.005603252483435506, kappa=1.2085404839979075, theta=0.004208634722987824, xi=1.0449138820677895,
            rho=-0.29566330506471905)
CM = 7.99928266381862   # oracle/anchors.py heston_price, bound 10000
dates = 1
grid = [(5, 1e-3, 4096), (8, 1e-3, 4096), (12, 1e-3, 4096), (5, 1e-5, 4096), (12, 1e-5, 4096), (20, 1e-6, 4096)]
if C4:
    K, r, T, dates = 100.0, 0.03, 1.0, 12
    pars = dict(V0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7)
    CM = 9.242521073959065   # oracle/anchors.py heston_price, bound 600
    grid = [(5, 1e-3, 4096), (8, 1e-3, 4096), (5, 1e-5, 4096), (8, 1e-5, 4096), (12, 1e-6, 4096)]
eng = hh.default_engine(0)
m = heston_model(S0=100.0, r=r, T=T, **pars)
for n_std, cf_tol, max_terms in grid:
    cfg = abi.hh_bk_config()
    eng.lib.hh_default_bk_config(cfg)
    cfg.n_std, cfg.cf_tol, cfg.max_terms = n_std, cf_tol, max_terms
    sim = SimSpec(n_paths=n, n_steps=dates, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=11, bk=cfg)
    res, _ = eng.mc_european(m, sim, [(K, 1.0)], math.exp(-r * T))
    st = eng.bk_last_stats()
    print(json.dumps({"n_std": n_std, "cf_tol": cf_tol, "price": res[0].price, "std_error": res[0].std_error,
                      "z": (res[0].price - CM) / res[0].std_error, "mean_terms": st["mean_series_terms"],
                      "n_fallback": st["n_fallback"], "kernel_ms": res[0].kernel_ms}), flush=True)
