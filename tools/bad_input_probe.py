#!/usr/bin/env python
"""Bad inputs must come back as an error or as counted non-finite results — never as a crash or a kernel that does not
return. Every call runs in this process under the caller's `timeout`; the last line must be reached.
   timeout 300 python tools/bad_input_probe.py"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model

eng = hh.default_engine(0)
nan, inf = float("nan"), float("inf")
EM, XS, XT, BK = abi.HH_SCHEME_EM, abi.HH_SCHEME_EXACT_STEPS, abi.HH_SCHEME_EXACT_TERMINAL, abi.HH_SCHEME_HESTON_BK
cases = []
for name, kw in [("nan kappa", dict(kappa=nan)), ("nan xi", dict(xi=nan)), ("inf V0", dict(V0=inf)), ("negative V0", dict(V0=-0.04)),
                 ("nan S0", dict(S0=nan)), ("zero S0", dict(S0=0.0)), ("negative S0", dict(S0=-5.0)), ("rho 1.5", dict(rho=1.5)),
                 ("huge xi", dict(xi=1e6)), ("tiny xi", dict(xi=1e-12)), ("huge kappa", dict(kappa=1e9)), ("zero T", dict(T=0.0)),
                 ("negative T", dict(T=-1.0)), ("nan T", dict(T=nan)), ("huge T", dict(T=1e6)), ("nan r", dict(r=nan))]:
    for scheme in (EM, BK):
        cases.append(("heston %s / %s" % (name, "EM" if scheme == EM else "BK"), "eur", heston_model, kw, scheme))
    cases.append(("heston %s / LSM" % name, "lsm", heston_model, kw, EM))
    cases.append(("heston %s / path-dependent" % name, "pd", heston_model, kw, EM))
for name, kw in [("nan sigma", dict(sigma=nan)), ("inf sigma", dict(sigma=inf)), ("negative sigma", dict(sigma=-0.2)), ("nan S0", dict(S0=nan)),
                 ("zero T", dict(T=0.0)), ("huge sigma", dict(sigma=1e3))]:
    for scheme in (EM, XS, XT):
        cases.append(("gbm %s / scheme %d" % (name, scheme), "eur", gbm_model, kw, scheme))
    cases.append(("gbm %s / LSM" % name, "lsm", gbm_model, kw, XS))
outcomes = {}
for label, kind, mk, kw, scheme in cases:
    t0 = time.perf_counter()
    try:
        m = mk(**kw)
        steps = 1 if scheme == XT else 6
        sim = SimSpec(n_paths=20000, n_steps=steps, scheme=scheme, base_seed=5)
        if kind == "eur":
            r, _ = eng.mc_european(m, sim, [(100.0, 1.0)], 0.97)
            out = "ran: price %r nonfinite %d" % (r[0].price, r[0].n_nonfinite)
        elif kind == "lsm":
            o = eng.lsm_american(m, sim, (100.0, -1.0), 3, 0.99)[0]
            out = "ran: price %r" % o.price
        else:
            r, _ = eng.mc_path_dependent(m, sim, [(abi.HH_PD_ASIAN_ARITH, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_UP_OUT, 100.0, 1.0, 130.0, 0.0)], 0.97, 1)
            out = "ran: price %r nonfinite %d" % (r[0].price, r[0].n_nonfinite)
    except Exception as e:  # noqa: BLE001
        out = "%s: %s" % (type(e).__name__, str(e)[:90])
    dt = time.perf_counter() - t0
    print("%-44s %6.0f ms  %s" % (label, dt * 1e3, out), flush=True)
    outcomes[label] = dt
for bad in (dict(n_paths=0), dict(n_paths=-5), dict(n_steps=0), dict(n_steps=-1), dict(n_paths=10, n_steps=10 ** 6)):
    try:
        sim = SimSpec(**{**dict(n_paths=1000, n_steps=4, scheme=EM, base_seed=1), **bad})
        r, _ = eng.mc_european(heston_model(), sim, [(100.0, 1.0)], 0.97)
        print("sim %r -> ran, price %r" % (bad, r[0].price), flush=True)
    except Exception as e:  # noqa: BLE001
        print("sim %r -> %s: %s" % (bad, type(e).__name__, str(e)[:100]), flush=True)
r, _ = eng.mc_european(heston_model(), SimSpec(n_paths=100000, n_steps=50, base_seed=1), [(100.0, 1.0)], math.exp(-0.03))
print("context still usable: price %.4f; slowest case %.0f ms" % (r[0].price, 1e3 * max(outcomes.values())))
