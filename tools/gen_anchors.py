#!/usr/bin/env python
"""Closed-form / Fourier / lattice anchors of the five BASELINE configs, written once to tests/golden/config_anchors.json
so that bench.py's GPU arm can print a `check` per config WITHOUT importing oracle/ (only its cpu_baseline leg may).

    python tools/gen_anchors.py          # needs scipy; a minute on one core

Sources: oracle/anchors.py (Black-Scholes black_scholes.jl:38-64, CRR cox_ross_rubinstein.jl:99-141, Carr-Madan
carr_madan.jl:47-92 with the Heston characteristic function heston.jl:307-319, alpha = 1, bound = 32 as in
test/agreement/montecarlo_heston.jl:47). Greeks of the Carr-Madan price are central finite differences with Richardson
extrapolation (two steps), accurate to ~1e-7 relative."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import anchors as A  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "config_anchors.json")
H = dict(S0=100.0, r=0.03, T=1.0, V0=0.04, kappa=2.0, theta=0.04, sigma=0.3, rho=-0.7)   # C2, C4, C5


def hp(K, **kw):
    p = dict(H, **kw)
    return A.heston_price(p["S0"], K, p["r"], p["T"], p["V0"], p["kappa"], p["theta"], p["sigma"], p["rho"])


def d1(K, name, h):
    f = lambda e: (hp(K, **{name: H[name] + e}) - hp(K, **{name: H[name] - e})) / (2 * e)
    return (4 * f(h / 2) - f(h)) / 3


def d2(K, name, h):
    f = lambda e: (hp(K, **{name: H[name] + e}) - 2 * hp(K) + hp(K, **{name: H[name] - e})) / (e * e)
    return (4 * f(h / 2) - f(h)) / 3


def main():
    out = {"generator": "tools/gen_anchors.py", "heston_parameters": H}
    out["c1_black_scholes_call"] = A.bs_price(100.0, 100.0, 0.05, 0.2, 1.0)
    out["c2_carr_madan_call"] = hp(100.0)
    # C3: American put S0 = K = 100, r = 0.05, sigma = 0.2, T = 1; CRR with 1000 steps (the reference's own comparison,
    # test/agreement/american_options.jl:24-26) and the Bermudan value on the 50 exercise dates of the config (CRR with
    # 5000 steps, exercise allowed every 100th step)
    out["c3_crr_american_put_1000"] = A.crr_price(100.0, 100.0, 0.05, 0.2, 1.0, 1000, cp=-1.0, american=True)
    out["c3_crr_bermudan_put_50_dates"] = bermudan_crr(100.0, 100.0, 0.05, 0.2, 1.0, 5000, 50)
    strikes = np.linspace(60.0, 140.0, 64)
    steps = dict(S0=0.5, V0=2e-3, r=2e-3, kappa=2e-2, theta=2e-3, sigma=5e-3, rho=1e-2)
    c5 = {"strikes": strikes.tolist(), "price": [hp(float(K)) for K in strikes]}
    for name, h in steps.items():
        c5["d_" + name] = [d1(float(K), name, h) for K in strikes]
    c5["d2_S0"] = [d2(float(K), "S0", 1.0) for K in strikes]
    out["c5"] = c5
    # "next" rows: closed forms under Black-Scholes r = 0.05, sigma = 0.2, S0 = K = 100, T = 1 (12 monitoring dates),
    # and the Heston European put at the C2 parameters (put-call parity): the American put under Heston must exceed it
    out["next_geometric_asian_call_12_dates"] = A.geometric_asian_price(100.0, 100.0, 0.05, 0.2, 1.0, 12, 1.0)
    out["next_digital_cash_call"] = A.digital_price(100.0, 100.0, 0.05, 0.2, 1.0, 1.0, 1.0)
    out["next_heston_european_put"] = out["c2_carr_madan_call"] - 100.0 + 100.0 * math.exp(-0.03)
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT)


def bermudan_crr(S, K, r, sigma, T, steps, ndates):
    dt = T / steps
    u = math.exp(sigma * math.sqrt(dt))
    p = (math.exp(r * dt) - 1 / u) / (u - 1 / u)
    disc = math.exp(-r * dt)
    every = steps // ndates
    j = np.arange(steps + 1)
    v = np.maximum(K - S * u ** (2.0 * j - steps), 0.0)
    for n in range(steps - 1, -1, -1):
        v = disc * (p * v[1:] + (1 - p) * v[:-1])
        if n % every == 0 and n > 0:
            jj = np.arange(n + 1)
            v = np.maximum(v, K - S * u ** (2.0 * jj - n))
    return float(v[0])


if __name__ == "__main__":
    main()
