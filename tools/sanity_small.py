#!/usr/bin/env python
"""One small invocation of every kernel family: European f64 / f32 / generic / GBM / tangents (specialised and generic),
Broadie-Kaya, path-dependent payoffs, LSM persistent and with outputs, peer-mailbox loopback. Written for compute-sanitizer;
that tool is closed on the GPU pool (profiles/r2_e_sanitizer.txt), so tests/test_gpu_guards.py runs this script with
HH_DEBUG_GUARDS=1 instead: guard bands around every device buffer, buffers pre-filled with NaN bytes."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model

eng = hh.default_engine(0)
m, g = heston_model(), gbm_model()
pay = [(100.0, 1.0), (95.0, -1.0), (110.0, 1.0)]
for anti in (0, 1):
    for prec in (abi.HH_PREC_F64, abi.HH_PREC_F32):
        r, t = eng.mc_european(m, SimSpec(n_paths=3001, n_steps=7, vr=anti, precision=prec, base_seed=1), pay, 0.97, want_terminal=True)
        assert np.all(np.isfinite(t))
    z = np.random.default_rng(1).standard_normal((1000, 6, 2))
    eng.mc_european(m, SimSpec(n_paths=1000, n_steps=6, vr=anti, rng_mode=abi.HH_RNG_NORMALS, normals=z), pay, 0.97, want_terminal=True)
    for sch, st in ((abi.HH_SCHEME_EM, 9), (abi.HH_SCHEME_EXACT_STEPS, 9), (abi.HH_SCHEME_EXACT_TERMINAL, 1)):
        eng.mc_european(g, SimSpec(n_paths=2049, n_steps=st, scheme=sch, vr=anti, base_seed=2), pay, 0.95, want_terminal=True)
    seeds = np.arange(1, 1501, dtype=np.uint64)
    eng.mc_european(m, SimSpec(n_paths=1500, n_steps=5, vr=anti, seeds=seeds), pay, 0.97)
    tans = [abi.hh_tangent(dS0=1.0), abi.hh_tangent(dV0=1.0), abi.hh_tangent(dkappa=1.0, dtheta=0.5), abi.hh_tangent(dr=1.0, ddiscount=-0.9),
            abi.hh_tangent(dxi=1.0)]
    eng.tangent_sums(m, tans, SimSpec(n_paths=2000, n_steps=11, vr=anti, base_seed=3), pay)
    eng.tangent_sums(m, tans[:2], SimSpec(n_paths=500, n_steps=6, vr=anti, rng_mode=abi.HH_RNG_NORMALS,
                                           normals=np.random.default_rng(2).standard_normal((500, 6, 2))), pay)
    eng.tangent_sums(g, [abi.hh_tangent(dS0=1.0), abi.hh_tangent(dsigma=1.0)], SimSpec(n_paths=900, n_steps=5, vr=anti, base_seed=4), pay)
    out, tau, val, paths = eng.lsm_american(g, SimSpec(n_paths=5001, n_steps=12, scheme=abi.HH_SCHEME_EXACT_STEPS, vr=anti, base_seed=5),
                                            (100.0, -1.0), 3, math.exp(-0.05 / 12), want_stopping=True, want_paths=True)
    assert np.isfinite(out.price)
eng.mc_european(m, SimSpec(n_paths=600, n_steps=3, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=6), pay, 0.97, want_terminal=True)
# the opt-in 64-bit Philox stream, path-dependent payoffs (generic, specialised Heston, Broadie-Kaya dates), log-space and
# Broadie-Kaya LSM generators, second-order sums, the Broadie-Kaya probes
eng.mc_european(m, SimSpec(n_paths=3001, n_steps=8, rng_mode=abi.HH_RNG_PHILOX_64, base_seed=8), pay, 0.97, want_terminal=True)
pd = [(abi.HH_PD_ASIAN_ARITH, 100.0, 1.0, 0.0, 0.0), (abi.HH_PD_ASIAN_GEOM, 100.0, -1.0, 0.0, 0.0),
      (abi.HH_PD_UP_OUT, 100.0, 1.0, 125.0, 0.5), (abi.HH_PD_DOWN_IN, 100.0, -1.0, 80.0, 0.0), (abi.HH_PD_DIGITAL_CASH, 100.0, 1.0, 0.0, 2.0)]
for anti in (0, 1):
    eng.mc_path_dependent(m, SimSpec(n_paths=2001, n_steps=12, vr=anti, base_seed=9), pd, 0.97, monitor_every=3, want_stats=True)
    eng.mc_path_dependent(g, SimSpec(n_paths=2001, n_steps=12, vr=anti, base_seed=10), pd, 0.95, monitor_every=4, want_stats=True)
eng.mc_path_dependent(m, SimSpec(n_paths=700, n_steps=4, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=11), pd, 0.97, monitor_every=2)
eng.mc_path_dependent(m, SimSpec(n_paths=1500, n_steps=10, base_seed=12), [(abi.HH_PD_BS_CONTROL, 100.0, 1.0, 0.0, 0.0),
                                                                         (abi.HH_PD_VANILLA_MINUS_BS, 100.0, 1.0, 0.0, 0.8)], 0.97)
eng.lsm_american(m, SimSpec(n_paths=3000, n_steps=10, scheme=abi.HH_SCHEME_EM, base_seed=13), (100.0, -1.0), 3, 0.997, want_stopping=True)
eng.lsm_american(m, SimSpec(n_paths=900, n_steps=4, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=14), (100.0, -1.0), 2, 0.99, want_paths=True)
eng.tangent_sums(m, tans, SimSpec(n_paths=2000, n_steps=11, base_seed=15), pay, spot_bump=0.5)
eng.bk_chf(m, 0.25, np.full(40, 0.04), np.linspace(0.01, 0.1, 40), np.tile(np.linspace(0.5, 60.0, 30), (40, 1)))
eng.bk_integral(m, 0.25, np.full(200, 0.04), np.linspace(0.005, 0.12, 200), np.linspace(0.01, 0.99, 200))
eng.bk_log_besseli(0.778, np.linspace(0.1, 80, 300) * np.exp(1j * np.linspace(-1.5, 1.5, 300)))
eng.bk_elementary("atan2", np.linspace(-3, 3, 100), np.linspace(3, -3, 100))
h = eng.peer_export()
eng.peer_connect(0, 1, [h])
eng.lsm_american(g, SimSpec(n_paths=2000, n_steps=5, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=7), (100.0, -1.0), 2, 0.99,
                 comm=abi.hh_comm(abi.hh_allreduce_fn(), None, 0, 1))
eng.peer_disconnect()
print("guard violations", eng.debug_check_guards())   # -1 unless HH_DEBUG_GUARDS=1
print("sanity_small ok")
