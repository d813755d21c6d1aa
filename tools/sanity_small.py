#!/usr/bin/env python
"""One small invocation of every kernel family (for compute-sanitizer runs): European f64 / f32 / generic / GBM / tangents
(specialised and generic), Broadie-Kaya, LSM persistent and with outputs, peer-mailbox loopback."""
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
h = eng.peer_export()
eng.peer_connect(0, 1, [h])
eng.lsm_american(g, SimSpec(n_paths=2000, n_steps=5, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=7), (100.0, -1.0), 2, 0.99,
                 comm=abi.hh_comm(abi.hh_allreduce_fn(), None, 0, 1))
eng.peer_disconnect()
print("sanity_small ok")
