#!/usr/bin/env python
"""Times the headline Heston EM kernel (kernel_ms from the library's CUDA events) for one tuning variant.
usage: HH_HESTON_VARIANT=k python tools/time_heston.py [paths] [steps] [anti]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 252
anti = int(sys.argv[3]) if len(sys.argv) > 3 else 0
prec = abi.HH_PREC_F32 if os.environ.get('HH_PREC', 'f64') == 'f32' else abi.HH_PREC_F64
eng = hh.default_engine(0)
m = abi.hh_model()
m.kind, m.flags = abi.HH_MODEL_HESTON, abi.HH_FLAG_SPLIT_STEP
m.S0, m.r, m.T = 100.0, 0.03, 1.0
m.V0, m.kappa, m.theta, m.xi, m.rho = 0.04, 2.0, 0.04, 0.3, -0.7
(m.m11, m.m12, m.m21, m.m22), _ = hh.corr_factor(m.rho, "cholesky")
D = math.exp(-0.03)
best = 1e30
for rep in range(5):
    res, _ = eng.mc_european(m, SimSpec(n_paths=n, n_steps=steps, vr=anti, precision=prec, base_seed=42 + rep), [(100.0, 1.0)], D)
    best = min(best, res[0].kernel_ms)
print(f"prec={os.environ.get('HH_PREC', 'f64')} variant={os.environ.get('HH_HESTON_VARIANT', '0')} paths={n} steps={steps} anti={anti} best_ms={best:.3f} "
      f"path_steps_per_s={n * steps * (2 if anti else 1) / best * 1e3:.4e} price={res[0].price:.5f} se={res[0].std_error:.5f}")
