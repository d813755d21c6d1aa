#!/usr/bin/env python
"""A job larger than the GPU's memory must come back as an error with a message, and the context must stay usable.
   python tools/oversize_probe.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model

eng = hh.default_engine(0)
for what, call in [
    ("LSM 5e8 x 50 (204 GB grid)", lambda: eng.lsm_american(gbm_model(), SimSpec(n_paths=500_000_000, n_steps=50, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1),
                                                            (100.0, -1.0), 3, 0.999)),
    ("Broadie-Kaya 2e9 x 12", lambda: eng.mc_european(heston_model(), SimSpec(n_paths=2_000_000_000, n_steps=12, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=1),
                                                       [(100.0, 1.0)], 0.97)),
]:
    try:
        out = call()
        print(what, "-> ran:", getattr(out[0], "price", None) if not isinstance(out[0], list) else out[0][0].price)
    except Exception as e:  # noqa: BLE001
        print(what, "->", type(e).__name__, str(e)[:200])
m = gbm_model()
o = eng.lsm_american(m, SimSpec(n_paths=100_000, n_steps=20, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=1), (100.0, -1.0), 3, math.exp(-m.r * m.T / 20))[0]
print("context still usable: LSM 1e5 x 20 price", round(o.price, 4))
