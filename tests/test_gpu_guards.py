"""Out-of-bounds writes and reads of uninitialised scratch memory, without compute-sanitizer (closed on this GPU pool,
profiles/r2_e_sanitizer.txt): tools/sanity_small.py — one small, ragged invocation of every kernel family — runs in a second
process with HH_DEBUG_GUARDS=1, where every device buffer carries 4 KB guard bands and starts filled with 0xFF bytes
(hh_ctx.h). No guard byte may change, and every result the script checks must stay finite."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env):
    env = dict(os.environ, **extra_env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanity_small.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "sanity_small ok" in p.stdout
    return int(re.search(r"guard violations (-?\d+)", p.stdout).group(1))


def test_no_kernel_writes_outside_its_buffers():
    assert run({"HH_DEBUG_GUARDS": "1"}) == 0


def test_guards_are_off_by_default():
    assert run({"HH_DEBUG_GUARDS": "0"}) == -1


def test_results_repeat_bit_for_bit_across_processes(cuda):
    """Race detector of last resort: the LSM backward kernel (mbarrier ring, warp-specialised producer, grid barriers) and
    the Broadie-Kaya pipeline (atomics in the counting sort) must give identical bits on every run."""
    import math
    import numpy as np
    from hedgehog_jl_b200 import _abi as abi
    from hedgehog_jl_b200.engine import SimSpec
    from helpers import gbm_model, heston_model
    g, m = gbm_model(), heston_model()
    ref = None
    for rep in range(12):
        out, tau, val, _ = cuda.lsm_american(g, SimSpec(n_paths=300_001, n_steps=20, scheme=abi.HH_SCHEME_EXACT_STEPS, base_seed=5,
                                                        vr=rep % 2 * 0), (100.0, -1.0), 3, math.exp(-0.05 / 20), want_stopping=True)
        _, t = cuda.mc_european(m, SimSpec(n_paths=20_001, n_steps=3, scheme=abi.HH_SCHEME_HESTON_BK, base_seed=6), [(100.0, 1.0)], 0.97,
                                want_terminal=True)
        cur = (out.price, out.sum, tau.tobytes(), val.tobytes(), t.tobytes())
        ref = ref or cur
        assert cur == ref, f"run {rep} differs from run 0"
