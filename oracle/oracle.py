"""ctypes binding of the CPU ORACLE (oracle/hh_oracle.c) — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
`OracleEngine` has the same methods as hedgehog.jl_b200.engine.CudaEngine so a test can run the host
layer's solve() against either and compare.

PARITY STATUS: "parity unpinned" per path (the reference is pure Julia, cannot run here, and holds no
golden vectors); pinned against the reference's deterministic known answers and statistical tests —
see oracle/hh_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import hedgehog_jl_b200 as hh  # noqa: E402  (struct layouts come from the product's ABI mirror)
from hedgehog_jl_b200 import _abi as abi  # noqa: E402
from hedgehog_jl_b200.engine import SimSpec, _payoff_array, _dp  # noqa: E402

LIB = os.path.join(_HERE, "_build", "libhh_oracle.so")
_lib = None


def build(force: bool = False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "hh_oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        L.hho_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.hho_normal_pair.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, dp, dp]
        L.hho_fill_normals.argtypes = [C.POINTER(abi.hh_model), C.POINTER(abi.hh_sim), dp]
        L.hho_threads.restype = C.c_int
        L.hho_set_threads.argtypes = [C.c_int]
        L.hho_mc_european.argtypes = [C.POINTER(abi.hh_model), C.POINTER(abi.hh_sim), C.POINTER(abi.hh_payoff), C.c_int,
                                      C.c_double, C.POINTER(abi.hh_result), dp, C.c_size_t]
        L.hho_mc_path_dependent.argtypes = [C.POINTER(abi.hh_model), C.POINTER(abi.hh_sim), C.c_int,
                                            C.POINTER(abi.hh_path_payoff), C.c_int, C.c_double, C.POINTER(abi.hh_result), dp]
        L.hho_heston_em_terminal_v.argtypes = [C.POINTER(abi.hh_model), C.POINTER(abi.hh_sim), dp]
        L.hho_mc_european_tangent_sums.argtypes = [C.POINTER(abi.hh_model), C.POINTER(abi.hh_tangent), C.c_int,
                                                   C.POINTER(abi.hh_sim), C.POINTER(abi.hh_payoff), C.c_int, dp]
        L.hho_lsm_american.argtypes = [C.POINTER(abi.hh_model), C.POINTER(abi.hh_sim), C.POINTER(abi.hh_payoff), C.c_int,
                                       C.c_double, C.POINTER(abi.hh_lsm_result), C.POINTER(C.c_int32), dp, dp, dp]
        L.hho_lsm_backward.argtypes = [dp, C.c_int64, C.c_int, C.POINTER(abi.hh_payoff), C.c_int, C.c_double,
                                       C.POINTER(abi.hh_lsm_result), C.POINTER(C.c_int32), dp, dp]
        _lib = L
    return _lib


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().hho_philox4x32_10(c, k, o)
    return list(o)


def normal_pair(key, idx, block, stream=0):
    a, b = C.c_double(), C.c_double()
    lib().hho_normal_pair(key, idx, block, stream, C.byref(a), C.byref(b))
    return a.value, b.value


def _raise(rc, what):
    if rc == abi.HH_ERR_ARG:
        raise ValueError(f"oracle {what}: bad argument")
    if rc == abi.HH_ERR_UNSUPPORTED:
        raise NotImplementedError(f"oracle {what}: unsupported")
    if rc:
        raise RuntimeError(f"oracle {what}: rc={rc}")


class OracleEngine:
    """CPU checker with CudaEngine's interface."""

    name = "oracle"

    def __init__(self, threads: int | None = None):
        self.lib = lib()
        if threads:
            self.lib.hho_set_threads(threads)

    @property
    def threads(self):
        return self.lib.hho_threads()

    def fill_normals(self, model, sim: SimSpec):
        s, keep = sim.to_c(hh.load_library())
        ncomp = 2 if model.kind == abi.HH_MODEL_HESTON else 1
        nsteps = 1 if sim.scheme == abi.HH_SCHEME_EXACT_TERMINAL else sim.n_steps
        z = np.empty((sim.n_paths, nsteps, ncomp))
        self.lib.hho_fill_normals(C.byref(model), C.byref(s), _dp(z))
        return z

    def mc_european(self, model, sim: SimSpec, payoffs, discount, want_terminal=False):
        s, keep = sim.to_c(hh.load_library())
        pa = _payoff_array(payoffs)
        res = (abi.hh_result * len(payoffs))()
        terminal, tptr, tlen = None, None, 0
        if want_terminal:
            tlen = sim.n_paths * (2 if sim.vr == abi.HH_VR_ANTITHETIC else 1)
            terminal = np.empty(tlen)
            tptr = _dp(terminal)
        _raise(self.lib.hho_mc_european(C.byref(model), C.byref(s), pa, len(payoffs), float(discount), res, tptr, tlen),
               "mc_european")
        return list(res), terminal

    def mc_path_dependent(self, model, sim: SimSpec, payoffs, discount, monitor_every=1, want_stats=False):
        from hedgehog_jl_b200.engine import path_payoff_array
        s, keep = sim.to_c(hh.load_library())
        pa = path_payoff_array(payoffs)
        res = (abi.hh_result * len(payoffs))()
        ncols = sim.n_paths * (2 if sim.vr == abi.HH_VR_ANTITHETIC else 1)
        stats = np.empty((abi.HH_PD_NSTATS, ncols)) if want_stats else None
        _raise(self.lib.hho_mc_path_dependent(C.byref(model), C.byref(s), int(monitor_every), pa, len(payoffs),
                                              float(discount), res, _dp(stats) if want_stats else None), "mc_path_dependent")
        return list(res), stats

    def heston_terminal_v(self, model, sim: SimSpec):
        s, keep = sim.to_c(hh.load_library())
        v = np.empty(sim.n_paths * (2 if sim.vr == abi.HH_VR_ANTITHETIC else 1))
        _raise(self.lib.hho_heston_em_terminal_v(C.byref(model), C.byref(s), _dp(v)), "terminal_v")
        return v

    def tangent_sums(self, model, tangents, sim: SimSpec, payoffs):
        s, keep = sim.to_c(hh.load_library())
        pa = _payoff_array(payoffs)
        nt = len(tangents)
        ta = (abi.hh_tangent * nt)(*tangents)
        out = np.zeros((len(payoffs), 2 + 2 * nt))
        _raise(self.lib.hho_mc_european_tangent_sums(C.byref(model), ta, nt, C.byref(s), pa, len(payoffs), _dp(out)),
               "tangent_sums")
        return out, 0.0

    def lsm_american(self, model, sim: SimSpec, payoff, degree, step_discount, want_stopping=False, want_paths=False,
                     comm=None, want_beta=False):
        s, keep = sim.to_c(hh.load_library())
        pa = _payoff_array([payoff])
        out = abi.hh_lsm_result()
        ncols = sim.n_paths * (2 if sim.vr == abi.HH_VR_ANTITHETIC else 1)
        stop_idx = np.empty(ncols, dtype=np.int32) if want_stopping else None
        stop_val = np.empty(ncols) if want_stopping else None
        paths = np.empty((ncols, sim.n_steps + 1)) if want_paths else None
        beta = np.zeros((sim.n_steps + 1, degree + 1)) if want_beta else None
        _raise(self.lib.hho_lsm_american(
            C.byref(model), C.byref(s), pa, int(degree), float(step_discount), C.byref(out),
            stop_idx.ctypes.data_as(C.POINTER(C.c_int32)) if want_stopping else None,
            _dp(stop_val) if want_stopping else None, _dp(paths) if want_paths else None,
            _dp(beta) if want_beta else None), "lsm_american")
        if want_beta:
            return out, stop_idx, stop_val, paths, beta
        return out, stop_idx, stop_val, paths

    def lsm_backward(self, grid, payoff, degree, step_discount):
        """grid: [n_steps+1, ncols] date-major."""
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        M, ncols = grid.shape[0] - 1, grid.shape[1]
        pa = _payoff_array([payoff])
        out = abi.hh_lsm_result()
        tau = np.empty(ncols, dtype=np.int32)
        val = np.empty(ncols)
        beta = np.zeros((M + 1, degree + 1))
        _raise(self.lib.hho_lsm_backward(_dp(grid), ncols, M, pa, int(degree), float(step_discount), C.byref(out),
                                         tau.ctypes.data_as(C.POINTER(C.c_int32)), _dp(val), _dp(beta)), "lsm_backward")
        return out, tau, val, beta
