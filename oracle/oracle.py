"""ctypes binding of the CPU ORACLE (oracle/hh_oracle.c) — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
`OracleEngine` has the same methods as hedgehog.jl_b200.engine.CudaEngine so a test can run the host
layer's solve() against either and compare.

This module shares NO code with the product: its ctypes structures are declared here from include/hedgehog_mc.h (the
interface the oracle's C file is compiled against), it never imports hedgehog.jl_b200 and never loads
libhedgehog_mc.so — a checker that marshalled through the product's layer would have a common-mode hole, and the
reference arm of bench.py must not show the product library among its loaded objects. Inputs are taken by duck typing:
any object with the fields of hh_model / hh_tangent / SimSpec (the product's or this module's) is copied field by field.

PARITY STATUS: "parity unpinned" per path (the reference is pure Julia, cannot run here, and holds no
golden vectors); pinned against the reference's deterministic known answers and statistical tests —
see oracle/hh_oracle.h. tools/dump_reference.jl + tests/test_reference_golden.py pin it per path wherever Julia runs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# ---- include/hedgehog_mc.h, restated for ctypes (checked against gcc's layout by tests/test_abi.py) ------------------
HH_OK, HH_ERR_ARG, HH_ERR_UNSUPPORTED = 0, -1, -2
HH_MODEL_GBM, HH_MODEL_HESTON = 0, 1
HH_SCHEME_EM, HH_SCHEME_EXACT_TERMINAL, HH_SCHEME_EXACT_STEPS, HH_SCHEME_HESTON_BK = 0, 1, 2, 3
HH_VR_NONE, HH_VR_ANTITHETIC, HH_VR_QUASI_RANDOM = 0, 1, 2
HH_PREC_F64, HH_PREC_F32 = 0, 1
HH_RNG_PHILOX, HH_RNG_NORMALS, HH_RNG_PHILOX_64 = 0, 1, 2
HH_FLAG_SPLIT_STEP, HH_FLAG_Q1_SQRT_MEAN = 1, 2
HH_PD_NSTATS = 5
_d = C.c_double


class o_model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_uint32)] + [(k, _d) for k in (
        "S0", "r", "T", "sigma", "V0", "kappa", "theta", "xi", "rho", "m11", "m12", "m21", "m22")]


class o_bk_config(C.Structure):
    _fields_ = [("n_std", C.c_int32), ("maxiter_newton", C.c_int32), ("maxiter_bisection", C.c_int32),
                ("max_terms", C.c_int32), ("h_fd", _d), ("cf_tol", _d), ("atol", _d)]


class o_sim(C.Structure):
    _fields_ = [("n_paths", C.c_int64), ("path_offset", C.c_int64), ("n_steps", C.c_int32), ("scheme", C.c_int32),
                ("vr", C.c_int32), ("precision", C.c_int32), ("rng_mode", C.c_int32), ("reserved", C.c_int32),
                ("base_seed", C.c_uint64), ("seeds", C.POINTER(C.c_uint64)), ("normals", C.POINTER(_d)),
                ("seeds_len", C.c_uint64), ("normals_len", C.c_uint64), ("bk", o_bk_config)]


class o_payoff(C.Structure):
    _fields_ = [("strike", _d), ("cp", _d)]


class o_path_payoff(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("strike", _d), ("cp", _d), ("barrier", _d), ("amount", _d)]


class o_result(C.Structure):
    _fields_ = [("sum", _d), ("sumsq", _d), ("n", C.c_int64), ("price", _d), ("std_error", _d),
                ("n_nonfinite", C.c_int64), ("n_fallback", C.c_int64), ("kernel_ms", _d)]


class o_tangent(C.Structure):
    _fields_ = [(k, _d) for k in ("dS0", "dr", "dsigma", "dV0", "dkappa", "dtheta", "dxi", "dm11", "dm12", "dm21", "dm22",
                                  "ddiscount")]


class o_lsm_result(C.Structure):
    _fields_ = [("sum", _d), ("sumsq", _d), ("n", C.c_int64), ("price", _d), ("std_error", _d),
                ("n_dates_skipped", C.c_int64), ("kernel_ms", _d), ("path_ms", _d), ("regress_ms", _d)]


@dataclass
class OSim:
    """SimulationConfig + execution knobs, the oracle's own mirror of hh_sim (same field names as the product's SimSpec)."""
    n_paths: int
    n_steps: int = 1
    scheme: int = HH_SCHEME_EM
    vr: int = HH_VR_NONE
    precision: int = HH_PREC_F64
    rng_mode: int = HH_RNG_PHILOX
    base_seed: int = 0
    path_offset: int = 0
    seeds: Optional[np.ndarray] = None
    normals: Optional[np.ndarray] = None


def _copy_fields(dst, src):
    for name, _ in dst._fields_:
        setattr(dst, name, getattr(src, name))
    return dst


def _model_c(model):
    return _copy_fields(o_model(), model)


def _sim_c(sim):
    """hh_sim from any object with SimSpec's fields. Returns (struct, keep-alive list)."""
    s = o_sim()
    s.n_paths, s.path_offset = int(sim.n_paths), int(sim.path_offset)
    s.n_steps, s.scheme, s.vr = int(sim.n_steps), int(sim.scheme), int(sim.vr)
    s.precision, s.rng_mode = int(sim.precision), int(sim.rng_mode)
    s.base_seed = int(sim.base_seed) & 0xFFFFFFFFFFFFFFFF
    keep = []
    if getattr(sim, "seeds", None) is not None:
        seeds = np.ascontiguousarray(sim.seeds, dtype=np.uint64)
        if seeds.shape[0] < sim.n_paths:  # montecarlo.jl:65-66
            raise ValueError(f"Number of seeds ({seeds.shape[0]}) must be ≥ number of trajectories ({sim.n_paths}).")
        keep.append(seeds)
        s.seeds = seeds.ctypes.data_as(C.POINTER(C.c_uint64))
        s.seeds_len = int(seeds.size)
    if getattr(sim, "normals", None) is not None:
        z = np.ascontiguousarray(sim.normals, dtype=np.float64)
        keep.append(z)
        s.normals = z.ctypes.data_as(C.POINTER(_d))
        s.normals_len = int(z.size)
    bk = getattr(sim, "bk", None)
    if bk is not None:
        _copy_fields(s.bk, bk)
    else:  # the reference's keyword defaults (sample_from_cf.jl:27,50,75,110-112); unused by the C oracle
        s.bk.n_std, s.bk.maxiter_newton, s.bk.maxiter_bisection, s.bk.max_terms = 5, 10, 100, 4096
        s.bk.h_fd, s.bk.cf_tol, s.bk.atol = 1e-2, 1e-3, 1e-4
    return s, keep


def _payoff_array(payoffs):
    arr = (o_payoff * len(payoffs))()
    for i, (k, cp) in enumerate(payoffs):
        arr[i].strike, arr[i].cp = float(k), float(cp)
    return arr


def _path_payoff_array(payoffs):
    arr = (o_path_payoff * len(payoffs))()
    for i, (kind, strike, cp, barrier, amount) in enumerate(payoffs):
        arr[i].kind, arr[i].strike, arr[i].cp, arr[i].barrier, arr[i].amount = int(kind), float(strike), float(cp), float(barrier), float(amount)
    return arr


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(_d))


def corr_factor(rho: float, mode: str = "cholesky"):
    """M with M M^T = [1 rho; rho 1] (heston.jl:18-20); the oracle's own copy for callers that must not import the product."""
    import math
    if mode == "cholesky":
        return (1.0, 0.0, rho, math.sqrt(1 - rho * rho))
    if mode == "sym_sqrt":
        p, m = math.sqrt(1 + rho), math.sqrt(1 - rho)
        return ((p + m) / 2, (p - m) / 2, (p - m) / 2, (p + m) / 2)
    raise ValueError(mode)


def heston_model(S0, r, T, V0, kappa, theta, xi, rho, split=True, corr="cholesky"):
    m = o_model()
    m.kind, m.flags = HH_MODEL_HESTON, (HH_FLAG_SPLIT_STEP if split else 0)
    m.S0, m.r, m.T, m.V0, m.kappa, m.theta, m.xi, m.rho = S0, r, T, V0, kappa, theta, xi, rho
    m.m11, m.m12, m.m21, m.m22 = corr_factor(rho, corr)
    return m


LIB = os.path.join(_HERE, "_build", "libhh_oracle.so")
_lib = None


def build(force: bool = False):
    srcs = [os.path.join(_HERE, "hh_oracle.c"), os.path.join(_HERE, "hh_oracle.h"),
            os.path.join(os.path.dirname(_HERE), "include", "hedgehog_mc.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
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
        L.hho_normal_pair64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, dp, dp]
        L.hho_fill_normals.argtypes = [C.POINTER(o_model), C.POINTER(o_sim), dp]
        L.hho_threads.restype = C.c_int
        L.hho_threads_used.restype = C.c_int
        L.hho_set_threads.argtypes = [C.c_int]
        L.hho_mc_european.argtypes = [C.POINTER(o_model), C.POINTER(o_sim), C.POINTER(o_payoff), C.c_int,
                                      C.c_double, C.POINTER(o_result), dp, C.c_size_t]
        L.hho_mc_path_dependent.argtypes = [C.POINTER(o_model), C.POINTER(o_sim), C.c_int,
                                            C.POINTER(o_path_payoff), C.c_int, C.c_double, C.POINTER(o_result), dp]
        L.hho_heston_em_terminal_v.argtypes = [C.POINTER(o_model), C.POINTER(o_sim), dp]
        L.hho_mc_european_tangent_sums.argtypes = [C.POINTER(o_model), C.POINTER(o_tangent), C.c_int,
                                                   C.POINTER(o_sim), C.POINTER(o_payoff), C.c_int, dp]
        L.hho_mc_european_second_sums.argtypes = [C.POINTER(o_model), C.POINTER(o_sim), C.POINTER(o_payoff), C.c_int,
                                                  C.c_double, dp]
        L.hho_lsm_american.argtypes = [C.POINTER(o_model), C.POINTER(o_sim), C.POINTER(o_payoff), C.c_int,
                                       C.c_double, C.POINTER(o_lsm_result), C.POINTER(C.c_int32), dp, dp, dp]
        L.hho_lsm_backward.argtypes = [dp, C.c_int64, C.c_int, C.POINTER(o_payoff), C.c_int, C.c_double,
                                       C.POINTER(o_lsm_result), C.POINTER(C.c_int32), dp, dp]
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


def normal_pair64(key, idx, step):
    a, b = C.c_double(), C.c_double()
    lib().hho_normal_pair64(key, idx, step, C.byref(a), C.byref(b))
    return a.value, b.value


def quasi_random_normals(base_seed: int, path_offset: int, n: int) -> np.ndarray:
    """Restatement of HH_VR_QUASI_RANDOM (include/hedgehog_mc.h): Z_g = Phi^-1(u_g), u_g the base-2 van der Corput point of the
    global trajectory index g under the splitmix64(base_seed) rotation, at the midpoints of the 2^-53 grid. The oracle's
    exact-terminal sampler consumes them in parity mode (rng_mode = HH_RNG_NORMALS)."""
    from scipy.special import ndtri
    M = (1 << 64) - 1
    z = (int(base_seed) + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    shift = z ^ (z >> 31)
    g = np.arange(path_offset, path_offset + n, dtype=np.uint64)
    r = np.zeros(n, dtype=np.uint64)
    for b in range(64):   # 64-bit bit reversal
        r |= ((g >> np.uint64(b)) & np.uint64(1)) << np.uint64(63 - b)
    with np.errstate(over="ignore"):
        pt = r + np.uint64(shift)   # wraps modulo 2^64
    u = ((pt >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    return ndtri(u)


def _raise(rc, what):
    if rc == HH_ERR_ARG:
        raise ValueError(f"oracle {what}: bad argument")
    if rc == HH_ERR_UNSUPPORTED:
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

    @property
    def threads_used(self):
        """OpenMP threads a parallel region really gets (omp_get_num_threads inside one)."""
        return self.lib.hho_threads_used()

    def fill_normals(self, model, sim):
        s, keep = _sim_c(sim)
        ncomp = 2 if model.kind == HH_MODEL_HESTON else 1
        nsteps = 1 if sim.scheme == HH_SCHEME_EXACT_TERMINAL else sim.n_steps
        z = np.empty((sim.n_paths, nsteps, ncomp))
        self.lib.hho_fill_normals(C.byref(_model_c(model)), C.byref(s), _dp(z))
        return z

    def mc_european(self, model, sim, payoffs, discount, want_terminal=False):
        s, keep = _sim_c(sim)
        pa = _payoff_array(payoffs)
        res = (o_result * len(payoffs))()
        terminal, tptr, tlen = None, None, 0
        if want_terminal:
            tlen = sim.n_paths * (2 if sim.vr == HH_VR_ANTITHETIC else 1)
            terminal = np.empty(tlen)
            tptr = _dp(terminal)
        _raise(self.lib.hho_mc_european(C.byref(_model_c(model)), C.byref(s), pa, len(payoffs), float(discount), res, tptr, tlen),
               "mc_european")
        return list(res), terminal

    def mc_path_dependent(self, model, sim, payoffs, discount, monitor_every=1, want_stats=False):
        s, keep = _sim_c(sim)
        pa = _path_payoff_array(payoffs)
        res = (o_result * len(payoffs))()
        ncols = sim.n_paths * (2 if sim.vr == HH_VR_ANTITHETIC else 1)
        stats = np.empty((HH_PD_NSTATS, ncols)) if want_stats else None
        _raise(self.lib.hho_mc_path_dependent(C.byref(_model_c(model)), C.byref(s), int(monitor_every), pa, len(payoffs),
                                              float(discount), res, _dp(stats) if want_stats else None), "mc_path_dependent")
        return list(res), stats

    def heston_terminal_v(self, model, sim):
        s, keep = _sim_c(sim)
        v = np.empty(sim.n_paths * (2 if sim.vr == HH_VR_ANTITHETIC else 1))
        _raise(self.lib.hho_heston_em_terminal_v(C.byref(_model_c(model)), C.byref(s), _dp(v)), "terminal_v")
        return v

    def tangent_sums(self, model, tangents, sim, payoffs, spot_bump: float = 0.0):
        s, keep = _sim_c(sim)
        pa = _payoff_array(payoffs)
        nt = len(tangents)
        ta = (o_tangent * nt)(*[_copy_fields(o_tangent(), t) for t in tangents])
        out = np.zeros((len(payoffs), 2 + 2 * nt))
        _raise(self.lib.hho_mc_european_tangent_sums(C.byref(_model_c(model)), ta, nt, C.byref(s), pa, len(payoffs), _dp(out)),
               "tangent_sums")
        if spot_bump > 0:
            second = np.zeros((len(payoffs), 4))
            _raise(self.lib.hho_mc_european_second_sums(C.byref(_model_c(model)), C.byref(s), pa, len(payoffs),
                                                        float(spot_bump), _dp(second)), "second_sums")
            return out, 0.0, second
        return out, 0.0

    def lsm_american(self, model, sim, payoff, degree, step_discount, want_stopping=False, want_paths=False,
                     comm=None, want_beta=False):
        s, keep = _sim_c(sim)
        pa = _payoff_array([payoff])
        out = o_lsm_result()
        ncols = sim.n_paths * (2 if sim.vr == HH_VR_ANTITHETIC else 1)
        stop_idx = np.empty(ncols, dtype=np.int32) if want_stopping else None
        stop_val = np.empty(ncols) if want_stopping else None
        paths = np.empty((ncols, sim.n_steps + 1)) if want_paths else None
        beta = np.zeros((sim.n_steps + 1, degree + 1)) if want_beta else None
        _raise(self.lib.hho_lsm_american(
            C.byref(_model_c(model)), C.byref(s), pa, int(degree), float(step_discount), C.byref(out),
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
        out = o_lsm_result()
        tau = np.empty(ncols, dtype=np.int32)
        val = np.empty(ncols)
        beta = np.zeros((M + 1, degree + 1))
        _raise(self.lib.hho_lsm_backward(_dp(grid), ncols, M, pa, int(degree), float(step_discount), C.byref(out),
                                         tau.ctypes.data_as(C.POINTER(C.c_int32)), _dp(val), _dp(beta)), "lsm_backward")
        return out, tau, val, beta
