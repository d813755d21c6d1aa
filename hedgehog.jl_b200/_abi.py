"""ctypes mirror of include/hedgehog_mc.h and the loader for libhedgehog_mc.so.

This is the Python stand-in for the Julia `ccall` stubs in julia/HedgehogB200.jl (no Julia
toolchain in this image); the struct layouts and entry points are identical.
The product path has no CPU fallback: if the library or a B200 is missing, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HH_VERSION = 200

HH_OK = 0
HH_ERR_ARG = -1
HH_ERR_UNSUPPORTED = -2
HH_ERR_CUDA = 1
HH_ERR_NOMEM = 2
HH_ERR_COMM = 3

HH_MODEL_GBM, HH_MODEL_HESTON = 0, 1
HH_SCHEME_EM, HH_SCHEME_EXACT_TERMINAL, HH_SCHEME_EXACT_STEPS, HH_SCHEME_HESTON_BK = 0, 1, 2, 3
HH_VR_NONE, HH_VR_ANTITHETIC, HH_VR_QUASI_RANDOM = 0, 1, 2
HH_PREC_F64, HH_PREC_F32 = 0, 1
HH_RNG_PHILOX, HH_RNG_NORMALS, HH_RNG_PHILOX_64 = 0, 1, 2
HH_FLAG_SPLIT_STEP, HH_FLAG_Q1_SQRT_MEAN = 1, 2


class hh_model(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("flags", C.c_uint32),
        ("S0", C.c_double), ("r", C.c_double), ("T", C.c_double), ("sigma", C.c_double),
        ("V0", C.c_double), ("kappa", C.c_double), ("theta", C.c_double), ("xi", C.c_double), ("rho", C.c_double),
        ("m11", C.c_double), ("m12", C.c_double), ("m21", C.c_double), ("m22", C.c_double),
    ]


class hh_bk_config(C.Structure):
    _fields_ = [
        ("n_std", C.c_int32), ("maxiter_newton", C.c_int32), ("maxiter_bisection", C.c_int32),
        ("max_terms", C.c_int32), ("h_fd", C.c_double), ("cf_tol", C.c_double), ("atol", C.c_double),
    ]


class hh_sim(C.Structure):
    _fields_ = [
        ("n_paths", C.c_int64), ("path_offset", C.c_int64),
        ("n_steps", C.c_int32), ("scheme", C.c_int32), ("vr", C.c_int32), ("precision", C.c_int32),
        ("rng_mode", C.c_int32), ("reserved", C.c_int32),
        ("base_seed", C.c_uint64),
        ("seeds", C.POINTER(C.c_uint64)), ("normals", C.POINTER(C.c_double)),
        ("seeds_len", C.c_uint64), ("normals_len", C.c_uint64),
        ("bk", hh_bk_config),
    ]


class hh_payoff(C.Structure):
    _fields_ = [("strike", C.c_double), ("cp", C.c_double)]


# hh_mc_path_dependent (include/hedgehog_mc.h): payoff kinds and per-column statistics
(HH_PD_VANILLA, HH_PD_ASIAN_ARITH, HH_PD_ASIAN_GEOM, HH_PD_UP_OUT, HH_PD_UP_IN, HH_PD_DOWN_OUT, HH_PD_DOWN_IN,
 HH_PD_DIGITAL_CASH, HH_PD_DIGITAL_ASSET, HH_PD_ASIAN_ARITH_MINUS_GEOM, HH_PD_BS_CONTROL, HH_PD_VANILLA_MINUS_BS) = range(12)
HH_PD_NKINDS = 12
HH_PD_NSTATS = 5


class hh_path_payoff(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("strike", C.c_double), ("cp", C.c_double),
                ("barrier", C.c_double), ("amount", C.c_double)]


class hh_result(C.Structure):
    _fields_ = [
        ("sum", C.c_double), ("sumsq", C.c_double), ("n", C.c_int64), ("price", C.c_double),
        ("std_error", C.c_double), ("n_nonfinite", C.c_int64), ("n_fallback", C.c_int64),
        ("kernel_ms", C.c_double),
    ]


class hh_tangent(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "dS0", "dr", "dsigma", "dV0", "dkappa", "dtheta", "dxi", "dm11", "dm12", "dm21", "dm22", "ddiscount")]


class hh_lsm_result(C.Structure):
    _fields_ = [
        ("sum", C.c_double), ("sumsq", C.c_double), ("n", C.c_int64), ("price", C.c_double),
        ("std_error", C.c_double), ("n_dates_skipped", C.c_int64), ("kernel_ms", C.c_double),
        ("path_ms", C.c_double), ("regress_ms", C.c_double),
    ]


hh_allreduce_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class hh_comm(C.Structure):
    _fields_ = [("allreduce_sum_f64", hh_allreduce_fn), ("user", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32)]


_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/hedgehog_mc.h declares
SYMBOLS = {
    "hh_version": (C.c_int, []),
    "hh_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "hh_destroy": (C.c_int, [C.c_void_p]),
    "hh_last_error": (C.c_char_p, [C.c_void_p]),
    "hh_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hh_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_size_t)]),
    "hh_default_bk_config": (None, [C.POINTER(hh_bk_config)]),
    "hh_bench_fp64_peak": (C.c_int, [C.c_void_p, _dp, _dp]),
    "hh_bench_heston_ablation": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, _dp]),
    "hh_mc_european": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.POINTER(hh_sim), C.POINTER(hh_payoff), C.c_int,
                                 C.c_double, C.POINTER(hh_result), _dp, C.c_size_t]),
    "hh_mc_european_launch": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.POINTER(hh_sim), C.POINTER(hh_payoff),
                                        C.c_int, C.c_int]),
    "hh_mc_european_collect": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(hh_result), _dp, C.c_size_t]),
    "hh_mc_european_tangent": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.POINTER(hh_tangent), C.c_int,
                                         C.POINTER(hh_sim), C.POINTER(hh_payoff), C.c_int, C.c_double,
                                         C.POINTER(hh_result), _dp, _dp]),
    "hh_mc_european_tangent_sums": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.POINTER(hh_tangent), C.c_int,
                                              C.POINTER(hh_sim), C.POINTER(hh_payoff), C.c_int, _dp, C.c_double, _dp, _dp]),
    "hh_lsm_american": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.POINTER(hh_sim), C.POINTER(hh_payoff), C.c_int,
                                  C.c_double, C.POINTER(hh_comm), C.POINTER(hh_lsm_result), C.POINTER(C.c_int32),
                                  _dp, _dp]),
    "hh_mc_path_dependent": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.POINTER(hh_sim), C.c_int,
                                       C.POINTER(hh_path_payoff), C.c_int, C.c_double, C.POINTER(hh_result), _dp, C.c_size_t]),
    "hh_bk_chf": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.c_double, _dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp]),
    "hh_bk_log_besseli": (C.c_int, [C.c_void_p, C.c_double, _dp, _dp, C.c_int, _dp, _dp]),
    "hh_bk_elementary": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp]),
    "hh_debug_check_guards": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "hh_bk_integral": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.c_double, C.POINTER(hh_bk_config), _dp, _dp, _dp,
                                 C.c_int, _dp]),
    "hh_bk_variance": (C.c_int, [C.c_void_p, C.POINTER(hh_model), C.c_double, _dp, C.c_int, C.c_uint64, _dp]),
    "hh_bk_last_stats": (C.c_int, [C.c_void_p, _dp]),
    "hh_peer_export": (C.c_int, [C.c_void_p, C.POINTER(C.c_ubyte)]),
    "hh_peer_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]),
    "hh_peer_disconnect": (C.c_int, [C.c_void_p]),
    "hh_peer_set_timeout": (C.c_int, [C.c_void_p, C.c_double]),
}
HH_IPC_HANDLE_BYTES = 64
HH_MAX_PEERS = 16
HH_ERR_PEER_TIMEOUT = 4

# HH_LIB_PATH: another build of the same library (A/B timing of two builds on one box); the default is the in-tree one
LIB_PATH = os.environ.get("HH_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhedgehog_mc.so")
_lib = None


class HedgehogB200Error(RuntimeError):
    """Raised for CUDA / library failures (Julia host: ErrorException)."""


def load_library(path: str | None = None):
    """dlopen libhedgehog_mc.so and bind every declared symbol. Fails loudly if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise HedgehogB200Error(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(hedgehog.jl_b200 has no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
