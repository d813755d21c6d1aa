"""CudaEngine — thin object wrapper over the C ABI (one hh_ctx = one GPU).

Everything numerical happens inside libhedgehog_mc.so; this file only marshals POD structs and
numpy buffers across ctypes, exactly what the Julia host does with `ccall`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _abi as abi


@dataclass
class SimSpec:
    """Python-side mirror of hh_sim (numpy arrays instead of raw pointers)."""
    n_paths: int
    n_steps: int = 1
    scheme: int = abi.HH_SCHEME_EM
    vr: int = abi.HH_VR_NONE
    precision: int = abi.HH_PREC_F64
    rng_mode: int = abi.HH_RNG_PHILOX
    base_seed: int = 0
    path_offset: int = 0
    seeds: Optional[np.ndarray] = None      # uint64[n_paths]
    normals: Optional[np.ndarray] = None    # float64[n_paths, n_steps, ncomp]
    bk: Optional[abi.hh_bk_config] = None
    job_paths: int = 0                      # trajectories of the whole job when this is one shard of it (hh_sim.reserved)

    def to_c(self, lib):
        s = abi.hh_sim()
        s.n_paths, s.path_offset = int(self.n_paths), int(self.path_offset)
        s.n_steps, s.scheme, s.vr = int(self.n_steps), int(self.scheme), int(self.vr)
        s.precision, s.rng_mode = int(self.precision), int(self.rng_mode)
        s.reserved = int(self.job_paths).bit_length() if self.job_paths else 0
        s.base_seed = int(self.base_seed) & 0xFFFFFFFFFFFFFFFF
        keep = []
        if self.seeds is not None:
            seeds = np.ascontiguousarray(self.seeds, dtype=np.uint64)
            if seeds.shape[0] < self.n_paths:  # montecarlo.jl:65-66
                raise ValueError(f"Number of seeds ({seeds.shape[0]}) must be ≥ number of trajectories ({self.n_paths}).")
            keep.append(seeds)
            s.seeds = seeds.ctypes.data_as(C.POINTER(C.c_uint64))
            s.seeds_len = int(seeds.size)
        if self.normals is not None:
            z = np.ascontiguousarray(self.normals, dtype=np.float64)
            # the C side reads n_paths * n_steps * components doubles behind this pointer and checks normals_len
            keep.append(z)
            s.normals = z.ctypes.data_as(C.POINTER(C.c_double))
            s.normals_len = int(z.size)
        if self.bk is not None:
            s.bk = self.bk
        else:
            lib.hh_default_bk_config(C.byref(s.bk))
        return s, keep


def path_payoff_array(payoffs):
    arr = (abi.hh_path_payoff * len(payoffs))()
    for i, (kind, strike, cp, barrier, amount) in enumerate(payoffs):
        arr[i].kind, arr[i].strike, arr[i].cp, arr[i].barrier, arr[i].amount = int(kind), float(strike), float(cp), float(barrier), float(amount)
    return arr


def _payoff_array(payoffs: Sequence[tuple]):
    arr = (abi.hh_payoff * len(payoffs))()
    for i, (k, cp) in enumerate(payoffs):
        arr[i].strike, arr[i].cp = float(k), float(cp)
    return arr


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CudaEngine:
    """One GPU. Raises HedgehogB200Error if the library or a B200 is not available (no CPU fallback)."""

    name = "cuda"

    def __init__(self, device: int = 0):
        self.lib = abi.load_library()
        self.device = device
        h = C.c_void_p()
        rc = self.lib.hh_create(C.byref(h), int(device))
        if rc != abi.HH_OK:
            raise abi.HedgehogB200Error(f"hh_create(device={device}) failed ({rc}): "
                                        f"{self.lib.hh_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.hh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc == abi.HH_OK:
            return
        msg = self.lib.hh_last_error(self.h).decode()
        if rc in (abi.HH_ERR_ARG,):
            raise ValueError(f"{what}: {msg}")  # Julia host: ArgumentError
        if rc == abi.HH_ERR_UNSUPPORTED:
            raise NotImplementedError(f"{what}: {msg}")  # Julia host: MethodError-like
        raise abi.HedgehogB200Error(f"{what} failed ({rc}): {msg}")

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self.lib.hh_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)), "hh_set_stream")

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_size_t()
        self._check(self.lib.hh_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)), "hh_device_info")
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def heston_ablation(self, n_paths: int, n_steps: int, rng_mode: int, part: int) -> float:
        """Device ms of the headline kernel with one part of its step removed (hh_bench_heston_ablation)."""
        ms = C.c_double()
        self._check(self.lib.hh_bench_heston_ablation(self.h, int(n_paths), int(n_steps), int(rng_mode), int(part),
                                                      C.byref(ms)), "hh_bench_heston_ablation")
        return ms.value

    def peer_set_timeout(self, seconds: float):
        self._check(self.lib.hh_peer_set_timeout(self.h, float(seconds)), "hh_peer_set_timeout")

    def fp64_peak(self):
        """(TFLOP/s, ms) of the in-library DFMA-chain microbenchmark."""
        t, ms = C.c_double(), C.c_double()
        self._check(self.lib.hh_bench_fp64_peak(self.h, C.byref(t), C.byref(ms)), "hh_bench_fp64_peak")
        return t.value, ms.value

    # -- European MC ---------------------------------------------------------------------------
    def mc_european(self, model: abi.hh_model, sim: SimSpec, payoffs: Sequence[tuple], discount: float,
                    want_terminal: bool = False):
        s, keep = sim.to_c(self.lib)
        pa = _payoff_array(payoffs)
        res = (abi.hh_result * len(payoffs))()
        terminal = None
        tptr, tlen = None, 0
        if want_terminal:
            tlen = sim.n_paths * (2 if sim.vr == abi.HH_VR_ANTITHETIC else 1)
            terminal = np.empty(tlen, dtype=np.float64)
            tptr = _dp(terminal)
        rc = self.lib.hh_mc_european(self.h, C.byref(model), C.byref(s), pa, len(payoffs), float(discount), res, tptr, tlen)
        self._check(rc, "hh_mc_european")
        del keep
        return list(res), terminal

    def mc_european_launch(self, model, sim: SimSpec, payoffs, want_terminal=False):
        s, keep = sim.to_c(self.lib)
        pa = _payoff_array(payoffs)
        self._pending = (keep, pa, len(payoffs), sim)
        self._check(self.lib.hh_mc_european_launch(self.h, C.byref(model), C.byref(s), pa, len(payoffs),
                                                   1 if want_terminal else 0), "hh_mc_european_launch")

    def mc_european_collect(self, discount: float, terminal_out: Optional[np.ndarray] = None):
        _, _, npay, sim = self._pending
        res = (abi.hh_result * npay)()
        tptr, tlen = None, 0
        if terminal_out is not None:
            tptr, tlen = _dp(terminal_out), terminal_out.size
        self._check(self.lib.hh_mc_european_collect(self.h, float(discount), res, tptr, tlen), "hh_mc_european_collect")
        self._pending = None
        return list(res)

    # -- peer mailboxes (multi-GPU LSM without a collective library) ------------------------------------------------
    peers = None  # (rank, world) once connected
    last_tangent_ms = 0.0  # device ms of the last hh_mc_european_tangent_sums call

    def peer_export(self) -> bytes:
        buf = (C.c_ubyte * abi.HH_IPC_HANDLE_BYTES)()
        self._check(self.lib.hh_peer_export(self.h, buf), "hh_peer_export")
        return bytes(buf)

    def peer_connect(self, rank: int, world: int, handles: Sequence[bytes]):
        blob = b"".join(handles)
        assert len(blob) == world * abi.HH_IPC_HANDLE_BYTES
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self.lib.hh_peer_connect(self.h, int(rank), int(world), arr), "hh_peer_connect")
        self.peers = (int(rank), int(world))

    def peer_disconnect(self):
        self._check(self.lib.hh_peer_disconnect(self.h), "hh_peer_disconnect")
        self.peers = None

    # -- tangents --------------------------------------------------------------------------------
    def tangent_sums(self, model, tangents: Sequence[abi.hh_tangent], sim: SimSpec, payoffs, spot_bump: float = 0.0):
        """Raw sums [npay, 2 + 2*ntan] and kernel ms; with spot_bump > 0 also the second-order sums [npay, 4]
        (sum sd, sum sd^2, sum dd, sum dd^2; include/hedgehog_mc.h) as a third element."""
        s, keep = sim.to_c(self.lib)
        pa = _payoff_array(payoffs)
        nt = len(tangents)
        ta = (abi.hh_tangent * nt)(*tangents)
        out = np.zeros((len(payoffs), 2 + 2 * nt), dtype=np.float64)
        second = np.zeros((len(payoffs), 4), dtype=np.float64) if spot_bump > 0 else None
        ms = C.c_double()
        rc = self.lib.hh_mc_european_tangent_sums(self.h, C.byref(model), ta, nt, C.byref(s), pa, len(payoffs),
                                                  _dp(out), float(spot_bump), _dp(second) if second is not None else None,
                                                  C.byref(ms))
        self._check(rc, "hh_mc_european_tangent_sums")
        del keep
        self.last_tangent_ms = ms.value
        if second is not None:
            return out, ms.value, second
        return out, ms.value

    # -- LSM ---------------------------------------------------------------------------------------
    def lsm_american(self, model, sim: SimSpec, payoff: tuple, degree: int, step_discount: float,
                     want_stopping: bool = False, want_paths: bool = False, comm: Optional[abi.hh_comm] = None):
        s, keep = sim.to_c(self.lib)
        pa = _payoff_array([payoff])
        out = abi.hh_lsm_result()
        ncols = sim.n_paths * (2 if sim.vr == abi.HH_VR_ANTITHETIC else 1)
        stop_idx = stop_val = paths = None
        ip = vp = pp = None
        if want_stopping:
            stop_idx = np.empty(ncols, dtype=np.int32)
            stop_val = np.empty(ncols, dtype=np.float64)
            ip, vp = stop_idx.ctypes.data_as(C.POINTER(C.c_int32)), _dp(stop_val)
        if want_paths:
            # reference layout: Matrix (nsteps+1) x ncols, column-major  == C-order (ncols, nsteps+1)
            paths = np.empty((ncols, sim.n_steps + 1), dtype=np.float64)
            pp = _dp(paths)
        rc = self.lib.hh_lsm_american(self.h, C.byref(model), C.byref(s), pa, int(degree), float(step_discount),
                                      C.byref(comm) if comm is not None else None, C.byref(out), ip, vp, pp)
        self._check(rc, "hh_lsm_american")
        del keep
        return out, stop_idx, stop_val, paths

    # -- path-dependent payoffs (hh_mc_path_dependent) -------------------------------------------------------
    def mc_path_dependent(self, model, sim: SimSpec, payoffs, discount: float, monitor_every: int = 1,
                          want_stats: bool = False):
        """payoffs: sequence of (kind, strike, cp, barrier, amount). Returns (results, stats) with stats the
        HH_PD_NSTATS x ncols array [S_T, A, G, max S, min S] or None."""
        s, keep = sim.to_c(self.lib)
        pa = path_payoff_array(payoffs)
        res = (abi.hh_result * len(payoffs))()
        ncols = sim.n_paths * (2 if sim.vr == abi.HH_VR_ANTITHETIC else 1)
        stats = np.empty((abi.HH_PD_NSTATS, ncols), dtype=np.float64) if want_stats else None
        rc = self.lib.hh_mc_path_dependent(self.h, C.byref(model), C.byref(s), int(monitor_every), pa, len(payoffs),
                                           float(discount), res, _dp(stats) if want_stats else None,
                                           stats.size if want_stats else 0)
        self._check(rc, "hh_mc_path_dependent")
        del keep
        return list(res), stats

    # -- Broadie-Kaya probes ---------------------------------------------------------------------------
    def bk_chf(self, model, tau: float, V0: np.ndarray, VT: np.ndarray, a: np.ndarray):
        V0 = np.ascontiguousarray(V0, dtype=np.float64)
        VT = np.ascontiguousarray(VT, dtype=np.float64)
        a = np.ascontiguousarray(a, dtype=np.float64)
        n, na = V0.shape[0], a.shape[1]
        re = np.empty((n, na)); im = np.empty((n, na))
        self._check(self.lib.hh_bk_chf(self.h, C.byref(model), float(tau), _dp(V0), _dp(VT), n, _dp(a), na, _dp(re), _dp(im)),
                    "hh_bk_chf")
        return re + 1j * im

    def bk_log_besseli(self, nu: float, z: np.ndarray):
        z = np.ascontiguousarray(z, dtype=np.complex128)
        zr, zi = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
        re = np.empty(z.shape[0]); im = np.empty(z.shape[0])
        self._check(self.lib.hh_bk_log_besseli(self.h, float(nu), _dp(zr), _dp(zi), z.shape[0], _dp(re), _dp(im)),
                    "hh_bk_log_besseli")
        return re + 1j * im

    def bk_elementary(self, kind: str, x, y=None):
        """The table-driven exp / sincos / log / atan2 of the Broadie-Kaya kernels (accuracy probes)."""
        k = {"exp": 0, "sincos": 1, "log": 2, "atan2": 3}[kind]
        x = np.ascontiguousarray(x, dtype=np.float64)
        yy = np.ascontiguousarray(y, dtype=np.float64) if y is not None else None
        a = np.empty(x.shape[0]); b = np.empty(x.shape[0])
        self._check(self.lib.hh_bk_elementary(self.h, k, _dp(x), _dp(yy) if yy is not None else None, x.shape[0], _dp(a), _dp(b)),
                    "hh_bk_elementary")
        return (a, b) if k == 1 else a

    def bk_integral(self, model, tau: float, V0, VT, U, cfg: Optional[abi.hh_bk_config] = None):
        """sample_from_cf for given (V0, VT, u) triples -> dict of arrays (x, mean, var, h, J, status, resid, evals)."""
        V0 = np.ascontiguousarray(V0, dtype=np.float64)
        VT = np.ascontiguousarray(VT, dtype=np.float64)
        U = np.ascontiguousarray(U, dtype=np.float64)
        out = np.empty((V0.shape[0], 8))
        self._check(self.lib.hh_bk_integral(self.h, C.byref(model), float(tau), C.byref(cfg) if cfg is not None else None,
                                            _dp(V0), _dp(VT), _dp(U), V0.shape[0], _dp(out)), "hh_bk_integral")
        names = ("x", "mean", "var", "h", "J", "status", "resid", "evals")
        return {k: out[:, i] for i, k in enumerate(names)}

    def bk_variance(self, model, tau: float, V0, seed: int):
        V0 = np.ascontiguousarray(V0, dtype=np.float64)
        VT = np.empty_like(V0)
        self._check(self.lib.hh_bk_variance(self.h, C.byref(model), float(tau), _dp(V0), V0.shape[0],
                                            int(seed) & 0xFFFFFFFFFFFFFFFF, _dp(VT)), "hh_bk_variance")
        return VT

    def debug_check_guards(self) -> int:
        """Guard-band bytes overwritten so far (HH_DEBUG_GUARDS=1), -1 when the guards are off."""
        v = C.c_int64(0)
        self._check(self.lib.hh_debug_check_guards(self.h, C.byref(v)), "hh_debug_check_guards")
        return int(v.value)

    def bk_last_stats(self) -> dict:
        out = np.zeros(5)
        self._check(self.lib.hh_bk_last_stats(self.h, _dp(out)), "hh_bk_last_stats")
        n = max(out[3], 1.0)
        return {"n_fallback": int(out[0]), "mean_series_terms": out[1] / n, "mean_cdf_evaluations": out[2] / n,
                "transitions": int(out[3]), "n_unbracketed": int(out[4])}


_default_engines: dict = {}


def default_engine(device: Optional[int] = None) -> CudaEngine:
    """The process-wide engine for `device` (default: LOCAL_RANK or 0). Never falls back to a CPU path."""
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    eng = _default_engines.get(device)
    if eng is None:
        eng = CudaEngine(device)
        _default_engines[device] = eng
    return eng
