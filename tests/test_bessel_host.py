"""CPU check of csrc/hh_bessel.cuh (host build through tools/bk_host_check.cpp): log I_nu(z) against AMOS (scipy.special.ive)
for every region of the routine — ascending series, Hankel expansion, Debye's uniform expansion (orders from 11.5 up,
between the two) and the continued fractions — and the characteristic function of the integrated variance against the
oracle's restatement of heston.jl:184-212. The device build differs only in its elementary functions (tables instead of
libm: tests/test_gpu_bk.py); the region logic, the series radius and the coefficients are the same source."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
from scipy.special import ive

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bk_host") / "libbk_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tools", "bk_host_check.cpp")], check=True)
    return ctypes.CDLL(out)


def _log_besseli(lib, nu, z):
    zr, zi = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
    o1, o2 = np.empty(z.size), np.empty(z.size)
    lib.bkh_log_besseli(ctypes.c_double(nu), P(zr), P(zi), z.size, P(o1), P(o2))
    return o1 + 1j * o2


@pytest.mark.parametrize("nu", [-0.933, -0.5, 0.0, 0.778, 6.0, 11.4, 11.6, 12.2, 15.0, 31.7, 89.0, 200.0])
def test_log_besseli_every_region(host_lib, nu):
    rng = np.random.default_rng(7)
    chunks = []
    for lo, hi in [(0.1, 6), (4, 40), (30, 300), (300, 5000), (5000, 50000)]:
        r = rng.uniform(lo, hi, 1500)
        th = rng.uniform(-1.5, 1.5, r.size)
        th[::4] = 0.0
        th[1::4] *= 0.2
        chunks.append(r * np.exp(1j * th))
    z = np.concatenate(chunks)
    with np.errstate(all="ignore"):
        ref = np.log(ive(nu, z).astype(complex)) + np.abs(z.real)
    ok = np.isfinite(ref) & (np.abs(np.abs(np.angle(z)) - np.pi / 2) > 0.05)
    got = _log_besseli(host_lib, nu, z[ok])
    d = got - ref[ok]
    d = d.real + 1j * ((d.imag + np.pi) % (2 * np.pi) - np.pi)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(d) / np.maximum(1.0, np.abs(ref[ok]))) < 5e-12


def test_nan_and_tiny_arguments_terminate(host_lib):
    z = np.array([complex(np.nan, 0.0), complex(1e-300, 0.0), complex(1e-200, 1e-200), complex(np.nan, np.nan)])
    for nu in (-0.99, 0.5, 40.0):
        got = _log_besseli(host_lib, nu, z)
        assert np.isnan(got[0].real) and np.isnan(got[3].real)
        assert np.isfinite(got[1].real) and np.isfinite(got[2].real)


@pytest.mark.parametrize("pars,tau", [(dict(kappa=2.0, theta=0.04, xi=0.3), 1 / 12), (dict(kappa=2.0, theta=0.04, xi=0.1), 0.25),
                                      (dict(kappa=5.0, theta=0.09, xi=0.1), 1 / 12), (dict(kappa=0.13, theta=0.016, xi=0.84), 0.225)])
def test_characteristic_function_matches_oracle(host_lib, pars, tau):
    from oracle import bk_ref as B
    rng = np.random.default_rng(5)
    for _ in range(6):
        V0, VT = pars["theta"] * rng.uniform(0.3, 2.0), pars["theta"] * rng.uniform(0.1, 3.0)
        cf = B.HestonCF(pars["kappa"], pars["theta"], pars["xi"], V0, VT, tau)
        mean, var = B.moments_from_cf(cf)
        h = math.pi / (mean + 5 * math.sqrt(max(var, 1e-12)))
        a = h * np.arange(1, 41)
        ref, th = np.empty(a.size, dtype=complex), math.nan
        for j in range(a.size):
            ref[j], th = cf.evaluate(a[j], th)
        o1, o2 = np.empty(a.size), np.empty(a.size)
        host_lib.bkh_chf(*(ctypes.c_double(x) for x in (pars["kappa"], pars["theta"], pars["xi"], tau, V0, VT)), P(a), a.size, P(o1), P(o2))
        zmax = 4 * pars["kappa"] * math.sqrt(V0 * VT) / (pars["xi"] ** 2 * -math.expm1(-pars["kappa"] * tau))
        assert np.max(np.abs(o1 + 1j * o2 - ref)) < max(1e-12, 4e-15 * zmax)
    # underflowed variances: finite, and the same numbers as at the floor (the reference has Inf - Inf here for nu < 0)
    a = np.array([0.01, 1.0, 50.0])
    o1, o2, f1, f2 = (np.empty(3) for _ in range(4))
    host_lib.bkh_chf(*(ctypes.c_double(x) for x in (pars["kappa"], pars["theta"], pars["xi"], tau, 1e-300, 1e-300)), P(a), 3, P(o1), P(o2))
    host_lib.bkh_chf(*(ctypes.c_double(x) for x in (pars["kappa"], pars["theta"], pars["xi"], tau, 1e-100, 1e-100)), P(a), 3, P(f1), P(f2))
    assert np.all(np.isfinite(o1)) and np.array_equal(o1, f1) and np.array_equal(o2, f2)
