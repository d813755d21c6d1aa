"""CPU ORACLE for the Broadie-Kaya exact Heston sampler — TEST INFRASTRUCTURE, not product code.

A numpy/scipy restatement of the reference's algorithm (paths relative to the reference checkout):
  sample_V_T            src/distributions/heston.jl:125-133      c * NoncentralChisq(d, lambda)
  HestonCFIterator      src/distributions/heston.jl:150-176
  evaluate_chf          src/distributions/heston.jl:184-212      (complex besseli + angle unwrapping)
  moments_from_cf       src/distributions/sample_from_cf.jl:50-64
  cdf_from_cf           src/distributions/sample_from_cf.jl:75-96
  inverse_cdf           src/distributions/sample_from_cf.jl:105-135
  sample_from_cf        src/distributions/sample_from_cf.jl:27-41
  sample_log_S_T        src/distributions/heston.jl:278-300

The modified Bessel function is scipy.special.ive, i.e. AMOS zbesi — the same library SpecialFunctions.besseli wraps in
the reference (third-party there; SpecialFunctions 2.5, Project.toml:40). Root finding: the reference calls
Roots.find_zero(f, x0, Order2(); atol, maxeval) and falls back to bisection; Order2's exact iterate sequence is
upstream behaviour that cannot be observed here, so `inverse_cdf` below is a secant iteration with the reference's
acceptance test (root >= 0 and |F(x) - u| <= atol) and its fallbacks. PARITY STATUS: "parity unpinned" per sample
(no golden vectors in the reference; its own inversion tolerance is 1e-4); the deterministic pieces (CF values, moments,
the Fourier CDF) are what the GPU is compared against at tight tolerance, the sampler statistically (Carr-Madan).

Only tests/ may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy.special import ive
from scipy.stats import norm


@dataclass
class HestonCF:
    """HestonCFIterator (heston.jl:150-176) for one (V0, VT) pair and horizon T."""
    kappa: float
    theta: float
    sigma: float
    V0: float
    VT: float
    T: float

    def __post_init__(self):
        k, s, T = self.kappa, self.sigma, self.T
        d = 4 * k * self.theta / s ** 2
        self.nu = 0.5 * d - 1
        E = -math.expm1(-k * T)
        self.zeta_k = E / k
        self.eta_k = k * (1 + math.exp(-k * T)) / E
        nu_k = math.sqrt(self.V0 * self.VT) * 4 * k * math.exp(-0.5 * k * T) / s ** 2 / E
        self.logI_k = log_besseli(self.nu, complex(nu_k))

    def evaluate(self, a: float, theta_prev: float):
        """evaluate_chf (heston.jl:184-212). Returns (phi, theta_unwrapped)."""
        k, s, T, V0, VT = self.kappa, self.sigma, self.T, self.V0, self.VT
        g = np.sqrt(complex(k ** 2, -2 * s ** 2 * a))
        eg = np.exp(-g * T)
        zeta_g = (1 - eg) / g
        eta_g = g * (1 + eg) / (1 - eg)
        nu_g = math.sqrt(V0 * VT) * 4 * g * np.exp(-0.5 * g * T) / s ** 2 / (1 - eg)
        first = np.exp(-0.5 * (g - k) * T) * (self.zeta_k / zeta_g)
        second = np.exp((V0 + VT) / s ** 2 * (self.eta_k - eta_g))
        th = math.atan2(nu_g.imag, nu_g.real)
        if math.isnan(theta_prev):
            thu = th
        else:
            dlt = th - theta_prev
            dlt -= 2 * math.pi * round(dlt / (2 * math.pi))
            thu = theta_prev + dlt
        z_unw = abs(nu_g) * complex(math.cos(thu), math.sin(thu))
        logI_g = log_besseli(self.nu, z_unw) + 1j * self.nu * (thu - th)
        return first * second * np.exp(logI_g - self.logI_k), thu


def log_besseli(nu: float, z: complex) -> complex:
    """log(besseli(nu, z)) through the exponentially scaled AMOS routine (finite where besseli itself overflows)."""
    return np.log(ive(nu, z)) + abs(z.real)


def moments_from_cf(cf: HestonCF, h: float = 1e-2):
    """sample_from_cf.jl:50-64 — central differences, the three evaluations share the unwrapping state."""
    th = math.nan
    pp, th = cf.evaluate(h, th)
    p0, th = cf.evaluate(0.0, th)
    pm, _ = cf.evaluate(-h, th)
    first = (pp - pm) / (2 * h)
    second = (pp - 2 * p0 + pm) / h ** 2
    mean = (-1j * first).real
    var = (-second - mean ** 2).real
    return mean, var


def cf_series(cf: HestonCF, h: float, cf_tol: float = 1e-3, max_terms: int = 10 ** 9):
    """The x-independent part of cdf_from_cf (sample_from_cf.jl:84-93): phi(h j) for j = 1.. until |phi|/j < pi cf_tol/2."""
    out = []
    th = math.nan
    for j in range(1, max_terms + 1):
        phi, th = cf.evaluate(h * j, th)
        out.append(phi)
        if abs(phi) / j < math.pi * cf_tol / 2:
            break
    return np.array(out)


def cdf_from_series(phis: np.ndarray, x: float, h: float) -> float:
    """cdf_from_cf (sample_from_cf.jl:75-96) given the precomputed phi(h j)."""
    if x < 0:
        return 0.0
    j = np.arange(1, len(phis) + 1)
    return h * x / math.pi + float(np.sum(2 / math.pi * np.sin(h * j * x) / j * phis.real))


def inverse_cdf(cdf, u, guess, max_guess, atol=1e-4, maxiter_newton=10, maxiter_bisection=100):
    """sample_from_cf.jl:105-135. Returns (x, status): 0 = secant accepted, 1 = bisection, 2 = fell back to max_guess."""
    f = lambda y: cdf(y) - u
    # secant from (guess, guess * (1 + 1e-3)) — stand-in for Roots.Order2 (upstream), <= maxiter_newton evaluations
    x0, f0 = guess, f(guess)
    x1 = guess * 1.001 + 1e-12
    f1 = f(x1)
    evals = 2
    sol = None
    while evals < maxiter_newton:
        if abs(f1) <= atol * 1e-6 or f1 == f0:
            break
        x2 = x1 - f1 * (x1 - x0) / (f1 - f0)
        x0, f0, x1 = x1, f1, x2
        f1 = f(x1)
        evals += 1
    sol = x1
    if math.isfinite(sol) and sol >= 0 and abs(f(sol)) <= atol:
        return sol, 0
    if f(0.0) * f(max_guess) > 0:
        return max_guess, 2  # the reference @warns and returns the fall-back (u ~ 1)
    lo, hi = 0.0, max_guess
    flo = f(lo)
    for _ in range(maxiter_bisection):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        if (fm > 0) == (flo > 0):
            lo, flo = mid, fm
        else:
            hi = mid
        if hi - lo <= atol:
            break
    return 0.5 * (lo + hi), 1


def sample_integral_V(cf: HestonCF, u: float, n: int = 5, h_fd: float = 1e-2, cf_tol: float = 1e-3, atol: float = 1e-4):
    """sample_from_cf (sample_from_cf.jl:27-41) for a GIVEN uniform u. Returns a dict with every intermediate."""
    mean, var = moments_from_cf(cf, h_fd)
    s2 = max(var, 1e-12)
    sd = math.sqrt(s2)
    ns = mean + sd * norm.ppf(u)
    guess = ns if ns > 0 else mean * 0.01
    max_guess = mean + 11 * sd
    h = math.pi / (mean + n * sd)
    phis = cf_series(cf, h, cf_tol)
    cdf = lambda x: cdf_from_series(phis, x, h)
    x, status = inverse_cdf(cdf, u, guess, max_guess, atol)
    return {"mean": mean, "var": var, "h": h, "J": len(phis), "phis": phis, "guess": guess, "max_guess": max_guess,
            "x": x, "status": status, "cdf": cdf}


def vt_params(kappa, theta, sigma, V0, T):
    """sample_V_T (heston.jl:125-133): V_T = c * NoncentralChisq(d, lambda)."""
    d = 4 * kappa * theta / sigma ** 2
    E = -math.expm1(-kappa * T)
    lam = 4 * kappa * math.exp(-kappa * T) * V0 / (sigma ** 2 * E)
    c = sigma ** 2 * E / (4 * kappa)
    return d, lam, c


def log_S_T(S0, V0, VT, IV, Z, kappa, theta, sigma, rho, r, T):
    """sample_log_S_T (heston.jl:278-300)."""
    mu = math.log(S0) + r * T - 0.5 * IV + (rho / sigma) * (VT - V0 - kappa * theta * T + kappa * IV)
    return mu + math.sqrt((1 - rho ** 2) * IV) * Z
