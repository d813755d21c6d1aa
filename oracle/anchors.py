"""Deterministic pricing anchors of the reference, restated (TEST INFRASTRUCTURE, not product code).

These are the oracles the reference's own Monte Carlo tests compare against:
  Black-Scholes analytic   src/pricing_methods/black_scholes.jl:38-64      KATs test/unit/black_scholes.jl:93,103,113,126
  Cox-Ross-Rubinstein      src/pricing_methods/cox_ross_rubinstein.jl:99-141  KATs test/unit/binomial_tree.jl:18,26
  Carr-Madan + Heston CF   src/pricing_methods/carr_madan.jl:47-92, src/distributions/heston.jl:307-319
They pin the statistical side of the oracle; tests/test_oracle_anchors.py checks the KATs.
"""
from __future__ import annotations

import cmath
import math

import numpy as np
from scipy import integrate
from scipy.stats import norm


def bs_price(S, K, r, sigma, T, cp=1.0):
    """black_scholes.jl:38-64 (forward-measure form; intrinsic if sigma == 0)."""
    D = math.exp(-r * T)
    F = S / D
    if sigma == 0:
        return D * max(cp * (F - K), 0.0)
    sq = math.sqrt(T)
    d1 = (math.log(F / K) + 0.5 * sigma ** 2 * T) / (sigma * sq)
    d2 = d1 - sigma * sq
    return D * cp * (F * norm.cdf(cp * d1) - K * norm.cdf(cp * d2))


def bs_greeks(S, K, r, sigma, T, cp=1.0):
    """Analytic delta, gamma, vega, rho (greeks_problem.jl:437-530 restated from the closed form)."""
    sq = math.sqrt(T)
    d1 = (math.log(S / K) + (r + 0.5 * sigma ** 2) * T) / (sigma * sq)
    d2 = d1 - sigma * sq
    delta = cp * norm.cdf(cp * d1)
    gamma = norm.pdf(d1) / (S * sigma * sq)
    vega = S * norm.pdf(d1) * sq
    rho = cp * K * T * math.exp(-r * T) * norm.cdf(cp * d2)
    return {"delta": delta, "gamma": gamma, "vega": vega, "rho": rho}


def crr_price(S, K, r, sigma, T, steps, cp=1.0, american=False, underlying="spot"):
    """cox_ross_rubinstein.jl:99-141: tree on the FORWARD with u = exp(sigma sqrt(dT)), p = 1/(1+u)."""
    D = math.exp(-r * T)
    fwd = S / D
    dT = T / steps
    u = math.exp(sigma * math.sqrt(dT))

    def forward_at(i):
        return fwd * u ** np.arange(-i, i + 1, 2, dtype=np.float64)

    def underlying_at(i):
        f = forward_at(i)
        if underlying == "spot":  # binomial_tree_underlying(..., ::Spot) :60-66
            return math.exp(-r * (steps - i) * dT) * f
        return f

    p = 1.0 / (1.0 + u)
    payoff = lambda s: np.maximum(cp * (s - K), 0.0)
    value = payoff(forward_at(steps))
    disc = math.exp(-r * dT)
    for step in range(steps - 1, -1, -1):
        cont = disc * (p * value[1:] + (1 - p) * value[:-1])
        value = np.maximum(cont, payoff(underlying_at(step))) if american else cont
    return float(value[0])


def heston_cf(u, S0, V0, kappa, theta, sigma, rho, r, T):
    """cf(::LogHestonDistribution, u)  heston.jl:307-319"""
    iu = 1j * u
    d1 = cmath.sqrt((kappa - rho * sigma * iu) ** 2 + sigma ** 2 * (iu + u ** 2))
    g = (kappa - rho * sigma * iu - d1) / (kappa - rho * sigma * iu + d1)
    Cc = (kappa * theta / sigma ** 2) * ((kappa - rho * sigma * iu - d1) * T
                                         - 2 * cmath.log((1 - g * cmath.exp(-d1 * T)) / (1 - g)))
    Dd = ((kappa - rho * sigma * iu - d1) / sigma ** 2) * ((1 - cmath.exp(-d1 * T)) / (1 - g * cmath.exp(-d1 * T)))
    return cmath.exp(Cc + Dd * V0 + iu * math.log(S0) + iu * r * T)


def gbm_cf(u, S0, r, sigma, T, q1_compat=False):
    mu = math.log(S0) + (r - sigma ** 2 / 2) * (math.sqrt(T) if q1_compat else T)
    return cmath.exp(1j * u * mu - sigma ** 2 * T / 2 * u ** 2)


def carr_madan_price(cf, S0, K, r, T, alpha=1.0, bound=32.0, cp=1.0):
    """carr_madan.jl:47-92: damped-call Fourier integral over (-bound, bound), then put-call parity."""
    logK = math.log(K)
    D = math.exp(-r * T)
    damp = math.exp(-alpha * logK) / (2 * math.pi)

    def integrand(v):
        num = D * cf(v - (alpha + 1) * 1j)
        den = alpha ** 2 + alpha - v ** 2 + v * (2 * alpha + 1) * 1j
        return (damp * num / den * cmath.exp(-1j * v * logK)).real

    val, _ = integrate.quad(integrand, -bound, bound, limit=400, epsabs=1e-12, epsrel=1e-12)
    return val if cp > 0 else val - S0 + K * D  # parity_transform payoffs.jl:172-194


def heston_price(S0, K, r, T, V0, kappa, theta, sigma, rho, alpha=1.0, bound=32.0, cp=1.0):
    return carr_madan_price(lambda u: heston_cf(u, S0, V0, kappa, theta, sigma, rho, r, T), S0, K, r, T, alpha, bound, cp)


# ---- closed forms for the path-dependent payoffs (roadmap Phase 5, derivatives_pricing_roadmap.md:73-80) --------------
def geometric_asian_price(S, K, r, sigma, T, m, cp=1.0):
    """Discretely monitored geometric-average option under Black-Scholes, dates t_i = i T / m, i = 1..m:
    log G is normal with mean log S + (r - sigma^2/2) T (m+1)/(2m) and variance sigma^2 T (m+1)(2m+1)/(6 m^2)."""
    mu = math.log(S) + (r - 0.5 * sigma ** 2) * T * (m + 1) / (2 * m)
    v = sigma ** 2 * T * (m + 1) * (2 * m + 1) / (6 * m * m)
    sv = math.sqrt(v)
    d2 = (mu - math.log(K)) / sv
    d1 = d2 + sv
    return math.exp(-r * T) * cp * (math.exp(mu + 0.5 * v) * norm.cdf(cp * d1) - K * norm.cdf(cp * d2))


def digital_price(S, K, r, sigma, T, cp=1.0, cash=None):
    """cash-or-nothing (cash = amount) or asset-or-nothing (cash = None) under Black-Scholes."""
    sq = sigma * math.sqrt(T)
    d1 = (math.log(S / K) + (r + 0.5 * sigma ** 2) * T) / sq
    d2 = d1 - sq
    if cash is None:
        return S * norm.cdf(cp * d1)
    return cash * math.exp(-r * T) * norm.cdf(cp * d2)


def up_and_in_call_price(S, K, H, r, sigma, T):
    """Continuously monitored up-and-in call, H > max(S, K) (reflection principle; Hull, Options Futures and Other
    Derivatives, 'Barrier options'), no rebate."""
    sq = sigma * math.sqrt(T)
    lam = (r + 0.5 * sigma ** 2) / sigma ** 2
    x1 = math.log(S / H) / sq + lam * sq
    y = math.log(H * H / (S * K)) / sq + lam * sq
    y1 = math.log(H / S) / sq + lam * sq
    D = math.exp(-r * T)
    return (S * norm.cdf(x1) - K * D * norm.cdf(x1 - sq)
            - S * (H / S) ** (2 * lam) * (norm.cdf(-y) - norm.cdf(-y1))
            + K * D * (H / S) ** (2 * lam - 2) * (norm.cdf(-y + sq) - norm.cdf(-y1 + sq)))


def discrete_barrier_shift(H, sigma, T, m, up=True):
    """Broadie-Glasserman-Kou continuity correction: a barrier monitored at m dates prices like a continuous one
    moved away from the spot by exp(+-0.5826 sigma sqrt(T/m))."""
    return H * math.exp((1.0 if up else -1.0) * 0.5826 * sigma * math.sqrt(T / m))
