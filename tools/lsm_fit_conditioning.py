#!/usr/bin/env python
"""Accuracy of the two ways of fitting the LSM continuation polynomial, against a 60-digit least-squares solution:
  * Chebyshev normal equations in u = ua_t S + ub_t with the per-date interval of hh_lsm_american (csrc/hh_lsm.cu), binary64;
  * QR on the raw Vandermonde matrix of S, binary64 — what Polynomials.fit does in the reference
    (least_squares_montecarlo.jl:124-126) and what the oracle restates.
Largest difference of the FITTED VALUES over the in-the-money points, for an at-the-money put (S0 = K = 100).
   python tools/lsm_fit_conditioning.py        (CPU only; numpy + mpmath)"""
import mpmath as mp
import numpy as np

mp.mp.dps = 60
rng = np.random.default_rng(0)


def trial(deg, sigma, t, S0=100.0, K=100.0, n=6000, r=0.03):
    S = S0 * np.exp((r - 0.5 * sigma ** 2) * t + sigma * np.sqrt(t) * rng.standard_normal(n))
    s = S[S < K]
    y = np.maximum(K - s * np.exp(0.1 * rng.standard_normal(s.size)), 0) * 0.99   # a cash-flow-like target
    lo, hi = min(S0 * np.exp((r - 0.5 * sigma ** 2) * t) / np.exp(5 * sigma * np.sqrt(t)), 0.9 * K), K
    T = np.polynomial.chebyshev.chebvander(2 * (s - lo) / (hi - lo) - 1, deg)
    fit_ne = T @ np.linalg.solve(T.T @ T, T.T @ y)
    V = np.vander(s, deg + 1, increasing=True)
    Q, R = np.linalg.qr(V)
    fit_qr = V @ np.linalg.solve(R, Q.T @ y)
    Tm, ym = mp.matrix(T.tolist()), mp.matrix(y.tolist())
    fit_ex = np.array([float(x) for x in Tm * mp.lu_solve(Tm.T * Tm, Tm.T * ym)])
    return np.max(np.abs(fit_ne - fit_ex)), np.max(np.abs(fit_qr - fit_ex)), np.linalg.cond(T.T @ T), np.linalg.cond(V)


for deg in (3, 4, 5, 6, 7, 8):
    for sigma, t in ((0.2, 0.5), (0.2, 0.02), (0.5, 1.0)):
        e_ne, e_qr, cg, cv = trial(deg, sigma, t)
        print(f"degree {deg} sigma {sigma} t {t:4}: |Chebyshev normal equations - exact| {e_ne:.1e}   |QR on raw monomials - exact| {e_qr:.1e}"
              f"   cond(Gram) {cg:.1e}   cond(Vandermonde) {cv:.1e}")
