"""Calibration on the B200 Monte Carlo method — mirror of src/calibration/calibration.jl ("next" row N2 of SURVEY.md §8f).

The reference minimises sum_k (price_k(x) - quote_k)^2 over the parameters selected by Accessors lenses, with the gradient
from AutoForwardDiff through `solve(BasketPricingProblem, method)` — one full simulation per parameter per payoff.
Here the objective AND its gradient with respect to all calibrated parameters, for every quote that shares an expiry and
a call/put flag, come from ONE launch of the tangent kernel on common random numbers (fixed Philox key), so the Monte
Carlo objective is a smooth deterministic function of the parameters and L-BFGS converges as on an analytic pricer.
Nothing numerical happens here besides the optimiser loop (scipy's L-BFGS-B / Brent, the counterparts of
OptimizationOptimJL.LBFGS / Brent); prices and tangents come from libhedgehog_mc.so.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional, Sequence

import numpy as np

from . import api
from . import greeks as _g


@dataclass(frozen=True)
class CalibrationProblem:  # calibration.jl:16-29
    pricing_problem: api.BasketPricingProblem
    pricing_method: Any
    accessors: Sequence[Any]
    quotes: Sequence[float]
    initial_guess: Sequence[float]


@dataclass(frozen=True)
class OptimizerAlgo:  # calibration.jl:46-58 (AutoForwardDiff + LBFGS)
    diff: str = "forward"        # the in-kernel pathwise tangents (the reference: AutoForwardDiff)
    optim_algo: str = "L-BFGS-B"


@dataclass(frozen=True)
class RootFinderAlgo:  # calibration.jl:105-117 (Brent on (1e-6, 5.0))
    root_method: str = "brentq"


@dataclass
class CalibrationResult:
    u: np.ndarray                 # fitted parameters (the reference's result.u)
    objective: float
    iterations: int
    evaluations: int
    success: bool
    message: str = ""
    history: list = field(default_factory=list)


def _set_all(basket, accessors, x):
    """foldl(set, zip(accessors, x); init = basket)   calibration.jl:78-82"""
    prob = api.PricingProblem(basket.payoffs[0], basket.market_inputs)
    for lens, val in zip(accessors, x):
        prob = _g.set(prob, lens, float(val))
    return api.BasketPricingProblem(basket.payoffs, prob.market_inputs)


def basket_prices_and_jacobian(basket, accessors, method, *, engine=None, shard=None, group=None):
    """prices[k] and d price_k / d accessor_j for every payoff of the basket: one tangent launch per (expiry, call/put)."""
    n = len(basket.payoffs)
    prices = np.zeros(n)
    jac = np.zeros((n, len(accessors)))
    groups: dict = {}
    for i, p in enumerate(basket.payoffs):
        if not isinstance(p.exercise_style, api.European):
            raise TypeError("MC calibration prices European payoffs")
        groups.setdefault((p.expiry, type(p.call_put)), []).append(i)
    for idxs in groups.values():
        for c0 in range(0, len(idxs), 256):
            chunk = idxs[c0:c0 + 256]
            prob = api.PricingProblem(basket.payoffs[chunk[0]], basket.market_inputs)
            g, _, pr = _g._forward_ad(prob, list(accessors), method, engine, shard, group,
                                      strikes=[basket.payoffs[i].strike for i in chunk])
            prices[chunk] = pr
            jac[chunk, :] = g
    return prices, jac


def solve_calibration(calib: CalibrationProblem, algo, *, lb=None, ub=None, maxiters: int = 100, f_abstol: float = 0.0,
                      g_tol: float = 1e-10, engine=None, shard=None, group=None) -> CalibrationResult:
    quotes = np.asarray(calib.quotes, dtype=np.float64)
    basket = calib.pricing_problem
    if isinstance(algo, RootFinderAlgo):  # calibration.jl:126-144
        if len(calib.accessors) != 1 or len(quotes) != 1:
            raise AssertionError("Root-finding only supports calibration of a single parameter and a single quote")
        from scipy.optimize import brentq
        lens = calib.accessors[0]
        pp = api.PricingProblem(basket.payoffs[0], basket.market_inputs)
        evals = [0]

        def f(x):
            evals[0] += 1
            return _g._price(_g.set(pp, lens, float(x)), calib.pricing_method, engine, shard, group) - quotes[0]
        root, info = brentq(f, 1e-6, 5.0, full_output=True, maxiter=maxiters)
        return CalibrationResult(np.array([root]), abs(f(root)) ** 2, info.iterations, evals[0], info.converged)

    if not isinstance(algo, OptimizerAlgo):
        raise TypeError(f"unknown calibration algorithm {algo!r}")
    from scipy.optimize import minimize
    hist = []

    def fun(x):  # objective calibration.jl:75-88 and its exact gradient 2 J^T (prices - quotes)
        prices, jac = basket_prices_and_jacobian(_set_all(basket, calib.accessors, x), calib.accessors, calib.pricing_method,
                                                 engine=engine, shard=shard, group=group)
        err = prices - quotes
        val = float(err @ err)
        hist.append(val)
        return val, 2.0 * (jac.T @ err)

    bounds = None
    if lb is not None or ub is not None:
        n = len(calib.accessors)
        lo = [None] * n if lb is None else list(lb)
        hi = [None] * n if ub is None else list(ub)
        bounds = list(zip(lo, hi))
    x0 = np.asarray(calib.initial_guess, dtype=np.float64)[:len(calib.accessors)]
    res = minimize(fun, x0, jac=True, method=algo.optim_algo, bounds=bounds,
                   options={"maxiter": maxiters, "ftol": f_abstol, "gtol": g_tol})
    return CalibrationResult(np.asarray(res.x), float(res.fun), int(res.nit), int(res.nfev), bool(res.success), str(res.message), hist)
