"""Greeks on the B200 Monte Carlo method — mirror of src/greeks/greeks_problem.jl.

The reference differentiates by re-running `solve` with a ForwardDiff.Dual (or bumps) once per lens
(greeks_problem.jl:249-262, 559-568). Here every ForwardAD lens of a problem becomes one tangent
direction of the SAME simulation: `BatchGreekProblem` with up to 8 lenses is one kernel launch.
FiniteDifference re-solves with bumped inputs exactly like the reference (same seeds => common random numbers).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Any, Sequence

import numpy as np

from . import _abi as abi
from . import api


# ---- lenses (greeks_problem.jl:18-130, pricing_methods.jl:26-60) ----------------------------------------
@dataclass(frozen=True)
class SpotLens:
    def __call__(self, prob):
        return prob.market_inputs.spot


@dataclass(frozen=True)
class VolLens:
    strike: Any = 1
    expiry: Any = 1

    def __call__(self, prob):
        return prob.market_inputs.sigma.sigma  # FlatVolSurface (greeks_problem.jl:119-121)


@dataclass(frozen=True)
class ZeroRateSpineLens:
    i: int = 1

    def __call__(self, prob):
        return prob.market_inputs.rate.rate  # flat curve (pricing_methods.jl:30-32)


@dataclass(frozen=True)
class FieldLens:
    """`@optic _.market_inputs.<name>` for Heston scalars: V0, kappa (κ), theta (θ), sigma (σ), rho (ρ), spot."""
    name: str

    def __call__(self, prob):
        return getattr(prob.market_inputs, self.name)


def optic(path: str):
    """optic("market_inputs.spot") ~ @optic _.market_inputs.spot"""
    assert path.startswith("market_inputs."), path
    name = path.split(".", 1)[1]
    alias = {"κ": "kappa", "θ": "theta", "σ": "sigma", "ρ": "rho"}
    name = alias.get(name, name)
    if name == "spot":
        return SpotLens()
    if name in ("rate.rate",):
        return ZeroRateSpineLens(1)
    return FieldLens(name)


def set(prob, lens, newval):  # noqa: A001 - same name as Accessors.set in the reference
    m = prob.market_inputs
    if isinstance(lens, SpotLens):
        m2 = _replace_inputs(m, spot=newval)
    elif isinstance(lens, VolLens):
        m2 = _replace_inputs(m, sigma=api.FlatVolSurface(newval))
    elif isinstance(lens, ZeroRateSpineLens):
        m2 = _replace_inputs(m, rate=api.FlatRateCurve(m.rate.reference_date, newval))  # pricing_methods.jl:56-59
    elif isinstance(lens, FieldLens):
        m2 = _replace_inputs(m, **{lens.name: newval})
    else:
        raise TypeError(f"unknown lens {lens!r}")
    return api.PricingProblem(prob.payoff, m2)


def _replace_inputs(m, **kw):
    if isinstance(m, api.BlackScholesInputs):
        d = dict(reference_date=m.referenceDate, rate=m.rate, spot=m.spot, sigma=m.sigma)
        d.update(kw)
        return api.BlackScholesInputs(d["reference_date"], d["rate"], d["spot"], d["sigma"])
    d = dict(reference_date=m.referenceDate, rate=m.rate, spot=m.spot, V0=m.V0, kappa=m.kappa, theta=m.theta,
             sigma=m.sigma, rho=m.rho)
    d.update(kw)
    return api.HestonInputs(d["reference_date"], d["rate"], d["spot"], d["V0"], d["kappa"], d["theta"], d["sigma"], d["rho"])


# ---- methods & problems ------------------------------------------------------------------------------------
class ForwardAD: pass
class FDForward: pass
class FDBackward: pass
class FDCentral: pass


@dataclass(frozen=True)
class FiniteDifference:  # greeks_problem.jl:204-220
    bump: float
    scheme: Any = FDCentral()


@dataclass(frozen=True)
class GreekProblem:  # :231-234
    pricing_problem: Any
    wrt: Any


@dataclass(frozen=True)
class SecondOrderGreekProblem:  # :341-345
    pricing_problem: Any
    wrt1: Any
    wrt2: Any


@dataclass(frozen=True)
class BatchGreekProblem:  # :541-544
    pricing_problem: Any
    lenses: Sequence[Any]


@dataclass
class GreekResult:
    greek: Any
    std_error: Any = None


# ---- lens -> tangent seed ------------------------------------------------------------------------------------
def tangent_of(prob, method, lens) -> abi.hh_tangent:
    t = abi.hh_tangent()
    m = prob.market_inputs
    heston = isinstance(m, api.HestonInputs)
    if isinstance(lens, SpotLens) or (isinstance(lens, FieldLens) and lens.name == "spot"):
        t.dS0 = 1.0
    elif isinstance(lens, VolLens):
        if heston:
            raise TypeError("VolLens addresses a vol surface; HestonInputs has none")
        t.dsigma = 1.0
    elif isinstance(lens, ZeroRateSpineLens):
        # the rate moves the drift (montecarlo.jl:150) AND the discount factor (rate_curve.jl:149-150)
        t.dr = 1.0
        t.ddiscount = -api.yearfrac(m.rate.reference_date, prob.payoff.expiry) * api.df(m.rate, prob.payoff.expiry)
    elif isinstance(lens, FieldLens) and heston:
        if lens.name == "V0":
            t.dV0 = 1.0
        elif lens.name == "kappa":
            t.dkappa = 1.0
        elif lens.name == "theta":
            t.dtheta = 1.0
        elif lens.name == "sigma":
            t.dxi = 1.0
        elif lens.name == "rho":
            _, (t.dm11, t.dm12, t.dm21, t.dm22) = api.corr_factor(float(m.rho), method.corr_mode)
        else:
            raise TypeError(f"no tangent rule for field {lens.name!r}")
    else:
        raise TypeError(f"no tangent rule for lens {lens!r}")
    return t


def _forward_ad(prob, lenses, method, engine, shard, group, strikes=None, spot_bump=None):
    """d price / d lens for all lenses (chunks of 8 directions per launch). Returns (greeks, stderrs, prices), and with
    `spot_bump` (absolute, greeks_problem.jl:395-412) a fourth element: the second derivative in the spot from the SAME
    launch, {"fd": D mean(sd)/eps^2 (the reference's three-solve form), "pathwise": D mean(dd)/(2 eps), and their standard
    errors} per payoff (include/hedgehog_mc.h, hh_mc_european_tangent_sums)."""
    if isinstance(method, api.LSM):
        raise NotImplementedError("ForwardAD through LSM is not on the GPU path; use FiniteDifference")
    if not isinstance(prob.payoff.exercise_style, api.European):
        raise TypeError("pathwise Greeks are defined for European payoffs")
    eng = engine or api.default_engine()
    shard, reduce = api._shard_and_reduce(shard, group, getattr(eng, "device", None))
    mdl = api._model_of(prob, method)
    scheme = api._scheme_of(method)
    sim = api._sim_of(method, scheme, shard)
    cp = prob.payoff.call_put()
    payoffs = [(prob.payoff.strike, cp)] if strikes is None else [(float(k), cp) for k in strikes]
    D = api.df(prob.market_inputs.rate, prob.payoff.expiry)
    greeks = np.zeros((len(payoffs), len(lenses)))
    stderrs = np.zeros_like(greeks)
    prices = np.zeros(len(payoffs))
    second = None
    for c0 in range(0, len(lenses), 8):
        chunk = lenses[c0:c0 + 8]
        tans = [tangent_of(prob, method, L) for L in chunk]
        nt = len(tans)
        sec = None
        if spot_bump and c0 == 0:
            sums, _, sec = eng.tangent_sums(mdl, tans, sim, payoffs, spot_bump=float(spot_bump))
        else:
            sums, _ = eng.tangent_sums(mdl, tans, sim, payoffs)
        n = np.array([float(sim.n_paths)])
        if reduce is not None:
            parts = [sums.ravel(), n] + ([sec.ravel()] if sec is not None else [])
            flat = reduce(np.concatenate(parts))
            k = sums.size
            if sec is not None:
                sec = flat[k + 1:].reshape(sec.shape)
            sums, n = flat[:k].reshape(sums.shape), flat[k:k + 1]
        N = n[0]
        if sec is not None:
            eps = float(spot_bump)
            m_sd, m_dd = sec[:, 0] / N, sec[:, 2] / N
            v_sd = np.maximum((sec[:, 1] - N * m_sd * m_sd) / max(N - 1, 1), 0.0)
            v_dd = np.maximum((sec[:, 3] - N * m_dd * m_dd) / max(N - 1, 1), 0.0)
            second = {"fd": D * m_sd / eps ** 2, "fd_stderr": D * np.sqrt(v_sd / N) / eps ** 2,
                      "pathwise": D * m_dd / (2 * eps), "pathwise_stderr": D * np.sqrt(v_dd / N) / (2 * eps), "bump": eps}
        mean = sums[:, 0] / N
        prices[:] = D * mean
        for q, t in enumerate(tans):
            dmean = sums[:, 2 + q] / N
            # price = D mean(payoff)  =>  d price = dD mean(payoff) + D mean(d payoff)   (montecarlo.jl:489-490)
            greeks[:, c0 + q] = t.ddiscount * mean + D * dmean
            var = np.maximum((sums[:, 2 + nt + q] - N * dmean * dmean) / max(N - 1, 1), 0.0)
            stderrs[:, c0 + q] = D * np.sqrt(var / N)
    if spot_bump:
        return greeks, stderrs, prices, second
    return greeks, stderrs, prices


def _price(prob, method, engine, shard, group):
    m2 = replace(method, ensemble=False) if isinstance(method, api.MonteCarlo) else method
    return api.solve(prob, m2, engine=engine, shard=shard, group=group).price


def solve_greek(gprob, gmethod, pricing_method, *, engine=None, shard=None, group=None):
    kw = dict(engine=engine, shard=shard, group=group)
    if isinstance(gprob, BatchGreekProblem):  # greeks_problem.jl:559-568 -> Dict(lens => greek)
        if isinstance(gmethod, ForwardAD):
            g, _, _ = _forward_ad(gprob.pricing_problem, list(gprob.lenses), pricing_method, engine, shard, group)
            return {lens: float(g[0, i]) for i, lens in enumerate(gprob.lenses)}
        return {lens: solve_greek(GreekProblem(gprob.pricing_problem, lens), gmethod, pricing_method, **kw).greek
                for lens in gprob.lenses}

    prob = gprob.pricing_problem
    if isinstance(gprob, GreekProblem):
        lens = gprob.wrt
        if isinstance(gmethod, ForwardAD):  # :249-262
            g, se, _ = _forward_ad(prob, [lens], pricing_method, engine, shard, group)
            return GreekResult(float(g[0, 0]), float(se[0, 0]))
        if isinstance(gmethod, FiniteDifference):  # :279-329, RELATIVE bump
            x0, eps = lens(prob), gmethod.bump
            f = lambda x: _price(set(prob, lens, x), pricing_method, engine, shard, group)
            if isinstance(gmethod.scheme, FDForward):
                return GreekResult((f(x0 * (1 + eps)) - f(x0)) / (x0 * eps))
            if isinstance(gmethod.scheme, FDBackward):
                return GreekResult((f(x0) - f(x0 * (1 - eps))) / (x0 * eps))
            return GreekResult((f(x0 * (1 + eps)) - f(x0 * (1 - eps))) / (2 * eps * x0))
        raise TypeError(f"unknown Greek method {gmethod!r}")

    if isinstance(gprob, SecondOrderGreekProblem):
        l1, l2 = gprob.wrt1, gprob.wrt2
        x0, y0 = l1(prob), l2(prob)
        if isinstance(gmethod, FiniteDifference):  # :395-422, ABSOLUTE bump
            eps = gmethod.bump
            if (_is_spot(l1) and _is_spot(l2) and isinstance(pricing_method, api.MonteCarlo)
                    and isinstance(prob.payoff.exercise_style, api.European)
                    and not isinstance(pricing_method.strategy, api.HestonBroadieKaya) and pricing_method.precision == "f64"
                    and pricing_method.control_variate is None and 0 < eps < x0):
                # gamma: the reference's three solves at S0 - eps, S0, S0 + eps share their seeds, and every scheme here is
                # linear in S0, so the three payoffs come from ONE simulation (hh_mc_european_tangent_sums, second_sums)
                _, _, _, sec = _forward_ad(prob, [SpotLens()], pricing_method, engine, shard, group, spot_bump=eps)
                return GreekResult(float(sec["fd"][0]), float(sec["fd_stderr"][0]))
            f = lambda x, y: _price(set(set(prob, l1, x), l2, y), pricing_method, engine, shard, group)
            if l1 == l2:
                return GreekResult((f(x0 + eps, y0 + eps) - 2 * f(x0, y0) + f(x0 - eps, y0 - eps)) / eps ** 2)
            return GreekResult((f(x0 + eps, y0 + eps) - f(x0 + eps, y0 - eps) - f(x0 - eps, y0 + eps)
                                + f(x0 - eps, y0 - eps)) / (4 * eps ** 2))
        if isinstance(gmethod, ForwardAD):
            # The reference nests Duals (:360-380); the pathwise second derivative of a vanilla payoff is
            # a.s. zero (its own MC test uses FD "due to AD instability", test/agreement/greeks_agreement.jl:219-224).
            # Here: central difference, on common random numbers, of the in-kernel first-order tangent.
            eps = 1e-2 * abs(x0) if x0 != 0 else 1e-4
            if _is_spot(l1) and _is_spot(l2) and isinstance(pricing_method, api.MonteCarlo) and x0 > 0:
                # both bumped deltas come from the one simulation (linear in S0): no extra launches
                _, _, _, sec = _forward_ad(prob, [SpotLens()], pricing_method, engine, shard, group, spot_bump=eps)
                return GreekResult(float(sec["pathwise"][0]), float(sec["pathwise_stderr"][0]))
            up, _, _ = _forward_ad(set(prob, l1, x0 + eps), [l2], pricing_method, engine, shard, group)
            dn, _, _ = _forward_ad(set(prob, l1, x0 - eps), [l2], pricing_method, engine, shard, group)
            return GreekResult(float((up[0, 0] - dn[0, 0]) / (2 * eps)))
    raise TypeError(f"unknown Greek problem {gprob!r}")


def strike_grid_greeks(prob, strikes, lenses, pricing_method, *, engine=None, shard=None, group=None, gamma_bump=None):
    """Config C5: all `lenses` x all `strikes` from ONE simulation. Returns (prices[k], greeks[k, lens], stderr[k, lens]);
    with `gamma_bump` (the absolute spot bump of SecondOrderGreekProblem + FiniteDifference, greeks_problem.jl:395-412) a
    fourth element {"fd", "fd_stderr", "pathwise", "pathwise_stderr"}: gamma per strike from the same launch."""
    out = _forward_ad(prob, list(lenses), pricing_method, engine, shard, group, strikes=list(strikes), spot_bump=gamma_bump)
    if gamma_bump:
        g, se, prices, second = out
        return prices, g, se, second
    g, se, prices = out
    return prices, g, se


def _is_spot(lens):
    return isinstance(lens, SpotLens) or (isinstance(lens, FieldLens) and lens.name == "spot")
