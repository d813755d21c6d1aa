"""Path-dependent payoffs priced on the Monte Carlo path — SURVEY §8(f) N4.

The reference has none of these yet; its roadmap lists them as Phase 5 (derivatives_pricing_roadmap.md:73-80:
arithmetic / geometric Asians, digitals, discretely monitored barriers, "Monitoring / Averaging modifiers"). The types
below follow the conventions of `VanillaOption` (payoffs.jl:101-140: strike, expiry, call_put) so that
`solve(PricingProblem(payoff, market_inputs), MonteCarlo(dynamics, strategy, config))` reads like the reference's
European solve; all numerics run in `hh_mc_path_dependent` (csrc/hh_pathdep.cu)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Optional, Sequence

import numpy as np

from . import _abi as abi
from . import api


# ---- modifiers ----------------------------------------------------------------------------------------------------------
class ArithmeticAverage: pass


class GeometricControlVariate:
    """Control variate for the arithmetic Asian under Black-Scholes (roadmap "control variates"): the kernel
    accumulates the payoff DIFFERENCE arithmetic - geometric on the same trajectories (correlation ~0.999) and the host
    adds the closed-form price of the discretely monitored geometric Asian."""
class GeometricAverage: pass
class Up: pass
class Down: pass
class KnockIn: pass
class KnockOut: pass
class AssetOrNothing: pass


@dataclass(frozen=True)
class CashOrNothing:
    amount: float = 1.0


@dataclass(frozen=True)
class Monitoring:
    """Monitoring dates = every `every`-th step of SimulationConfig.steps (t = 0 excluded, expiry included)."""
    every: int = 1


class _PathPayoff:
    exercise_style = api.European()
    underlying = api.Spot()

    def _base(self, strike, expiry_date, call_put, monitoring):
        object.__setattr__(self, "strike", strike)
        object.__setattr__(self, "expiry", api.to_ticks(expiry_date))
        object.__setattr__(self, "call_put", call_put)
        object.__setattr__(self, "monitoring", monitoring or Monitoring())


@dataclass(frozen=True, init=False)
class AsianOption(_PathPayoff):
    """max(cp (avg S - K), 0), fixed strike, average over the monitoring dates."""
    strike: float
    expiry: int
    call_put: Any
    averaging: Any
    monitoring: Monitoring

    control_variate: Any

    def __init__(self, strike, expiry_date, call_put, averaging=None, monitoring=None, control_variate=None):
        self._base(strike, expiry_date, call_put, monitoring)
        object.__setattr__(self, "averaging", averaging or ArithmeticAverage())
        object.__setattr__(self, "control_variate", control_variate)
        if control_variate is not None and isinstance(self.averaging, GeometricAverage):
            raise ValueError("the geometric Asian has a closed form; the control variate applies to the arithmetic average")

    def abi_tuple(self):
        kind = abi.HH_PD_ASIAN_GEOM if isinstance(self.averaging, GeometricAverage) else abi.HH_PD_ASIAN_ARITH
        if self.control_variate is not None:
            kind = abi.HH_PD_ASIAN_ARITH_MINUS_GEOM
        return (kind, self.strike, self.call_put(), 0.0, 0.0)


@dataclass(frozen=True, init=False)
class BarrierOption(_PathPayoff):
    """Vanilla payoff at expiry, switched on (KnockIn) or off (KnockOut) when the spot is at or beyond `barrier` on a
    monitoring date; otherwise `rebate`, paid at expiry."""
    strike: float
    barrier: float
    expiry: int
    call_put: Any
    direction: Any
    knock: Any
    rebate: float
    monitoring: Monitoring

    def __init__(self, strike, barrier, expiry_date, call_put, direction, knock, rebate=0.0, monitoring=None):
        self._base(strike, expiry_date, call_put, monitoring)
        object.__setattr__(self, "barrier", barrier)
        object.__setattr__(self, "direction", direction)
        object.__setattr__(self, "knock", knock)
        object.__setattr__(self, "rebate", rebate)

    def abi_tuple(self):
        up, out = isinstance(self.direction, Up), isinstance(self.knock, KnockOut)
        kind = {(True, True): abi.HH_PD_UP_OUT, (True, False): abi.HH_PD_UP_IN,
                (False, True): abi.HH_PD_DOWN_OUT, (False, False): abi.HH_PD_DOWN_IN}[(up, out)]
        return (kind, self.strike, self.call_put(), self.barrier, self.rebate)


@dataclass(frozen=True, init=False)
class DigitalOption(_PathPayoff):
    """CashOrNothing(amount) or AssetOrNothing() if cp (S_T - K) > 0."""
    strike: float
    expiry: int
    call_put: Any
    payout: Any
    monitoring: Monitoring

    def __init__(self, strike, expiry_date, call_put, payout=None, monitoring=None):
        self._base(strike, expiry_date, call_put, monitoring)
        object.__setattr__(self, "payout", payout or CashOrNothing(1.0))

    def abi_tuple(self):
        if isinstance(self.payout, AssetOrNothing):
            return (abi.HH_PD_DIGITAL_ASSET, self.strike, self.call_put(), 0.0, 0.0)
        return (abi.HH_PD_DIGITAL_CASH, self.strike, self.call_put(), 0.0, self.payout.amount)


def geometric_asian_closed_form(S, K, r, sigma, T, m, cp):
    """Discretely monitored geometric-average option under Black-Scholes, dates i T / m, i = 1..m: log G is normal with
    mean log S + (r - sigma^2/2) T (m+1)/(2m) and variance sigma^2 T (m+1)(2m+1)/(6 m^2)."""
    from statistics import NormalDist
    N = NormalDist().cdf
    mu = math.log(S) + (r - 0.5 * sigma * sigma) * T * (m + 1) / (2 * m)
    v = sigma * sigma * T * (m + 1) * (2 * m + 1) / (6 * m * m)
    sv = math.sqrt(v)
    d2 = (mu - math.log(K)) / sv
    d1 = d2 + sv
    return math.exp(-r * T) * cp * (math.exp(mu + 0.5 * v) * N(cp * d1) - K * N(cp * d2))


def _abi_tuple(p):
    if isinstance(p, api.VanillaOption):
        if not isinstance(p.exercise_style, api.European) or not isinstance(p.underlying, api.Spot):
            raise TypeError("path-dependent baskets take European options on the Spot underlying")
        return (abi.HH_PD_VANILLA, p.strike, p.call_put(), 0.0, 0.0)
    return p.abi_tuple()


def is_path_payoff(p) -> bool:
    return isinstance(p, _PathPayoff)


def solve_path_dependent(payoffs: Sequence, market_inputs, method, *, engine=None, shard=None, group=None):
    """Prices `payoffs` (same expiry, same monitoring) on ONE set of trajectories; returns (price, std_error) per
    payoff and the launch statistics. Multi-GPU: each rank simulates its block of trajectories and the (sum, sumsq, n)
    triples are sum-reduced, as in the European solve."""
    p0 = payoffs[0]
    every = getattr(p0, "monitoring", Monitoring()).every
    for p in payoffs:
        if p.expiry != p0.expiry:
            raise ValueError("payoffs priced on common trajectories must share the expiry")
        if getattr(p, "monitoring", Monitoring(every)).every != every:
            raise ValueError("payoffs priced on common trajectories must share the monitoring dates")
    eng = engine or api.default_engine()
    shard, reduce = api._shard_and_reduce(shard, group, getattr(eng, "device", None))
    prob0 = api.PricingProblem(p0, market_inputs)
    mdl = api._model_of(prob0, method)
    scheme = api._scheme_of(method, for_lsm=True)                             # stepping form of BlackScholesExact
    if scheme == abi.HH_SCHEME_HESTON_BK:                                       # exact transitions between config.steps dates
        from dataclasses import replace
        method = replace(method, bk_steps_from_config=True)
    sim = api._sim_of(method, scheme, shard)
    discount = api.df(market_inputs.rate, p0.expiry)
    results, _ = eng.mc_path_dependent(mdl, sim, [_abi_tuple(p) for p in payoffs], discount, every)
    sums = np.array([[r.sum, r.sumsq, float(r.n)] for r in results], dtype=np.float64)
    if reduce is not None:
        sums = reduce(sums)
    out = []
    for p, (s, q, n) in zip(payoffs, sums):
        mean = s / n
        var = max((q - n * mean * mean) / (n - 1), 0.0) if n > 1 else 0.0
        price = discount * mean
        if getattr(p, "control_variate", None) is not None:
            if not isinstance(market_inputs, api.BlackScholesInputs):
                raise TypeError("GeometricControlVariate needs the closed-form geometric Asian: BlackScholesInputs only")
            T = api.yearfrac(market_inputs.referenceDate, p.expiry)
            price += geometric_asian_closed_form(market_inputs.spot, p.strike, api.zero_rate(market_inputs.rate, 0.0),
                                                 api.get_vol(market_inputs.sigma), T, sim.n_steps // every, p.call_put())
        out.append((price, discount * math.sqrt(var / n)))
    stats = {"kernel_ms": results[0].kernel_ms, "n_nonfinite": results[0].n_nonfinite, "n_local": sim.n_paths,
             "n_total": int(sums[0][2])}
    return out, stats


# ---- Black-Scholes control variate for Heston vanilla prices (roadmap "Control variates using Black-Scholes", N3) -----
@dataclass(frozen=True)
class BlackScholesControlVariate:
    """MonteCarlo(HestonDynamics(), EulerMaruyama(), config, control_variate=BlackScholesControlVariate()).
    The kernel advances a log-GBM trajectory on the same Brownian increments next to every Heston trajectory
    (HH_PD_VANILLA_MINUS_BS); price = D mean(X - beta Y) + beta BS(sigma_cv). beta = None: estimated from a pilot run of
    `pilot` trajectories on a separate seed (the same on every rank), so that the main estimator stays unbiased."""
    beta: Optional[float] = None
    pilot: int = 50_000


def bs_control_sigma(V0, kappa, theta, T):
    """sigma_cv of the control: sqrt of the mean of E[V_t] = theta + (V0 - theta) e^(-kappa t) over [0, T]
    (the same formula as bs_control_variance in csrc/hh_pathdep.cu)."""
    kT = kappa * T
    w = -math.expm1(-kT) / kT if abs(kT) > 1e-8 else 1.0 - 0.5 * kT
    return math.sqrt(max(theta + (V0 - theta) * w, 1e-12))


def black_scholes_closed_form(S, K, r, sigma, T, cp):
    from statistics import NormalDist
    N = NormalDist().cdf
    sq = sigma * math.sqrt(T)
    d1 = (math.log(S / K) + (r + 0.5 * sigma * sigma) * T) / sq
    return cp * (S * N(cp * d1) - K * math.exp(-r * T) * N(cp * (d1 - sq)))


def _var(r):
    n = float(r.n)
    mean = r.sum / n
    return max((r.sumsq - n * mean * mean) / (n - 1), 0.0)


def solve_bs_control(prob, method, *, engine=None, shard=None, group=None):
    from .engine import SimSpec
    cv = method.control_variate
    if not (isinstance(method.dynamics, api.HestonDynamics) and isinstance(method.strategy, api.EulerMaruyama)):
        raise TypeError("BlackScholesControlVariate runs next to HestonDynamics + EulerMaruyama")
    if not isinstance(prob.payoff, api.VanillaOption) or not isinstance(prob.payoff.underlying, api.Spot):
        raise TypeError("BlackScholesControlVariate prices European vanilla options on the Spot underlying")
    eng = engine or api.default_engine()
    shard, reduce = api._shard_and_reduce(shard, group, getattr(eng, "device", None))
    mdl = api._model_of(prob, method)
    sim = api._sim_of(method, abi.HH_SCHEME_EM, shard)
    K, cp = prob.payoff.strike, prob.payoff.call_put()
    discount = api.df(prob.market_inputs.rate, prob.payoff.expiry)
    sigma_cv = bs_control_sigma(mdl.V0, mdl.kappa, mdl.theta, mdl.T)
    bs = black_scholes_closed_form(mdl.S0, K, mdl.r, sigma_cv, mdl.T, cp)
    beta = cv.beta
    if beta is None:  # Cov(X, Y) = (Var X + Var Y - Var(X - Y)) / 2 from three sums of ONE pilot launch
        # explicit per-trajectory seeds leave base_seed unset: the pilot's key then derives from the first seed
        base = method.config.base_seed if method.config.base_seed is not None else int(method.config.seeds[0])
        seed = (int(base) ^ 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        pilot = SimSpec(n_paths=max(2, min(cv.pilot, method.config.trajectories)), n_steps=sim.n_steps, scheme=abi.HH_SCHEME_EM,
                        vr=sim.vr, base_seed=seed)
        px, py, pd = eng.mc_path_dependent(mdl, pilot, [(abi.HH_PD_VANILLA, K, cp, 0.0, 0.0), (abi.HH_PD_BS_CONTROL, K, cp, 0.0, 0.0),
                                                        (abi.HH_PD_VANILLA_MINUS_BS, K, cp, 0.0, 1.0)], discount, 1)[0]
        vy = _var(py)
        beta = 0.5 * (_var(px) + vy - _var(pd)) / vy if vy > 0.0 else 0.0
    (res,), _ = eng.mc_path_dependent(mdl, sim, [(abi.HH_PD_VANILLA_MINUS_BS, K, cp, 0.0, beta)], discount, 1)
    sums = np.array([res.sum, res.sumsq, float(res.n)])
    if reduce is not None:
        sums = reduce(sums)
    s, q, n = sums
    mean = s / n
    var = max((q - n * mean * mean) / (n - 1), 0.0) if n > 1 else 0.0
    stats = {"kernel_ms": res.kernel_ms, "n_nonfinite": res.n_nonfinite, "n_local": sim.n_paths, "n_total": int(n),
             "beta": beta, "sigma_cv": sigma_cv, "control_price": bs}
    return api.MonteCarloSolution(prob, method, discount * mean + beta * bs, None, discount * math.sqrt(var / n), stats)
