"""Host-side mirror of Hedgehog.jl's problem / method / solution types for the Monte Carlo path.

Same names, argument meaning and error behaviour as the reference (paths relative to the
reference checkout), so the parity tests read like the reference's own tests:

    payoff  = VanillaOption(100.0, date(2021, 1, 1), European(), Call(), Spot())       # payoffs.jl:101-140
    market  = HestonInputs(date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)  # market_inputs.jl:55-88
    prob    = PricingProblem(payoff, market)                                            # pricing_methods.jl:19-22
    method  = MonteCarlo(HestonDynamics(), EulerMaruyama(), SimulationConfig(10**6, steps=252))
    sol     = solve(prob, method)              # -> MonteCarloSolution(problem, method, price, ensemble)

The Julia version of this file is julia/HedgehogB200.jl (same scalar extraction, `ccall` instead of
ctypes). All numerics run in libhedgehog_mc.so on the GPU; nothing here computes a price.
"""
from __future__ import annotations

import datetime as _dt
import math
import secrets
from dataclasses import dataclass, field, replace
from typing import Any, Optional, Sequence

import numpy as np

from . import _abi as abi
from .engine import SimSpec, default_engine

# ---- dates: ACT/365 in millisecond ticks (date_functions.jl:1-3, 15-58, 87-89) ----------------------
MILLISECONDS_IN_DAY = 86400000
MILLISECONDS_IN_YEAR_365 = 365 * 86400 * 1000


def to_ticks(x):
    """Date -> ms since the Julia Dates epoch (Dates.date2epochdays == date.toordinal)."""
    if isinstance(x, _dt.datetime):
        d = x.date().toordinal() * MILLISECONDS_IN_DAY
        return d + ((x.hour * 60 + x.minute) * 60 + x.second) * 1000 + x.microsecond // 1000
    if isinstance(x, _dt.date):
        return x.toordinal() * MILLISECONDS_IN_DAY
    return x


def yearfrac(start, stop):
    return (to_ticks(stop) - to_ticks(start)) / MILLISECONDS_IN_YEAR_365


def add_yearfrac(t, yf):
    return to_ticks(t) + yf * MILLISECONDS_IN_YEAR_365


# ---- market inputs -------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FlatRateCurve:  # rate_curve.jl:32-35
    reference_date: int
    rate: float

    @staticmethod
    def of(rate, reference_date=_dt.date(1, 1, 1)):
        return FlatRateCurve(to_ticks(reference_date), rate)


def zero_rate(curve: FlatRateCurve, ticks):  # rate_curve.jl:185-186
    return curve.rate


def df(curve: FlatRateCurve, ticks):  # rate_curve.jl:149-150
    return math.exp(-zero_rate(curve, ticks) * yearfrac(curve.reference_date, to_ticks(ticks)))


@dataclass(frozen=True)
class FlatVolSurface:  # vol_surface.jl:73-98
    sigma: float


def get_vol(surface: FlatVolSurface, *_):
    return surface.sigma


@dataclass(frozen=True)
class BlackScholesInputs:  # market_inputs.jl:21-36
    referenceDate: int
    rate: FlatRateCurve
    spot: float
    sigma: FlatVolSurface

    def __init__(self, reference_date, rate, spot, sigma):
        ref = to_ticks(reference_date)
        object.__setattr__(self, "referenceDate", ref)
        object.__setattr__(self, "rate", rate if isinstance(rate, FlatRateCurve) else FlatRateCurve(ref, rate))
        object.__setattr__(self, "spot", spot)
        object.__setattr__(self, "sigma", sigma if isinstance(sigma, FlatVolSurface) else FlatVolSurface(sigma))


@dataclass(frozen=True)
class HestonInputs:  # market_inputs.jl:55-88 ; positional order (ref, rate, spot, V0, κ, θ, σ, ρ)
    referenceDate: int
    rate: FlatRateCurve
    spot: float
    V0: float
    kappa: float
    theta: float
    sigma: float
    rho: float

    def __init__(self, reference_date, rate, spot, V0, kappa, theta, sigma, rho):
        ref = to_ticks(reference_date)
        object.__setattr__(self, "referenceDate", ref)
        object.__setattr__(self, "rate", rate if isinstance(rate, FlatRateCurve) else FlatRateCurve(ref, rate))
        for k, v in (("spot", spot), ("V0", V0), ("kappa", kappa), ("theta", theta), ("sigma", sigma), ("rho", rho)):
            object.__setattr__(self, k, v)


# ---- payoffs (payoffs.jl) ----------------------------------------------------------------------------------
class European: pass
class American: pass
class Spot: pass
class Forward: pass


class Call:
    def __call__(self):
        return 1.0


class Put:
    def __call__(self):
        return -1.0


@dataclass(frozen=True)
class VanillaOption:  # payoffs.jl:101-140
    strike: float
    expiry: int
    exercise_style: Any
    call_put: Any
    underlying: Any

    def __init__(self, strike, expiry_date, exercise_style, call_put, underlying):
        object.__setattr__(self, "strike", strike)
        object.__setattr__(self, "expiry", to_ticks(expiry_date))
        object.__setattr__(self, "exercise_style", exercise_style)
        object.__setattr__(self, "call_put", call_put)
        object.__setattr__(self, "underlying", underlying)

    def __call__(self, spot):  # payoffs.jl:154-156 (host-side convenience; the GPU evaluates payoffs in-kernel)
        return np.maximum(self.call_put() * (np.asarray(spot) - self.strike), 0.0)


@dataclass(frozen=True)
class PricingProblem:  # pricing_methods.jl:19-22
    payoff: VanillaOption
    market_inputs: Any


# ---- method types (montecarlo.jl:15-131) ----------------------------------------------------------------------
class LognormalDynamics: pass
class HestonDynamics: pass
class NoVarianceReduction: pass
class Antithetic: pass
class QuasiRandom: pass   # HH_VR_QUASI_RANDOM: randomised van der Corput points in the one-draw exact sampler (roadmap :164)
class EulerMaruyama: pass
class HestonBroadieKaya: pass
class BlackScholesExact: pass


@dataclass(frozen=True)
class SimulationConfig:
    """montecarlo.jl:58-79. `seeds=None` draws fresh randomness like the reference, but as ONE 64-bit
    Philox key (the trajectory index goes into the counter) instead of one UInt64 per path; pass
    `base_seed` to fix it. An explicit `seeds` vector is honoured as per-trajectory keys and must
    have length >= trajectories (ArgumentError -> ValueError)."""
    trajectories: int
    steps: int = 1
    variance_reduction: Any = field(default_factory=NoVarianceReduction)  # Q2: the code default, not the docstring's
    seeds: Optional[np.ndarray] = None
    base_seed: Optional[int] = None

    def __post_init__(self):
        if self.seeds is not None:
            s = np.asarray(self.seeds, dtype=np.uint64)
            if s.shape[0] < self.trajectories:
                raise ValueError(f"Number of seeds ({s.shape[0]}) must be ≥ number of trajectories ({self.trajectories}).")
            object.__setattr__(self, "seeds", s)
        elif self.base_seed is None:
            object.__setattr__(self, "base_seed", secrets.randbits(64))


@dataclass(frozen=True)
class MonteCarlo:  # montecarlo.jl:127-131
    dynamics: Any
    strategy: Any
    config: SimulationConfig
    # execution knobs of this build (not in the reference)
    precision: str = "f64"            # "f64" | "f32" fast mode
    corr_mode: str = "cholesky"       # factor of [1 ρ; ρ 1]: "cholesky" | "sym_sqrt" | "svd" (parity-mode mapping only)
    split_step: bool = True           # StochasticDiffEq EM() default [upstream]
    q1_compat: bool = True            # marginal_law's sqrt(α) in the mean (montecarlo.jl:302)
    ensemble: bool = True             # materialise MonteCarloSolution.ensemble (terminal prices) on the host
    normals: Optional[np.ndarray] = None  # parity mode: pre-generated standard normals [path, step, comp]
    bk_steps_from_config: bool = False    # False: exact strategies ignore `steps` like the reference (Q6)
    control_variate: Any = None           # pathdep.BlackScholesControlVariate(): Heston + EulerMaruyama vanilla prices
    rng: str = "philox"                   # "philox": Philox4x32-10, 52-bit uniforms, one block per Heston step (default);
                                          # "philox64": opt-in HH_RNG_PHILOX_64, one block per TWO steps (Heston + EulerMaruyama, f64)


B200MonteCarlo = MonteCarlo


@dataclass(frozen=True)
class LSM:  # least_squares_montecarlo.jl:12-34
    mc_method: MonteCarlo
    degree: int

    def __init__(self, *args):
        if len(args) == 2:
            mc, degree = args
        else:
            dynamics, strategy, config, degree = args
            mc = MonteCarlo(dynamics, strategy, config)
        object.__setattr__(self, "mc_method", mc)
        object.__setattr__(self, "degree", int(degree))


# ---- solutions (pricing_solutions.jl:22-27, 78-84) ----------------------------------------------------------
@dataclass
class MonteCarloSolution:
    problem: PricingProblem
    method: Any
    price: float
    ensemble: Any                 # terminal prices (or a (plus, minus) tuple for antithetic); None if not requested
    std_error: float = float("nan")   # not in the reference (Q10)
    stats: dict = field(default_factory=dict)


@dataclass
class LSMSolution:
    problem: PricingProblem
    method: Any
    price: float
    stopping_info: Any            # list of (tau, value) like the reference, or None if not requested
    spot_paths: Any               # (nsteps+1, ncols) matrix like the reference, or None
    std_error: float = float("nan")
    stats: dict = field(default_factory=dict)


# ---- scalar extraction (identical for every scheme) ------------------------------------------
def corr_factor(rho: float, mode: str):
    """M with M Mᵀ = [1 ρ; ρ 1], and dM/dρ. Any factor gives the same law (heston.jl:18-20)."""
    if mode == "cholesky":
        c = math.sqrt(1 - rho * rho)
        return (1.0, 0.0, rho, c), (0.0, 0.0, 1.0, -rho / c)
    if mode == "sym_sqrt":
        p, m = math.sqrt(1 + rho), math.sqrt(1 - rho)
        a, b = (p + m) / 2, (p - m) / 2
        da, db = (1 / (2 * p) - 1 / (2 * m)) / 2, (1 / (2 * p) + 1 / (2 * m)) / 2
        return (a, b, b, a), (da, db, db, da)
    if mode == "svd":  # U * sqrt(S), U = [[1, 1], [1, -1]] / sqrt(2), S = (1+ρ, 1-ρ)
        p, m = math.sqrt((1 + rho) / 2), math.sqrt((1 - rho) / 2)
        dp, dm = 1 / (4 * p), -1 / (4 * m)
        return (p, m, p, -m), (dp, dm, dp, -dm)
    raise ValueError(f"unknown corr_mode {mode!r}")


def _model_of(prob: PricingProblem, method: MonteCarlo):
    m = prob.market_inputs
    mdl = abi.hh_model()
    mdl.flags = (abi.HH_FLAG_SPLIT_STEP if method.split_step else 0) | (abi.HH_FLAG_Q1_SQRT_MEAN if method.q1_compat else 0)
    mdl.S0 = float(m.spot)
    mdl.r = float(zero_rate(m.rate, 0.0))                         # montecarlo.jl:150 (Q3)
    mdl.T = float(yearfrac(m.referenceDate, prob.payoff.expiry))  # montecarlo.jl:147
    if isinstance(method.dynamics, LognormalDynamics):
        if not isinstance(m, BlackScholesInputs):
            raise TypeError("LognormalDynamics needs BlackScholesInputs")  # MethodError in the reference
        mdl.kind = abi.HH_MODEL_GBM
        mdl.sigma = float(get_vol(m.sigma, None, None))          # montecarlo.jl:151 (Q4)
    elif isinstance(method.dynamics, HestonDynamics):
        if not isinstance(m, HestonInputs):
            raise TypeError("HestonDynamics needs HestonInputs")
        mdl.kind = abi.HH_MODEL_HESTON
        mdl.V0, mdl.kappa, mdl.theta, mdl.xi, mdl.rho = map(float, (m.V0, m.kappa, m.theta, m.sigma, m.rho))
        (mdl.m11, mdl.m12, mdl.m21, mdl.m22), _ = corr_factor(mdl.rho, method.corr_mode)
    else:
        raise TypeError(f"unknown dynamics {method.dynamics!r}")
    return mdl


def _scheme_of(method: MonteCarlo, for_lsm: bool = False) -> int:
    s, d = method.strategy, method.dynamics
    if isinstance(s, EulerMaruyama):
        return abi.HH_SCHEME_EM
    if isinstance(s, BlackScholesExact):
        if not isinstance(d, LognormalDynamics):
            raise TypeError("BlackScholesExact needs LognormalDynamics")
        return abi.HH_SCHEME_EXACT_STEPS if for_lsm else abi.HH_SCHEME_EXACT_TERMINAL
    if isinstance(s, HestonBroadieKaya):
        if not isinstance(d, HestonDynamics):
            raise TypeError("HestonBroadieKaya needs HestonDynamics")
        return abi.HH_SCHEME_HESTON_BK
    raise TypeError(f"unknown strategy {s!r}")


def _sim_of(method: MonteCarlo, scheme: int, shard=None) -> SimSpec:
    cfg = method.config
    exact = scheme in (abi.HH_SCHEME_EXACT_TERMINAL, abi.HH_SCHEME_HESTON_BK)
    n, off = cfg.trajectories, 0
    if shard is not None:
        rank, world = shard
        lo, hi = (cfg.trajectories * rank) // world, (cfg.trajectories * (rank + 1)) // world
        n, off = hi - lo, lo
    sim = SimSpec(n_paths=n, path_offset=off, scheme=scheme, job_paths=cfg.trajectories if shard is not None else 0,
                  n_steps=cfg.steps if (not exact or method.bk_steps_from_config) else 1,
                  vr=(abi.HH_VR_ANTITHETIC if isinstance(cfg.variance_reduction, Antithetic) else
                      abi.HH_VR_QUASI_RANDOM if isinstance(cfg.variance_reduction, QuasiRandom) else abi.HH_VR_NONE),
                  precision=abi.HH_PREC_F32 if method.precision == "f32" else abi.HH_PREC_F64)
    if method.rng not in ("philox", "philox64"):
        raise ValueError(f"unknown rng {method.rng!r} (philox | philox64)")
    if method.rng == "philox64":
        if method.normals is not None:
            raise ValueError("rng='philox64' selects an in-kernel stream; it cannot be combined with pre-generated normals")
        sim.rng_mode = abi.HH_RNG_PHILOX_64
    if method.normals is not None:
        sim.rng_mode = abi.HH_RNG_NORMALS
        z = np.asarray(method.normals, dtype=np.float64)
        sim.normals = z[off:off + n]
    elif cfg.seeds is not None:
        if exact:
            sim.base_seed = int(cfg.seeds[0])   # Q6: ONE stream seeded by seeds[1] (montecarlo.jl:456)
        else:
            sim.seeds = cfg.seeds[off:off + n]  # remake(prob; seed = seeds[i]) (montecarlo.jl:331)
    else:
        sim.base_seed = int(cfg.base_seed)
    return sim


def _shard_and_reduce(shard, group, device=None):
    """(rank, world) and a function that sum-reduces a float64 vector over ranks (torch.distributed). `device`: the
    engine's CUDA device, where the NCCL collectives of this rank run."""
    if shard is not None:
        return shard, None
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None, None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None, None
    from .distributed import allreduce_sum_f64
    return (dist.get_rank(group), dist.get_world_size(group)), (lambda v: allreduce_sum_f64(v, group, device))


# ---- solve: European Monte Carlo (montecarlo.jl:478-493) ---------------------------------------------------------
def _solve_european(prob, method, engine, shard, group, strikes=None):
    if not isinstance(prob.payoff.underlying, Spot):
        raise TypeError("MonteCarlo prices options on the Spot underlying")  # dispatch: VanillaOption{..,European,C,Spot}
    eng = engine or default_engine()
    shard, reduce = _shard_and_reduce(shard, group, getattr(eng, "device", None))
    mdl = _model_of(prob, method)
    sim = _sim_of(method, _scheme_of(method), shard)
    cp = prob.payoff.call_put()
    payoffs = [(prob.payoff.strike, cp)] if strikes is None else [(float(k), cp) for k in strikes]
    discount = df(prob.market_inputs.rate, prob.payoff.expiry)     # montecarlo.jl:489
    want_terminal = bool(method.ensemble)
    results, terminal = eng.mc_european(mdl, sim, payoffs, discount, want_terminal)
    sums = np.array([[r.sum, r.sumsq, float(r.n)] for r in results], dtype=np.float64)
    if reduce is not None:
        sums = reduce(sums)
    out = []
    for k, r in enumerate(results):
        s, q, n = sums[k]
        mean = s / n
        var = max((q - n * mean * mean) / (n - 1), 0.0) if n > 1 else 0.0
        out.append((discount * mean, discount * math.sqrt(var / n)))   # montecarlo.jl:490
    ens = None
    if want_terminal:
        N = sim.n_paths
        ens = (terminal[:N], terminal[N:]) if sim.vr == abi.HH_VR_ANTITHETIC else terminal  # montecarlo.jl:398-402
    stats = {"kernel_ms": results[0].kernel_ms, "n_nonfinite": results[0].n_nonfinite,
             "n_fallback": results[0].n_fallback, "n_local": sim.n_paths, "n_total": int(sums[0][2])}
    return out, ens, stats


def solve(prob, method, *args, engine=None, shard=None, group=None, **kw):
    """Hedgehog.solve for the Monte Carlo path. Dispatches like the reference:
    PricingProblem × MonteCarlo (European), PricingProblem × LSM (American), and the Greek problems."""
    from . import greeks as _g
    from . import calibration as _c
    if isinstance(prob, _c.CalibrationProblem):
        return _c.solve_calibration(prob, method, *args, engine=engine, shard=shard, group=group, **kw)
    if isinstance(prob, (_g.GreekProblem, _g.BatchGreekProblem, _g.SecondOrderGreekProblem)):
        return _g.solve_greek(prob, method, *args, engine=engine, shard=shard, group=group, **kw)
    from . import pathdep as _pd
    if isinstance(prob, BasketPricingProblem):
        if isinstance(method, MonteCarlo) and any(_pd.is_path_payoff(p) for p in prob.payoffs):
            res, stats = _pd.solve_path_dependent(prob.payoffs, prob.market_inputs, method, engine=engine, shard=shard, group=group)
            return [MonteCarloSolution(PricingProblem(p, prob.market_inputs), method, price, None, se, stats)
                    for p, (price, se) in zip(prob.payoffs, res)]
        return _solve_basket(prob, method, engine, shard, group)
    if isinstance(method, MonteCarlo) and _pd.is_path_payoff(prob.payoff):
        ((price, se),), stats = _pd.solve_path_dependent([prob.payoff], prob.market_inputs, method, engine=engine, shard=shard, group=group)
        return MonteCarloSolution(prob, method, price, None, se, stats)
    if isinstance(method, LSM):
        from .lsm import solve_lsm
        return solve_lsm(prob, method, engine=engine, shard=shard, group=group, **kw)
    if isinstance(method, MonteCarlo):
        if not isinstance(prob.payoff.exercise_style, European):
            raise TypeError("solve(::PricingProblem, ::MonteCarlo) is defined for European exercise; use LSM")
        if method.control_variate is not None:
            return _pd.solve_bs_control(prob, method, engine=engine, shard=shard, group=group)
        (price_se,), ens, stats = _solve_european(prob, method, engine, shard, group)
        return MonteCarloSolution(prob, method, price_se[0], ens, price_se[1], stats)
    raise TypeError(f"no B200 solve for method {type(method).__name__}")


# ---- "next" row N1: a strike grid priced on ONE simulation (src/calibration/basket.jl:10-38) ---------------------
@dataclass(frozen=True)
class BasketPricingProblem:
    payoffs: Sequence[VanillaOption]
    market_inputs: Any


def _solve_basket(bprob, method, engine, shard, group):
    """The reference loops solve() per payoff (basket.jl:35-38); payoffs sharing expiry and call/put are
    priced here in one launch on common paths."""
    groups: dict = {}
    for i, p in enumerate(bprob.payoffs):
        groups.setdefault((p.expiry, type(p.call_put), type(p.exercise_style)), []).append(i)
    sols = [None] * len(bprob.payoffs)
    for (_, _, ex), idxs in groups.items():
        if ex is not European or not isinstance(method, MonteCarlo):
            for i in idxs:
                sols[i] = solve(PricingProblem(bprob.payoffs[i], bprob.market_inputs), method, engine=engine, shard=shard, group=group)
            continue
        for c0 in range(0, len(idxs), 256):
            chunk = idxs[c0:c0 + 256]
            p0 = bprob.payoffs[chunk[0]]
            m2 = replace(method, ensemble=False)
            res, _, stats = _solve_european(PricingProblem(p0, bprob.market_inputs), m2, engine, shard, group,
                                            strikes=[bprob.payoffs[i].strike for i in chunk])
            for i, (price, se) in zip(chunk, res):
                sols[i] = MonteCarloSolution(PricingProblem(bprob.payoffs[i], bprob.market_inputs), method, price, None, se, stats)
    return sols
