"""solve(::PricingProblem{American}, ::LSM) on the GPU — mirror of src/pricing_methods/least_squares_montecarlo.jl:99-136."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as abi
from . import api


def solve_lsm(prob, method, *, engine=None, shard=None, group=None, stopping_info=True, spot_paths=False):
    """`stopping_info` / `spot_paths` choose whether LSMSolution carries the per-column (tau, value) pairs and the
    (nsteps+1) x ncols spot matrix (4 GB at config C3, hence off by default; the reference always returns both)."""
    if not isinstance(prob.payoff.exercise_style, api.American):
        raise TypeError("solve(::PricingProblem, ::LSM) is defined for American exercise")
    mc = method.mc_method
    eng = engine or api.default_engine()
    shard, reduce = api._shard_and_reduce(shard, group, getattr(eng, "device", None))
    mdl = api._model_of(prob, mc)
    scheme = api._scheme_of(mc, for_lsm=True)
    if scheme == abi.HH_SCHEME_HESTON_BK:   # exact transitions between config.steps exercise dates
        from dataclasses import replace
        mc = replace(mc, bk_steps_from_config=True)
    sim = api._sim_of(mc, scheme, shard)
    m = prob.market_inputs
    T = api.yearfrac(m.referenceDate, prob.payoff.expiry)                                 # lsm.jl:104
    step_discount = api.df(m.rate, api.add_yearfrac(m.referenceDate, T / sim.n_steps))   # lsm.jl:110
    comm = None
    keep = None
    err = None
    if reduce is not None:
        from .distributed import allreduce_max_int, make_comm, peer_comm
        dev = getattr(eng, "device", None)
        peer = getattr(eng, "peers", None) == tuple(shard)
        if peer:
            comm = peer_comm(eng)              # moments exchanged in the pass kernel's tail over peer memory
            # the in-kernel wait for a peer is bounded (hh_peer_set_timeout): enter the solve together, so that ordinary skew
            # between ranks (module load, a 4 GB allocation, garbage collection) cannot be mistaken for a dead peer
            import torch.distributed as dist
            dist.barrier(group)
        else:
            comm, keep = make_comm(shard, group, dev)  # NCCL all-reduce through the hh_comm callback
        try:
            out, tau, val, paths = eng.lsm_american(mdl, sim, (prob.payoff.strike, prob.payoff.call_put()), method.degree,
                                                    step_discount, want_stopping=stopping_info, want_paths=spot_paths, comm=comm)
        except Exception as e:  # noqa: BLE001 - re-raised below, after every rank has learnt about it
            err = e
        if peer:
            # every rank learns whether ANY rank failed before the sums are reduced: no rank is left waiting in a collective
            # for one that raised (a timed-out exchange already fails on all ranks together, see hh_peer_set_timeout)
            if allreduce_max_int(0 if err is None else 1, group, dev):
                raise err if err is not None else abi.HedgehogB200Error(
                    "hh_lsm_american failed on another rank; the peer connection is dead: connect_peers() again")
        elif err is not None:
            raise err
    else:
        out, tau, val, paths = eng.lsm_american(mdl, sim, (prob.payoff.strike, prob.payoff.call_put()), method.degree,
                                                step_discount, want_stopping=stopping_info, want_paths=spot_paths, comm=comm)
    del keep
    s = np.array([out.sum, out.sumsq, float(out.n)])
    if reduce is not None:
        s = reduce(s)
    n = s[2]
    mean = s[0] / n
    var = max((s[1] - n * mean * mean) / (n - 1), 0.0) if n > 1 else 0.0
    # the reference's Vector{Tuple{Int,S}} (:112); "arrays" keeps the two numpy arrays (a Python list of 1e7 tuples costs
    # seconds to build and is only useful for small cases)
    if stopping_info == "arrays":
        info = (tau, val)
    else:
        info = list(zip(tau.tolist(), val.tolist())) if stopping_info else None
    stats = {"kernel_ms": out.kernel_ms, "path_ms": out.path_ms, "regress_ms": out.regress_ms,
             "n_dates_skipped": out.n_dates_skipped, "n_cols_local": int(out.n), "n_cols_total": int(n)}
    return api.LSMSolution(prob, method, float(mean), info, None if paths is None else paths.T,
                           float(np.sqrt(var / n)), stats)
