"""Worker of tests/test_multirank_gloo.py: one rank of a world_size-2 gloo job on the CPU.

The host layer under test is the product's (api.solve sharding, distributed.allreduce_sum_f64, distributed.make_comm);
the arithmetic engine injected into it is the CPU oracle, because there is no GPU in the CPU test tier."""
import datetime as dt
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

import hedgehog_jl_b200 as hh  # noqa: E402
from hedgehog_jl_b200 import distributed as hd  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = O.OracleEngine(threads=2)
    res = {}

    # European Heston EM: every rank simulates its block of the global trajectory index, [sum, sumsq, n] allreduced
    N = 20_001  # odd: uneven shards
    payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
    market = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    prob = hh.PricingProblem(payoff, market)
    for name, vr in (("novr", hh.NoVarianceReduction()), ("anti", hh.Antithetic())):
        method = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(),
                               hh.SimulationConfig(N, steps=16, variance_reduction=vr, base_seed=7), ensemble=False)
        sol = hh.solve(prob, method, engine=eng)                      # sharded: picks (rank, world) up from torch.distributed
        one = hh.solve(prob, method, engine=eng, shard=(0, 1))        # the whole job on this rank, no collective
        res[name] = {"sharded": [sol.price, sol.std_error, sol.stats["n_local"], sol.stats["n_total"]],
                     "single": [one.price, one.std_error]}

    # explicit per-trajectory seeds are sharded with the trajectories
    seeds = np.random.Generator(np.random.Philox(3)).integers(0, 2**63, size=N, dtype=np.uint64)
    method = hh.MonteCarlo(hh.LognormalDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(N, steps=8, seeds=seeds), ensemble=False)
    bs = hh.PricingProblem(payoff, hh.BlackScholesInputs(dt.date(2020, 1, 1), 0.05, 100.0, 0.2))
    res["seeds"] = {"sharded": hh.solve(bs, method, engine=eng).price, "single": hh.solve(bs, method, engine=eng, shard=(0, 1)).price}

    # a strike grid on common paths (BasketPricingProblem) and batch Greeks (tangent sums) reduce the same way
    basket = hh.BasketPricingProblem([hh.VanillaOption(k, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())
                                      for k in (80.0, 100.0, 120.0)], market)
    method = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(N, steps=16, base_seed=11), ensemble=False)
    res["basket"] = {"sharded": [s.price for s in hh.solve(basket, method, engine=eng)],
                     "single": [s.price for s in hh.solve(basket, method, engine=eng, shard=(0, 1))]}
    lenses = [hh.SpotLens(), hh.ZeroRateSpineLens(1), hh.optic("market_inputs.V0"), hh.optic("market_inputs.rho")]
    g = hh.solve(hh.BatchGreekProblem(prob, lenses), hh.ForwardAD(), method, engine=eng)
    g1 = hh.solve(hh.BatchGreekProblem(prob, lenses), hh.ForwardAD(), method, engine=eng, shard=(0, 1))
    res["greeks"] = {"sharded": [float(g[l]) for l in lenses], "single": [float(g1[l]) for l in lenses]}

    # path-dependent payoffs on common trajectories: the same shard + [sum, sumsq, n] reduction
    exp = dt.date(2020, 12, 31)
    pd_basket = hh.BasketPricingProblem([hh.AsianOption(100.0, exp, hh.Call(), monitoring=hh.Monitoring(4)),
                                         hh.BarrierOption(100.0, 120.0, exp, hh.Call(), hh.Up(), hh.KnockOut(), monitoring=hh.Monitoring(4)),
                                         hh.DigitalOption(95.0, exp, hh.Put(), monitoring=hh.Monitoring(4))], market)
    res["pathdep"] = {"sharded": [[s.price, s.std_error] for s in hh.solve(pd_basket, method, engine=eng)],
                      "single": [[s.price, s.std_error] for s in hh.solve(pd_basket, method, engine=eng, shard=(0, 1))]}

    # Black-Scholes control variate: the pilot (beta) is the same launch on every rank, the controlled sums are sharded
    cvm = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), hh.SimulationConfig(N, steps=16, base_seed=11), ensemble=False,
                        control_variate=hh.BlackScholesControlVariate(pilot=4000))
    a, b = hh.solve(prob, cvm, engine=eng), hh.solve(prob, cvm, engine=eng, shard=(0, 1))
    res["bs_control"] = {"sharded": [a.price, a.std_error, a.stats["beta"]], "single": [b.price, b.std_error, b.stats["beta"]]}

    # the hh_comm callback the LSM driver calls between pass and fit: in-place sum-allreduce of a buffer
    comm, keep = hd.make_comm((rank, world))
    buf = np.arange(12, dtype=np.float64) * (rank + 1)
    rc = comm.allreduce_sum_f64(None, buf.ctypes.data, buf.size, None)
    res["comm"] = {"rc": int(rc), "buf": buf.tolist(), "rank": int(comm.rank), "world": int(comm.world)}
    res["allreduce"] = hd.allreduce_sum_f64(np.array([[1.0 + rank, 2.0], [3.0, 4.0 * rank]])).tolist()

    dist.barrier()
    with open(f"{out_path}.{rank}", "w") as f:
        json.dump(res, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
