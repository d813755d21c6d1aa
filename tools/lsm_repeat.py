import sys, math, datetime as dt
sys.path.insert(0, ".")
import hedgehog_jl_b200 as hh
eng = hh.default_engine(0)
put = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.American(), hh.Put(), hh.Spot())
bs = hh.BlackScholesInputs(dt.date(2020, 1, 1), 0.05, 100.0, 0.2)
lsm = hh.LSM(hh.MonteCarlo(hh.LognormalDynamics(), hh.BlackScholesExact(), hh.SimulationConfig(10_000_000, steps=50, base_seed=12345)), 3)
p = hh.PricingProblem(put, bs)
out = []
for i in range(30):
    s = hh.solve(p, lsm, engine=eng, stopping_info=False)
    out.append((round(s.stats["path_ms"], 3), round(s.stats["regress_ms"], 3)))
print(out)
