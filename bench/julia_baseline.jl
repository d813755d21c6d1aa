# julia_baseline.jl — times the REAL reference (Hedgehog.jl's CPU MonteCarlo / LSM path) on the BASELINE.json shapes at
# reduced path counts (the reference stores every step of every trajectory, BASELINE.md §2). Not runnable in the build
# image (no Julia); run it wherever Julia >= 1.10 and Hedgehog's dependencies are installed:
#     JULIA_NUM_THREADS=$(nproc) julia --project=/path/to/Hedgehog.jl bench/julia_baseline.jl
# Prints one JSON line per configuration: rate in the metric's unit, thread count, sample size.
using Hedgehog, Dates, BenchmarkTools, Printf, Random

ref = Date(2020, 1, 1)
expiry = ref + Day(365)                      # T = 1 exactly (ACT/365)
call(K) = VanillaOption(K, expiry, European(), Call(), Spot())
heston = HestonInputs(ref, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)   # test/agreement/montecarlo_heston.jl:13-22
bs = BlackScholesInputs(ref, 0.05, 100.0, 0.2)
nthreads = Threads.nthreads()
line(name, unit, rate, sample) = @printf("{\"config\": \"%s\", \"unit\": \"%s\", \"value\": %.6e, \"threads\": %d, \"sample\": \"%s\"}\n",
                                         name, unit, rate, nthreads, sample)

# C2: Heston Euler-Maruyama, 252 steps
let N = 100_000, steps = 252
    m = MonteCarlo(HestonDynamics(), EulerMaruyama(), SimulationConfig(N; steps = steps, seeds = rand(MersenneTwister(42), UInt64, N)))
    t = @belapsed solve($(PricingProblem(call(100.0), heston)), $m) samples = 3 evals = 1
    line("C2 Heston EM f64", "path-steps/s", N * steps / t, "$N paths x $steps steps")
end
# C1: GBM exact, one step
let N = 1_000_000
    m = MonteCarlo(LognormalDynamics(), BlackScholesExact(), SimulationConfig(N; steps = 1, seeds = rand(MersenneTwister(42), UInt64, N)))
    t = @belapsed solve($(PricingProblem(call(100.0), bs)), $m) samples = 5 evals = 1
    line("C1 GBM exact", "paths/s", N / t, "$N paths x 1 step")
end
# C3: American put, LSM, 50 dates, degree 3
let N = 100_000, steps = 50
    put = VanillaOption(100.0, expiry, American(), Put(), Spot())
    m = LSM(LognormalDynamics(), BlackScholesExact(), SimulationConfig(N; steps = steps, seeds = rand(MersenneTwister(12345), UInt64, N)), 3)
    t = @belapsed solve($(PricingProblem(put, bs)), $m) samples = 3 evals = 1
    line("C3 LSM", "path-dates/s", N * steps / t, "$N paths x $steps dates, degree 3")
end
# C4: Heston Broadie-Kaya (one transition to expiry: the reference's exact strategies ignore `steps`)
let N = 10_000
    m = MonteCarlo(HestonDynamics(), HestonBroadieKaya(), SimulationConfig(N; steps = 1, seeds = rand(MersenneTwister(42), UInt64, N)))
    t = @belapsed solve($(PricingProblem(call(100.0), heston)), $m) samples = 3 evals = 1
    line("C4 Heston BK", "transitions/s", N / t, "$N transitions")
end
# C5: batch Greeks by ForwardDiff through the simulation (one simulation per lens)
let N = 10_000, steps = 252
    m = MonteCarlo(HestonDynamics(), EulerMaruyama(), SimulationConfig(N; steps = steps, seeds = rand(MersenneTwister(42), UInt64, N)))
    lenses = (SpotLens(), ZeroRateSpineLens(1))
    g = BatchGreekProblem(PricingProblem(call(100.0), heston), lenses)
    t = @belapsed solve($g, ForwardAD(), $m) samples = 3 evals = 1
    line("C5 batch Greeks (2 lenses, 1 strike)", "path-steps/s", N * steps / t, "$N paths x $steps steps x $(length(lenses)) lenses")
end
