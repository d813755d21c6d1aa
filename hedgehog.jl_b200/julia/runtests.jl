# runtests.jl — the reference's own agreement tests, re-run through the B200 method types.
#
#   julia --project=<environment with Hedgehog> hedgehog.jl_b200/julia/runtests.jl        (needs a B200 and libhedgehog_mc.so)
#
# Every testset below is the reference's test of the same name (file:line cited) with `MonteCarlo(...)` / `LSM(...)` wrapped
# in `B200MonteCarlo(...)` / `B200LSM(...)`: same parameters, same path counts, same tolerances, same oracles
# (BlackScholesAnalytic, CarrMadan, CoxRossRubinsteinMethod) — computed by Hedgehog itself. This image has no Julia toolchain,
# so the file has not been executed here; tests/test_oracle_mc.py and tests/test_gpu_*.py run the same set-ups through the
# Python twin of the host layer.
using Test, Dates, Random, Statistics
using Hedgehog, Accessors
include(joinpath(@__DIR__, "HedgehogB200.jl"))
using .HedgehogB200

ref_date = Date(2020, 1, 1)

@testset "Black-Scholes Monte Carlo (test/agreement/montecarlo_black_scholes.jl:8-169)" begin
    payoff = VanillaOption(100.0, ref_date + Year(1), European(), Call(), Spot())
    prob = PricingProblem(payoff, BlackScholesInputs(ref_date, 0.05, 100.0, 0.2))
    analytic = Hedgehog.solve(prob, BlackScholesAnalytic()).price
    for strategy in (BlackScholesExact(), EulerMaruyama()), vr in (Hedgehog.NoVarianceReduction(), Antithetic())
        prices = map(1:5) do trial
            seeds = rand(MersenneTwister(42 + trial), UInt64, 10_000)                     # :60-70
            cfg = SimulationConfig(10_000; steps = 1, seeds = seeds, variance_reduction = vr)
            Hedgehog.solve(prob, B200MonteCarlo(LognormalDynamics(), strategy, cfg)).price
        end
        @test isapprox(mean(prices), analytic; rtol = 0.02)                               # :130
    end
end

@testset "Heston Monte Carlo vs Carr-Madan (test/agreement/montecarlo_heston.jl:8-144, 208-253)" begin
    payoff = VanillaOption(100.0, ref_date + Year(1), European(), Call(), Spot())
    prob = PricingProblem(payoff, HestonInputs(ref_date, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7))
    reference = Hedgehog.solve(prob, CarrMadan(1.0, 32.0, HestonDynamics())).price       # :47
    em = B200MonteCarlo(HestonDynamics(), EulerMaruyama(), SimulationConfig(5000; seeds = nothing))
    @test isapprox(Hedgehog.solve(prob, em).price, reference; rtol = 0.05)                # :116 (steps = 1, SURVEY Q9)
    em252 = B200MonteCarlo(HestonDynamics(), EulerMaruyama(), SimulationConfig(1_000_000; steps = 252); base_seed = 42, ensemble = false)
    @test isapprox(Hedgehog.solve(prob, em252).price, reference; rtol = 5e-3)
    bk = B200MonteCarlo(HestonDynamics(), HestonBroadieKaya(), SimulationConfig(10_000; seeds = nothing))
    @test isapprox(Hedgehog.solve(prob, bk).price, reference; rtol = 2e-2)                # :252
    # Q5: Antithetic + HestonBroadieKaya is a MethodError in the reference (montecarlo.jl:387)
    @test_throws MethodError Hedgehog.solve(prob, B200MonteCarlo(HestonDynamics(), HestonBroadieKaya(),
                                                                 SimulationConfig(100; variance_reduction = Antithetic())))
    # montecarlo.jl:65-66
    @test_throws ArgumentError SimulationConfig(100; seeds = UInt64[1, 2, 3])
end

@testset "LSM American put vs CRR (test/agreement/american_options.jl:9-52)" begin
    payoff = VanillaOption(100.0, ref_date + Day(365), American(), Put(), Spot())
    prob = PricingProblem(payoff, BlackScholesInputs(ref_date, 0.05, 100.0, 0.2))
    crr = Hedgehog.solve(prob, CoxRossRubinsteinMethod(1000)).price
    seeds = rand(Xoshiro(12345), UInt64, 50_000)
    cfg = SimulationConfig(50_000; steps = 100, seeds = seeds, variance_reduction = Antithetic())
    sol = Hedgehog.solve(prob, B200LSM(LognormalDynamics(), BlackScholesExact(), cfg, 5))
    @test isapprox(sol.price, crr; rtol = 0.02)                                           # :49
    @test size(sol.spot_paths) == (101, 100_000)                                          # least_squares_montecarlo.jl:50, 70-85
    @test length(sol.stopping_info) == 100_000 && all(1 <= t <= 100 for (t, _) in sol.stopping_info)
    quiet = Hedgehog.solve(prob, B200LSM(LognormalDynamics(), BlackScholesExact(), cfg, 5; stopping_info = false, spot_paths = false))
    @test quiet.price == sol.price && isempty(quiet.stopping_info) && isempty(quiet.spot_paths)
end

@testset "Monte Carlo vs analytic Greeks (test/agreement/greeks_agreement.jl:170-241)" begin
    payoff = VanillaOption(1.0, Date(2021, 1, 1), European(), Call(), Spot())
    prob = PricingProblem(payoff, BlackScholesInputs(ref_date, 0.03, 1.0, 1.0))
    vol_lens, spot_lens, rate_lens = VolLens(1, 1), @optic(_.market_inputs.spot), ZeroRateSpineLens(1)
    seeds = rand(Xoshiro(42), 1:10^9, 100_000)
    mc = B200MonteCarlo(LognormalDynamics(), BlackScholesExact(), SimulationConfig(100_000; seeds = seeds))
    an = BlackScholesAnalytic()
    @test isapprox(Hedgehog.solve(prob, mc).price, Hedgehog.solve(prob, an).price; rtol = 3e-2)           # :211
    # ForwardAD through Hedgehog's GENERIC solve (greeks_problem.jl:249-262): the Dual reaches solve(::PricingProblem,
    # ::B200MonteCarlo), which unpacks it into one tangent direction of the kernel
    g = Hedgehog.solve(GreekProblem(prob, spot_lens), ForwardAD(), mc)
    @test g isa NamedTuple && haskey(g, :greek)                                                            # :261
    @test isapprox(g.greek, Hedgehog.solve(GreekProblem(prob, spot_lens), AnalyticGreek(), an).greek; rtol = 3e-2)   # :219
    gp2 = SecondOrderGreekProblem(prob, spot_lens, spot_lens)
    @test isapprox(Hedgehog.solve(gp2, FiniteDifference(1e-1), mc).greek, Hedgehog.solve(gp2, AnalyticGreek(), an).greek; rtol = 2e-1)   # :224
    @test isapprox(Hedgehog.solve(GreekProblem(prob, vol_lens), ForwardAD(), mc).greek,
                   Hedgehog.solve(GreekProblem(prob, vol_lens), AnalyticGreek(), an).greek; rtol = 1e-1)                                 # :232
    @test isapprox(Hedgehog.solve(GreekProblem(prob, rate_lens), ForwardAD(), mc).greek,
                   Hedgehog.solve(GreekProblem(prob, rate_lens), ForwardAD(), an).greek; rtol = 1e-2)                                    # :240
    # BatchGreekProblem: ONE simulation for all lenses, same Dict as the reference's loop (:559-568)
    batch = Hedgehog.solve(BatchGreekProblem(prob, (spot_lens, vol_lens, rate_lens)), ForwardAD(), mc)
    @test isapprox(batch[spot_lens], g.greek; rtol = 1e-12)
    @test_throws ArgumentError Hedgehog.solve(gp2, ForwardAD(), mc)      # nested duals: the message names FiniteDifference
end

@testset "Heston Greeks and a basket on common trajectories" begin
    inputs = HestonInputs(ref_date, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    expiry = ref_date + Day(365)
    mc = B200MonteCarlo(HestonDynamics(), EulerMaruyama(), SimulationConfig(2_000_000; steps = 252); base_seed = 7, ensemble = false)
    payoffs = [VanillaOption(k, expiry, European(), Call(), Spot()) for k in 80.0:5.0:120.0]
    basket = Hedgehog.solve(BasketPricingProblem(payoffs, inputs), mc)                   # one launch for the nine strikes
    for (p, s) in zip(payoffs, basket.solutions)
        cm = Hedgehog.solve(PricingProblem(p, inputs), CarrMadan(1.0, 32.0, HestonDynamics())).price
        @test isapprox(s.price, cm; rtol = 1e-2)
    end
    prob = PricingProblem(payoffs[5], inputs)
    lenses = (@optic(_.market_inputs.spot), @optic(_.market_inputs.V0), @optic(_.market_inputs.κ),
              @optic(_.market_inputs.θ), @optic(_.market_inputs.σ), @optic(_.market_inputs.ρ), ZeroRateSpineLens(1))
    greeks = Hedgehog.solve(BatchGreekProblem(prob, lenses), ForwardAD(), mc)
    cmm = CarrMadan(1.0, 32.0, HestonDynamics())
    for lens in lenses
        fd = Hedgehog.solve(GreekProblem(prob, lens), FiniteDifference(1e-4), cmm).greek
        @test isapprox(greeks[lens], fd; rtol = 5e-2, atol = 2e-3)
    end
end
