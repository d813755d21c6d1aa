# dump_reference.jl — runs the REAL Hedgehog.jl package on small fixed inputs and writes what it computes, so that the CPU
# oracle (oracle/) and the CUDA kernels can be pinned to the package itself instead of to a restatement of it.
#
#   julia --project=<environment with Hedgehog and its dependencies> tools/dump_reference.jl [path/to/Hedgehog.jl] [outdir]
#
# (default outdir: tests/golden/reference_dump). tests/test_reference_golden.py consumes the directory when it exists and
# skips loudly when it does not. This image has no Julia toolchain, so the dump cannot be produced here (SURVEY 8c:
# "parity unpinned"); the script is written against the reference sources by reading, every call cites its file:line.
#
# What is dumped, and through which door of the reference:
#   heston_em / gbm_em   Euler-Maruyama trajectories on KNOWN Brownian increments. The reference's own hook for that is
#                        `remake(prob; noise = NoiseGrid(t, W))` (src/pricing_methods/montecarlo.jl:252-263, its antithetic
#                        pass); here W is the cumulative sum of increments built from dumped normals. The SDE comes from
#                        `Hedgehog.sde_problem` (:166-202), the solver call is the one of `simulate_paths` (:349-351):
#                        `StochasticDiffEq.solve(prob, EM(); dt)`. Settles: split-step EM, full truncation, sqrt placement.
#   heston_em_corr       the same Heston problem on the reference's OWN `CorrelatedWienerProcess` (heston.jl:18-20) fed by a
#                        replay RNG that returns the dumped standard normals: settles the factor M of [1 rho; rho 1].
#                        (Optional: needs DiffEqNoiseProcess to accept a custom AbstractRNG; recorded as unavailable if not.)
#   gbm_exact_terminal   `Hedgehog.solve(prob, MonteCarlo(LognormalDynamics(), BlackScholesExact(), cfg))` (:454-493) at
#                        T = 366/365, where the sqrt(alpha)-in-the-mean quirk (:302, SURVEY Q1) is visible; the normals are
#                        re-drawn from the same `Xoshiro(seeds[1])` (:456).
#   lsm / lsm_antithetic `Hedgehog.solve(prob, LSM(...))` (least_squares_montecarlo.jl:99-136): spot grid, stopping_info,
#                        price. The consumer feeds the dumped grid to its own backward induction (Polynomials.fit = QR).
#   gbm_exact_steps      one `GeometricBrownianMotionProcess` trajectory set on a replay RNG: settles S += S (e^x - 1).
#                        (Optional, like heston_em_corr.)
#   bk_*                 Broadie-Kaya deterministic pieces for three parameter sets: `HestonCFIterator` (heston.jl:150-176),
#                        `evaluate_chf` (:184-212), `moments_from_cf` (sample_from_cf.jl:50-64), `cdf_from_cf` (:75-96) and
#                        `inverse_cdf` (:105-135) with the arguments `sample_from_cf` (:27-41) builds.
#
# Format: one little-endian raw file per array (`<name>.f64` / `<name>.i64`, Julia column-major) and `manifest.txt` with
# lines `array <name> <dtype> <dim1> <dim2> ...` and `meta <key> <value>`.
using Pkg
length(ARGS) >= 1 && isdir(ARGS[1]) && Pkg.develop(path = ARGS[1])
using Hedgehog, Dates, Random, Statistics, LinearAlgebra
using StochasticDiffEq, DiffEqNoiseProcess
import SciMLBase

const OUT = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden", "reference_dump")
mkpath(OUT)
const MANIFEST = String[]

function dump(name::AbstractString, a::AbstractArray{Float64})
    open(joinpath(OUT, name * ".f64"), "w") do io
        write(io, htol.(vec(collect(a))))
    end
    push!(MANIFEST, "array $name f64 " * join(size(a), " "))
end
function dump(name::AbstractString, a::AbstractArray{<:Integer})
    open(joinpath(OUT, name * ".i64"), "w") do io
        write(io, htol.(Int64.(vec(collect(a)))))
    end
    push!(MANIFEST, "array $name i64 " * join(size(a), " "))
end
meta(key, value) = push!(MANIFEST, "meta $key $value")

meta("hedgehog_version", string(pkgversion(Hedgehog)))
meta("julia_version", string(VERSION))
for dep in ("StochasticDiffEq", "DiffEqNoiseProcess", "Distributions", "SpecialFunctions", "Roots", "Polynomials", "ForwardDiff")
    for (uuid, info) in Pkg.dependencies()
        info.name == dep && meta("dep_" * dep, string(info.version))
    end
end

# ---- a RNG that replays a fixed list of standard normals (and uniforms) ---------------------------------------------------
mutable struct ReplayRNG <: Random.AbstractRNG
    z::Vector{Float64}
    pos::Int
end
ReplayRNG(z::AbstractVector{Float64}) = ReplayRNG(collect(z), 0)
next!(r::ReplayRNG) = (r.pos += 1; r.z[r.pos])
Random.randn(r::ReplayRNG, ::Type{Float64}) = next!(r)
Random.randn(r::ReplayRNG) = next!(r)
Random.randn!(r::ReplayRNG, a::AbstractArray{Float64}) = (for i in eachindex(a); a[i] = next!(r); end; a)
Random.rand(r::ReplayRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) = next!(r)
Random.seed!(r::ReplayRNG, args...) = r     # solve(...; seed) must not rewind or reseed the replay
Base.copy(r::ReplayRNG) = ReplayRNG(copy(r.z), r.pos)

ref_date = Date(2020, 1, 1)

# ---- 1. Heston Euler-Maruyama on known increments (config C2 parameters, test/agreement/montecarlo_heston.jl:13-22) --------
function dump_em(tag, prob, method, ncomp::Int, npaths::Int, steps::Int, M::Matrix{Float64})
    sde = Hedgehog.sde_problem(prob, method)                              # montecarlo.jl:166-202, 236-241
    T = sde.tspan[2]
    dt = T / steps                                                        # :349
    tgrid = collect(range(0.0, T; length = steps + 1))
    rng = Xoshiro(20261018)
    Z = randn(rng, ncomp, steps, npaths)                                  # the normals, dumped
    W = zeros(ncomp, steps + 1, npaths)                                   # the grid handed to NoiseGrid, dumped
    uT = zeros(ncomp, npaths)
    for p in 1:npaths
        for n in 1:steps
            W[:, n+1, p] = W[:, n, p] + sqrt(dt) * (M * Z[:, n, p])
        end
        Wp = ncomp == 1 ? [W[1, n, p] for n in 1:steps+1] : [W[:, n, p] for n in 1:steps+1]
        noise = NoiseGrid(tgrid, Wp)                                      # the reference's hook, :258
        sol = StochasticDiffEq.solve(SciMLBase.remake(sde; noise = noise), EM(); dt = dt)   # :259, :351
        last_u = last(sol.u)
        uT[:, p] .= ncomp == 1 ? [last_u[1]] : last_u
    end
    ST = exp.(uT[1, :])                                                   # final_sample, :398
    payoffs = prob.payoff.(ST)                                            # reduce_payoffs, :428
    price = df(prob.market_inputs.rate, prob.payoff.expiry) * mean(payoffs)   # :489-490
    dump("$(tag)_Z", Z); dump("$(tag)_W", W); dump("$(tag)_uT", uT); dump("$(tag)_ST", ST)
    meta("$(tag)_price", repr(price)); meta("$(tag)_T", repr(T)); meta("$(tag)_steps", steps)
    return sde, Z, T
end

heston_in = HestonInputs(ref_date, 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
call_1y = VanillaOption(100.0, ref_date + Day(365), European(), Call(), Spot())
hprob = PricingProblem(call_1y, heston_in)
hmethod = MonteCarlo(HestonDynamics(), EulerMaruyama(), SimulationConfig(64; steps = 16))
rho = -0.7
Mchol = [1.0 0.0; rho sqrt(1 - rho^2)]
meta("heston_em_params", "S0=100 r=0.03 V0=0.04 kappa=2 theta=0.04 xi=0.3 rho=-0.7 K=100 cp=1")
meta("heston_em_M", "cholesky (only used to BUILD W; consumers read W)")
hsde, hZ, hT = dump_em("heston_em", hprob, hmethod, 2, 64, 16, Mchol)
# a violent parameter set: variance hits zero often, so full truncation and the sqrt placement matter on most steps
heston_wild = HestonInputs(ref_date, 0.03, 100.0, 0.01, 0.5, 0.01, 1.0, -0.9)
meta("heston_em_wild_params", "S0=100 r=0.03 V0=0.01 kappa=0.5 theta=0.01 xi=1.0 rho=-0.9 K=100 cp=1")
dump_em("heston_em_wild", PricingProblem(call_1y, heston_wild), hmethod, 2, 64, 16, [1.0 0.0; -0.9 sqrt(1 - 0.81)])

bs_in = BlackScholesInputs(ref_date, 0.05, 100.0, 0.2)
meta("gbm_em_params", "S0=100 r=0.05 sigma=0.2 K=100 cp=1")
dump_em("gbm_em", PricingProblem(call_1y, bs_in), MonteCarlo(LognormalDynamics(), EulerMaruyama(), SimulationConfig(64; steps = 16)),
        1, 64, 16, fill(1.0, 1, 1))

# ---- 1b. the reference's own correlated Wiener process on replayed normals (optional) --------------------------------------
try
    steps, npaths = 16, 64
    dt = hT / steps
    uT = zeros(2, npaths)
    for p in 1:npaths
        rng = ReplayRNG(vec(hZ[:, :, p]))                                 # (Z1, Z2) of step 1, then of step 2, ...
        noise = CorrelatedWienerProcess([1 rho; rho 1], 0.0, zeros(2); rng = rng, reseed = false)   # heston.jl:18-20
        sol = StochasticDiffEq.solve(SciMLBase.remake(hsde; noise = noise), EM(); dt = dt)
        uT[:, p] .= last(sol.u)
    end
    dump("heston_em_corr_uT", uT)
    meta("heston_em_corr", "ok")
catch err
    meta("heston_em_corr", "unavailable: " * replace(sprint(showerror, err), '\n' => ' ')[1:min(end, 200)])
end

# ---- 2. exact GBM terminal law through solve (C1), T = 366/365 so that Q1 is visible ------------------------------------------
call_leap = VanillaOption(100.0, ref_date + Year(1), European(), Call(), Spot())     # 2020 is a leap year: 366 days
seeds = UInt64.(42:42+999)
sol1 = Hedgehog.solve(PricingProblem(call_leap, bs_in),
                      MonteCarlo(LognormalDynamics(), BlackScholesExact(), SimulationConfig(1000; seeds = seeds)))
dump("gbm_exact_terminal_ST", Float64.(sol1.ensemble))
let r1 = Xoshiro(seeds[1])                                               # montecarlo.jl:456-457: ONE Xoshiro(seeds[1]) stream,
    dump("gbm_exact_terminal_Z", [randn(r1) for _ in 1:1000])            # drawn one scalar at a time like rand(rng, ::Normal)
end
meta("gbm_exact_terminal_price", repr(sol1.price))
meta("gbm_exact_terminal_T", repr(yearfrac(bs_in.referenceDate, call_leap.expiry)))
sol1a = Hedgehog.solve(PricingProblem(call_leap, bs_in),
                       MonteCarlo(LognormalDynamics(), BlackScholesExact(),
                                  SimulationConfig(1000; seeds = seeds, variance_reduction = Antithetic())))
dump("gbm_exact_terminal_anti_plus", Float64.(sol1a.ensemble[1]))
dump("gbm_exact_terminal_anti_minus", Float64.(sol1a.ensemble[2]))
meta("gbm_exact_terminal_anti_price", repr(sol1a.price))

# ---- 3. Longstaff-Schwartz through solve (C3 parameters, test/agreement/american_options.jl:11-16) -------------------------
put_1y = VanillaOption(100.0, ref_date + Day(365), American(), Put(), Spot())
for (tag, vr) in (("lsm", Hedgehog.NoVarianceReduction()), ("lsm_antithetic", Antithetic()))
    cfg = SimulationConfig(256; steps = 10, seeds = UInt64.(12345:12345+255), variance_reduction = vr)
    s = Hedgehog.solve(PricingProblem(put_1y, bs_in), LSM(LognormalDynamics(), BlackScholesExact(), cfg, 3))
    dump("$(tag)_spot_paths", Float64.(s.spot_paths))                      # (steps + 1) x columns, :50, :135
    dump("$(tag)_tau", [t for (t, _) in s.stopping_info])
    dump("$(tag)_value", Float64[v for (_, v) in s.stopping_info])
    meta("$(tag)_price", repr(s.price)); meta("$(tag)_degree", 3)
    meta("$(tag)_params", "S0=100 r=0.05 sigma=0.2 K=100 cp=-1 T=1 steps=10")
end

# ---- 3b. GeometricBrownianMotionProcess on replayed normals (optional): the S-space increment form ---------------------------
try
    steps, npaths = 10, 32
    Zg = randn(Xoshiro(7), steps, npaths)
    S = zeros(steps + 1, npaths)
    for p in 1:npaths
        noise = GeometricBrownianMotionProcess(0.05, 0.2, 0.0, 100.0; rng = ReplayRNG(Zg[:, p]), reseed = false)   # montecarlo.jl:156
        sol = SciMLBase.solve(NoiseProblem(noise, (0.0, 1.0)); dt = 1.0 / steps)                                    # :157
        S[:, p] .= [u[1] for u in sol.u]
    end
    dump("gbm_exact_steps_Z", Zg); dump("gbm_exact_steps_S", S)
    meta("gbm_exact_steps", "ok")
catch err
    meta("gbm_exact_steps", "unavailable: " * replace(sprint(showerror, err), '\n' => ' ')[1:min(end, 200)])
end

# ---- 4. Broadie-Kaya deterministic pieces -------------------------------------------------------------------------------------
bk_sets = Dict(
    "bk_c2" => (S0 = 100.0, V0 = 0.04, κ = 2.0, θ = 0.04, σ = 0.3, ρ = -0.7, r = 0.03, τ = 1 / 12),
    "bk_q8" => (S0 = 100.0, V0 = 1.5, κ = 0.04, θ = 0.3, σ = -0.6, ρ = 0.04, r = 0.05, τ = 364 / 365),   # montecarlo_heston.jl:161-170 (Q8)
    "bk_case1" => (S0 = 100.0, V0 = 0.010201, κ = 6.21, θ = 0.019, σ = 0.61, ρ = -0.7, r = 0.0319, τ = 1.0),
)
for (tag, q) in bk_sets
    dist = Hedgehog.LogHestonDistribution(q.S0, q.V0, q.κ, q.θ, q.σ, q.ρ, q.r, q.τ)   # heston.jl:102-111
    rng = Xoshiro(99)
    VTs = [Hedgehog.sample_V_T(rng, dist) for _ in 1:24]                  # :125-133
    us = rand(Xoshiro(100), 24)
    maxJ = 4096
    phis = fill(NaN, 2, maxJ, 24)        # (re, im) of evaluate_chf(iter, h j, theta_prev), j = 1..J
    out = zeros(10, 24)                  # VT, u, mean, variance, h, J, x, F(x) - u, logIk, F(max_guess) - u
    mom = zeros(6, 24)                   # (re, im) of phi(+h0), phi(0), phi(-h0), theta carried as moments_from_cf does
    for (i, VT) in enumerate(VTs)
        it = Hedgehog.HestonCFIterator(VT, dist)                          # :163-176
        θp = NaN
        ϕp, θp = Hedgehog.evaluate_chf(it, 1e-2, θp)                       # sample_from_cf.jl:52-54
        ϕ0, θp = Hedgehog.evaluate_chf(it, 0.0, θp)
        ϕm, _ = Hedgehog.evaluate_chf(it, -1e-2, θp)
        mom[:, i] .= (real(ϕp), imag(ϕp), real(ϕ0), imag(ϕ0), real(ϕm), imag(ϕm))
        mean_, var_ = Hedgehog.moments_from_cf(it)                        # :50-64
        s2 = max(var_, 1e-12)                                             # :32
        u = us[i]
        guess = mean_ + sqrt(s2) * Hedgehog.quantile(Hedgehog.Normal(), u) # :33
        guess = guess > 0 ? guess : mean_ * 0.01                          # :34
        max_guess = mean_ + 11 * sqrt(s2)                                 # :35
        h = π / (mean_ + 5 * sqrt(s2))                                    # :37
        θj = NaN
        J = 0
        for j in 1:maxJ                                                   # the loop of cdf_from_cf, :84-93
            ϕ, θj = Hedgehog.evaluate_chf(it, h * j, θj)
            phis[1, j, i] = real(ϕ); phis[2, j, i] = imag(ϕ)
            J = j
            abs(ϕ) / j < π * 1e-3 / 2 && break
        end
        cdf = x -> Hedgehog.cdf_from_cf(it, x, h)                         # :38
        x = Hedgehog.inverse_cdf(cdf, u, guess, max_guess)                # :39
        out[:, i] .= (VT, u, mean_, var_, h, J, x, cdf(x) - u, real(it.logIκ), cdf(max_guess) - u)
    end
    dump("$(tag)_phis", phis[:, 1:Int(maximum(out[6, :])), :]); dump("$(tag)_out", out); dump("$(tag)_moments", mom)
    meta("$(tag)_params", "S0=$(q.S0) V0=$(q.V0) kappa=$(q.κ) theta=$(q.θ) xi=$(q.σ) rho=$(q.ρ) r=$(q.r) tau=$(repr(q.τ))")
end

open(joinpath(OUT, "manifest.txt"), "w") do io
    foreach(l -> println(io, l), MANIFEST)
end
println("wrote ", length(MANIFEST), " manifest lines to ", OUT)
