# HedgehogB200.jl — Julia host layer of libhedgehog_mc.so (B200 / sm_100a Monte Carlo pricing path).
#
# Adds methods to Hedgehog.solve for two new method types, `B200MonteCarlo` and `B200LSM`, that carry the same fields
# as Hedgehog's `MonteCarlo` (src/pricing_methods/montecarlo.jl:127-131) and `LSM`
# (src/pricing_methods/least_squares_montecarlo.jl:31-34) and return the same solution types
# (`MonteCarloSolution`, `LSMSolution`, src/solutions/pricing_solutions.jl:22-27, 78-84). Nothing numerical happens in
# Julia: scalars are extracted exactly as the reference extracts them and handed to the C ABI (include/hedgehog_mc.h)
# through `ccall`. There is no CUDA.jl and no CPU fallback: without the library or a B200 every call throws.
#
# Greeks need no lens table here. Every extracted scalar may be a `ForwardDiff.Dual`: the solve then unpacks values and
# partials, runs ONE tangent launch (in-kernel dual numbers, hh_mc_european_tangent_sums) and rebuilds `price::Dual`. So
#   * Hedgehog's own `solve(::GreekProblem, ::ForwardAD, method)` (greeks_problem.jl:249-262) works unchanged and returns
#     its `(greek = deriv,)`; FiniteDifference and SecondOrderGreekProblem + FiniteDifference work unchanged too
#     (deterministic Philox streams: common random numbers);
#   * `OptimizerAlgo()`'s `AutoForwardDiff` (calibration.jl:57, 74-98) differentiates through
#     `solve(::BasketPricingProblem, ::B200MonteCarlo)` — all quotes of an expiry and all parameters in one launch;
#   * `solve(::BatchGreekProblem, ::ForwardAD, ::B200MonteCarlo)` seeds one partial per lens and prices once.
#
# NOTE: this image has no Julia toolchain, so this file has not been executed here. hedgehog.jl_b200/api.py is its Python
# (ctypes) twin that the test-suite runs against the same ABI, and tests/test_julia_mirror.py checks every `struct` and
# every `ccall` signature below against include/hedgehog_mc.h. See INTEGRATION.md.
module HedgehogB200

using Hedgehog
using Hedgehog: PricingProblem, BasketPricingProblem, VanillaOption, European, American, Spot, AbstractPricingMethod,
                AbstractMarketInputs, BlackScholesInputs, HestonInputs, PriceDynamics, LognormalDynamics, HestonDynamics,
                SimulationStrategy, SimulationConfig, EulerMaruyama, BlackScholesExact, HestonBroadieKaya,
                NoVarianceReduction, Antithetic, MonteCarlo, LSM, MonteCarloSolution, LSMSolution, BatchGreekProblem,
                ForwardAD, yearfrac, add_yearfrac, zero_rate, df, get_vol
import ForwardDiff
using ForwardDiff: Dual, Partials, value, partials
using Libdl

export B200MonteCarlo, B200LSM, b200_library!, AsianOption, BarrierOption, DigitalOption, solve_with_bs_control,
       peer_export, peer_connect, QuasiRandom

# ---- library handle -----------------------------------------------------------------------------------------------
const LIB = Ref{String}(get(ENV, "HEDGEHOG_MC_LIB", joinpath(@__DIR__, "..", "libhedgehog_mc.so")))
b200_library!(path::AbstractString) = (LIB[] = path)

const HH_VERSION = 200
const HH_OK = Cint(0)
const HH_ERR_ARG = Cint(-1)
const HH_ERR_UNSUPPORTED = Cint(-2)
const HH_MODEL_GBM, HH_MODEL_HESTON = Cint(0), Cint(1)
const HH_SCHEME_EM, HH_SCHEME_EXACT_TERMINAL, HH_SCHEME_EXACT_STEPS, HH_SCHEME_HESTON_BK = Cint(0), Cint(1), Cint(2), Cint(3)
const HH_RNG_PHILOX, HH_RNG_NORMALS, HH_RNG_PHILOX_64 = Cint(0), Cint(1), Cint(2)
const HH_FLAG_SPLIT_STEP, HH_FLAG_Q1_SQRT_MEAN = UInt32(1), UInt32(2)
const HH_IPC_HANDLE_BYTES = 64

# ---- POD mirrors of include/hedgehog_mc.h (field order and types are the ABI; checked by tests/test_julia_mirror.py) ---
struct HHModel            # hh_model
    kind::Int32
    flags::UInt32
    S0::Float64
    r::Float64
    T::Float64
    sigma::Float64
    V0::Float64
    kappa::Float64
    theta::Float64
    xi::Float64
    rho::Float64
    m11::Float64
    m12::Float64
    m21::Float64
    m22::Float64
end
struct HHBkConfig         # hh_bk_config
    n_std::Int32
    maxiter_newton::Int32
    maxiter_bisection::Int32
    max_terms::Int32
    h_fd::Float64             # > 0: noise-aware (see include/hedgehog_mc.h); < 0: plain finite differences at |h_fd|
    cf_tol::Float64
    atol::Float64
end
HHBkConfig() = HHBkConfig(5, 10, 100, 4096, 1e-2, 1e-3, 1e-4)  # sample_from_cf.jl:27,50,75,110-112
struct HHSim              # hh_sim
    n_paths::Int64
    path_offset::Int64
    n_steps::Int32
    scheme::Int32
    vr::Int32
    precision::Int32
    rng_mode::Int32
    reserved::Int32
    base_seed::UInt64
    seeds::Ptr{UInt64}
    normals::Ptr{Float64}
    seeds_len::UInt64
    normals_len::UInt64
    bk::HHBkConfig
end
struct HHPayoff           # hh_payoff
    strike::Float64
    cp::Float64
end
struct HHResult           # hh_result
    sum::Float64
    sumsq::Float64
    n::Int64
    price::Float64
    std_error::Float64
    n_nonfinite::Int64
    n_fallback::Int64
    kernel_ms::Float64
end
struct HHTangent          # hh_tangent
    dS0::Float64
    dr::Float64
    dsigma::Float64
    dV0::Float64
    dkappa::Float64
    dtheta::Float64
    dxi::Float64
    dm11::Float64
    dm12::Float64
    dm21::Float64
    dm22::Float64
    ddiscount::Float64
end
struct HHLsmResult        # hh_lsm_result
    sum::Float64
    sumsq::Float64
    n::Int64
    price::Float64
    std_error::Float64
    n_dates_skipped::Int64
    kernel_ms::Float64
    path_ms::Float64
    regress_ms::Float64
end
struct HHComm             # hh_comm
    allreduce_sum_f64::Ptr{Cvoid}
    user::Ptr{Cvoid}
    rank::Int32
    world::Int32
end
struct HHPathPayoff       # hh_path_payoff
    kind::Int32
    reserved::Int32
    strike::Float64
    cp::Float64
    barrier::Float64
    amount::Float64
end

# ---- context ---------------------------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
end
const CTX = Dict{Int,Context}()

function context(device::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0")))
    get!(CTX, device) do
        v = ccall((:hh_version, LIB[]), Cint, ())
        v == HH_VERSION || error("libhedgehog_mc.so has ABI version $v, this host file was written for $HH_VERSION")
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:hh_create, LIB[]), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device)
        rc == HH_OK || error("hh_create(device=$device) failed ($rc): " *
                             unsafe_string(ccall((:hh_last_error, LIB[]), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = Context(h[])
        finalizer(c -> ccall((:hh_destroy, LIB[]), Cint, (Ptr{Cvoid},), c.h), ctx)
        ctx
    end
end

function check(ctx::Context, rc::Cint, what)
    rc == HH_OK && return
    msg = unsafe_string(ccall((:hh_last_error, LIB[]), Cstring, (Ptr{Cvoid},), ctx.h))
    rc == HH_ERR_ARG && throw(ArgumentError("$what: $msg"))           # mirrors montecarlo.jl:65-66
    rc == HH_ERR_UNSUPPORTED && throw(MethodError(Hedgehog.solve, (what, msg)))   # what the reference raises (SURVEY Q5)
    error("$what failed ($rc): $msg")
end

# ---- method types: the three fields of Hedgehog.MonteCarlo + execution options -------------------------------------------
struct B200MonteCarlo{P<:PriceDynamics,S<:SimulationStrategy,C<:SimulationConfig} <: AbstractPricingMethod
    dynamics::P
    strategy::S
    config::C
    ensemble::Bool      # materialise MonteCarloSolution.ensemble on the host (8 B per trajectory D2H)
    base_seed::Union{Nothing,UInt64}  # one Philox key + trajectory index in the counter, instead of config.seeds per path
    precision::Symbol   # :f64, or :f32 = the Float32 fast mode (Heston Euler-Maruyama only, HH_PREC_F32)
    rng::Symbol         # :philox (one Philox block per Heston step) or :philox64 (HH_RNG_PHILOX_64, opt-in fast stream)
    # one Julia process per GPU: this process simulates trajectories [N rank / world, N (rank + 1) / world) of the job and
    # `allreduce` sums a Vector{Float64} over the processes (e.g. v -> MPI.Allreduce(v, +, comm)); identity on one GPU
    rank::Int
    world::Int
    allreduce::Function
end
B200MonteCarlo(d, s, c; ensemble = true, base_seed = nothing, precision = :f64, rng = :philox, rank = 0, world = 1,
               allreduce = identity) =
    B200MonteCarlo(d, s, c, ensemble, base_seed === nothing ? nothing : UInt64(base_seed), precision, rng, rank, world, allreduce)
B200MonteCarlo(m::MonteCarlo; kw...) = B200MonteCarlo(m.dynamics, m.strategy, m.config; kw...)

struct B200LSM{M<:B200MonteCarlo} <: AbstractPricingMethod
    mc_method::M
    degree::Int
    stopping_info::Bool   # copy stopping_info back (12 B per column D2H); false: LSMSolution.stopping_info is empty
    spot_paths::Bool      # copy the (steps + 1) x columns spot matrix back (C3: 4.08 GB, 232 ms against 3.8 ms of kernels)
end
B200LSM(mc::B200MonteCarlo, degree::Int; stopping_info = true, spot_paths = true) = B200LSM(mc, degree, stopping_info, spot_paths)
B200LSM(d::PriceDynamics, s::SimulationStrategy, c::SimulationConfig, degree::Int; stopping_info = true, spot_paths = true, kw...) =
    B200LSM(B200MonteCarlo(d, s, c; kw...), degree, stopping_info, spot_paths)
B200LSM(m::LSM; stopping_info = true, spot_paths = true, kw...) =
    B200LSM(B200MonteCarlo(m.mc_method; kw...), m.degree, stopping_info, spot_paths)

# ---- scalar extraction, exactly as the reference does it (values may be ForwardDiff.Dual) -------------------------------
# Cholesky factor of [1 rho; rho 1] (any factor gives the same law, heston.jl:18-20); a Dual rho carries d(factor)/d rho
corr_factor(rho) = (one(rho), zero(rho), rho, sqrt(1 - rho^2))

# (S0, r, sigma, V0, kappa, theta, xi, rho, m11, m12, m21, m22): the differentiable scalars of hh_model, in its field order
function model_scalars(prob_inputs::BlackScholesInputs, ::LognormalDynamics)
    m = prob_inputs
    r = zero_rate(m.rate, 0.0)                                    # montecarlo.jl:150
    sigma = get_vol(m.sigma, nothing, nothing)                    # :151
    (m.spot, r, sigma, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0)
end
function model_scalars(prob_inputs::HestonInputs, ::HestonDynamics)
    m = prob_inputs
    r = zero_rate(m.rate, 0.0)                                    # :200
    (m.spot, r, 0.0, m.V0, m.κ, m.θ, m.σ, m.ρ, corr_factor(m.ρ)...)   # :201
end
model_kind(::LognormalDynamics) = HH_MODEL_GBM
model_kind(::HestonDynamics) = HH_MODEL_HESTON
model_flags(::LognormalDynamics) = HH_FLAG_SPLIT_STEP | HH_FLAG_Q1_SQRT_MEAN
model_flags(::HestonDynamics) = HH_FLAG_SPLIT_STEP

plain(x::Real) = Float64(x)
plain(x::Dual) = (v = value(x); v isa Dual &&
    throw(ArgumentError("nested dual numbers (second-order ForwardAD) are not propagated by the B200 kernels; use " *
                        "SecondOrderGreekProblem with FiniteDifference — the reference's own choice for Monte Carlo gamma " *
                        "(test/agreement/greeks_agreement.jl:219-224)")); Float64(v))
partial_of(x::Real, p) = 0.0
partial_of(x::Dual, p) = Float64(partials(x, p))

function hh_model(kind, flags, T, s::Tuple)
    T isa Dual && throw(ArgumentError("sensitivities to the expiry are not propagated by the B200 kernels; use FiniteDifference"))
    HHModel(kind, flags, plain(s[1]), plain(s[2]), Float64(T), plain.(s[3:12])...)
end
hh_tangent(s::Tuple, discount, p) =
    HHTangent(partial_of(s[1], p), partial_of(s[2], p), partial_of(s[3], p), partial_of(s[4], p), partial_of(s[5], p),
              partial_of(s[6], p), partial_of(s[7], p), partial_of(s[9], p), partial_of(s[10], p), partial_of(s[11], p),
              partial_of(s[12], p), partial_of(discount, p))

scheme_of(::EulerMaruyama, for_lsm) = HH_SCHEME_EM
scheme_of(::BlackScholesExact, for_lsm) = for_lsm ? HH_SCHEME_EXACT_STEPS : HH_SCHEME_EXACT_TERMINAL
scheme_of(::HestonBroadieKaya, for_lsm) = HH_SCHEME_HESTON_BK
vr_of(::NoVarianceReduction) = Cint(0)
vr_of(::Antithetic) = Cint(1)
# randomised van der Corput points in the one-draw exact sampler (HH_VR_QUASI_RANDOM; roadmap "quasi-random", not in Hedgehog)
struct QuasiRandom <: Hedgehog.VarianceReductionStrategy end
vr_of(::QuasiRandom) = Cint(2)

# this process's share [lo, hi) of the job's trajectories (contiguous blocks of the global index, SURVEY 8e)
shard(method::B200MonteCarlo) = (N = method.config.trajectories;
                                 (N * method.rank ÷ method.world, N * (method.rank + 1) ÷ method.world))

# `f(sim)` runs with the seed vector pinned for the duration of the ccall
function with_sim(f, method::B200MonteCarlo, scheme::Cint; dates_from_config::Bool = false)
    cfg = method.config
    exact = scheme == HH_SCHEME_EXACT_TERMINAL || scheme == HH_SCHEME_HESTON_BK
    # exact strategies ignore `steps` (montecarlo.jl:454-459); LSM and path-dependent payoffs under HestonBroadieKaya use
    # them as the number of exactly simulated dates
    steps = (exact && !dates_from_config) ? 1 : cfg.steps
    lo, hi = shard(method)
    prec = method.precision === :f32 ? Cint(1) : Cint(0)          # HH_PREC_F32 / HH_PREC_F64
    rng = method.rng === :philox64 ? HH_RNG_PHILOX_64 : HH_RNG_PHILOX
    # hh_sim.reserved: bit length of the JOB's trajectory count when this process simulates one shard of it (the LSM regression
    # fixes its Chebyshev interval from it, identically on every rank)
    job_bits = method.world > 1 ? Int32(ndigits(cfg.trajectories, base = 2)) : Int32(0)
    if method.base_seed !== nothing || exact
        key = method.base_seed === nothing ? UInt64(cfg.seeds[1]) : method.base_seed   # Xoshiro(seeds[1]) :456 -> ONE stream
        sim = HHSim(hi - lo, lo, steps, scheme, vr_of(cfg.variance_reduction), prec, rng, job_bits, key, C_NULL, C_NULL, 0, 0, HHBkConfig())
        return f(sim)
    end
    seeds = Vector{UInt64}(cfg.seeds[lo+1:hi])                      # remake(prob; seed = seeds[i]) :331
    GC.@preserve seeds begin
        sim = HHSim(hi - lo, lo, steps, scheme, vr_of(cfg.variance_reduction), prec, rng, job_bits, 0,
                    pointer(seeds), C_NULL, length(seeds), 0, HHBkConfig())
        f(sim)
    end
end

# ---- European Monte Carlo on one simulation: `strikes` payoffs of one expiry and one call/put flag ------------------------
# Returns (prices, ensemble). Plain inputs: hh_mc_european. Dual inputs: hh_mc_european_tangent_sums, one direction per
# partial (8 per launch), price::Dual rebuilt with the product rule for the discount factor.
function european_prices(inputs::AbstractMarketInputs, expiry, cp::Float64, strikes::Vector, method::B200MonteCarlo)
    ctx = context()
    any(k -> k isa Dual, strikes) &&
        throw(ArgumentError("sensitivities to the strike are not propagated by the B200 kernels; use FiniteDifference"))
    s = model_scalars(inputs, method.dynamics)
    T = yearfrac(inputs.referenceDate, expiry)                                    # montecarlo.jl:147
    discount = df(inputs.rate, expiry)                                            # :489
    model = hh_model(model_kind(method.dynamics), model_flags(method.dynamics), T, s)
    scheme = scheme_of(method.strategy, false)
    payoffs = [HHPayoff(Float64(k), cp) for k in strikes]
    npay = length(payoffs)
    (1 <= npay <= 256) || throw(ArgumentError("between 1 and 256 payoffs per launch (got $npay)"))
    DT = promote_type(map(typeof, s)..., typeof(discount))
    lo, hi = shard(method)
    nloc = hi - lo
    anti = method.config.variance_reduction isa Antithetic
    if !(DT <: Dual)
        terminal = method.ensemble ? Vector{Float64}(undef, anti ? 2nloc : nloc) : Float64[]
        res = Vector{HHResult}(undef, npay)
        with_sim(method, scheme) do sim
            GC.@preserve payoffs res terminal begin
                rc = ccall((:hh_mc_european, LIB[]), Cint,
                           (Ptr{Cvoid}, Ref{HHModel}, Ref{HHSim}, Ptr{HHPayoff}, Cint, Cdouble, Ptr{HHResult}, Ptr{Float64}, Csize_t),
                           ctx.h, model, sim, pointer(payoffs), npay, plain(discount), pointer(res),
                           method.ensemble ? pointer(terminal) : Ptr{Float64}(C_NULL), length(terminal))
                check(ctx, rc, "hh_mc_european")
            end
        end
        sums = method.allreduce(vcat([[r.sum, Float64(r.n)] for r in res]...))   # [sum_k, n_k] per payoff, over all processes
        prices = [plain(discount) * sums[2k-1] / sums[2k] for k in 1:npay]       # :490
        ensemble = (anti && method.ensemble) ? (terminal[1:nloc], terminal[nloc+1:end]) : terminal     # final_sample :398-402
        return prices, ensemble
    end
    # ---- Dual inputs -----------------------------------------------------------------------------------------------------
    # (MonteCarloSolution.ensemble is not materialised on the tangent path)
    NP = ForwardDiff.npartials(DT)
    sd = map(x -> convert(DT, x), s)
    dd = convert(DT, discount)
    mean_pay = zeros(npay)
    dprice = zeros(npay, NP)
    ntot = Float64(method.config.trajectories)
    for chunk in Iterators.partition(1:NP, 8)                                     # up to 8 directions per launch
        tans = [hh_tangent(sd, dd, p) for p in chunk]
        nt = length(tans)
        sums = zeros(npay * (2 + 2nt))
        ms = Ref{Cdouble}(0.0)
        with_sim(method, scheme) do sim
            GC.@preserve tans payoffs sums begin
                rc = ccall((:hh_mc_european_tangent_sums, LIB[]), Cint,
                           (Ptr{Cvoid}, Ref{HHModel}, Ptr{HHTangent}, Cint, Ref{HHSim}, Ptr{HHPayoff}, Cint, Ptr{Float64},
                            Cdouble, Ptr{Float64}, Ref{Cdouble}),
                           ctx.h, model, pointer(tans), nt, sim, pointer(payoffs), npay, pointer(sums), 0.0,
                           Ptr{Float64}(C_NULL), ms)
                check(ctx, rc, "hh_mc_european_tangent_sums")
            end
        end
        sums = method.allreduce(sums)
        for k in 1:npay
            base = (k - 1) * (2 + 2nt)
            mean_pay[k] = sums[base+1] / ntot
            for (q, p) in enumerate(chunk)   # d(D mean)/dp = D mean(dpayoff) + dD mean(payoff)
                dprice[k, p] = plain(dd) * sums[base+2+q] / ntot + partial_of(dd, p) * mean_pay[k]
            end
        end
    end
    Tag = ForwardDiff.tagtype(DT)
    prices = [Dual{Tag}(plain(dd) * mean_pay[k], Partials(ntuple(p -> dprice[k, p], NP))) for k in 1:npay]
    return prices, Float64[]
end

# ---- solve: European Monte Carlo (montecarlo.jl:478-493) ---------------------------------------------------------------
function Hedgehog.solve(prob::PricingProblem{VanillaOption{TS,TE,European,C,Spot},I},
                        method::B200MonteCarlo) where {TS,TE,C,I<:AbstractMarketInputs}
    prices, ensemble = european_prices(prob.market_inputs, prob.payoff.expiry, prob.payoff.call_put(), [prob.payoff.strike], method)
    return MonteCarloSolution(prob, method, prices[1], ensemble)    # :492
end

# ---- solve: a basket priced on common trajectories (src/calibration/basket.jl:35-38 loops solve per payoff) -----------
# Payoffs sharing expiry and call/put are one launch (up to 256 strikes); the reference's result type is kept.
function Hedgehog.solve(prob::BasketPricingProblem{P,M}, method::B200MonteCarlo) where {P<:VanillaOption,M<:AbstractMarketInputs}
    quiet = B200MonteCarlo(method.dynamics, method.strategy, method.config, false, method.base_seed, method.precision,
                           method.rng, method.rank, method.world, method.allreduce)
    sols = Vector{Any}(undef, length(prob.payoffs))
    groups = Dict{Any,Vector{Int}}()
    for (i, p) in enumerate(prob.payoffs)
        p.exercise_style isa European || throw(MethodError(Hedgehog.solve, (PricingProblem(p, prob.market_inputs), method)))
        push!(get!(groups, (p.expiry, p.call_put()), Int[]), i)
    end
    for ((expiry, cp), idxs) in groups, chunk in Iterators.partition(idxs, 256)
        prices, _ = european_prices(prob.market_inputs, expiry, cp, [prob.payoffs[i].strike for i in chunk], quiet)
        for (i, price) in zip(chunk, prices)
            sols[i] = MonteCarloSolution(PricingProblem(prob.payoffs[i], prob.market_inputs), method, price, Float64[])
        end
    end
    return Hedgehog.BasketPricingSolution(prob, [s for s in sols])
end

# ---- solve: BatchGreekProblem + ForwardAD = ONE simulation (greeks_problem.jl:559-568 loops over the lenses) -------------
struct B200Tag end
function Hedgehog.solve(gprob::BatchGreekProblem, ::ForwardAD, method::B200MonteCarlo)
    prob, lenses = gprob.pricing_problem, collect(gprob.lenses)
    NP = length(lenses)
    seeded = prob
    for (p, lens) in enumerate(lenses)     # x_p + eps_p, like ForwardDiff.derivative seeds its single partial (:257-260)
        x0 = lens(prob)
        seeded = Hedgehog.set(seeded, lens, Dual{B200Tag}(Float64(x0), Partials(ntuple(q -> q == p ? 1.0 : 0.0, NP))))
    end
    price = Hedgehog.solve(seeded, method).price
    Dict(lens => partials(price, p) for (p, lens) in enumerate(lenses))
end

# ---- solve: American LSM (least_squares_montecarlo.jl:99-136) ------------------------------------------------------------
function Hedgehog.solve(prob::PricingProblem{VanillaOption{TS,TE,American,C,S},I},
                        method::B200LSM) where {TS,TE,C,S,I<:AbstractMarketInputs}
    ctx = context()
    mc = method.mc_method
    m = prob.market_inputs
    T = yearfrac(m.referenceDate, prob.payoff.expiry)                       # :104
    model = hh_model(model_kind(mc.dynamics), model_flags(mc.dynamics), T, model_scalars(m, mc.dynamics))
    scheme = scheme_of(mc.strategy, true)
    payoff = HHPayoff(Float64(prob.payoff.strike), prob.payoff.call_put())
    nsteps = mc.config.steps
    step_discount = plain(df(m.rate, add_yearfrac(m.referenceDate, T / nsteps)))   # :110
    lo, hi = shard(mc)
    ncols = (hi - lo) * (mc.config.variance_reduction isa Antithetic ? 2 : 1)
    stop_idx = method.stopping_info ? Vector{Int32}(undef, ncols) : Int32[]
    stop_val = method.stopping_info ? Vector{Float64}(undef, ncols) : Float64[]
    spot = method.spot_paths ? Matrix{Float64}(undef, nsteps + 1, ncols) : Matrix{Float64}(undef, 0, 0)   # column = trajectory, :50
    out = Ref(HHLsmResult(0, 0, 0, 0, 0, 0, 0, 0, 0))
    # world > 1: the per-date regression moments are exchanged inside the kernel over the peers' mailboxes
    # (allreduce_sum_f64 == NULL, hedgehog_mc.h "peer mailboxes"); peer_connect must have been called on every rank
    comm = Ref(HHComm(C_NULL, C_NULL, mc.rank, mc.world))
    with_sim(mc, scheme; dates_from_config = scheme == HH_SCHEME_HESTON_BK) do sim
        GC.@preserve stop_idx stop_val spot comm begin
            rc = ccall((:hh_lsm_american, LIB[]), Cint,
                       (Ptr{Cvoid}, Ref{HHModel}, Ref{HHSim}, Ref{HHPayoff}, Cint, Cdouble, Ptr{HHComm}, Ref{HHLsmResult},
                        Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                       ctx.h, model, sim, payoff, method.degree, step_discount,
                       mc.world > 1 ? Base.unsafe_convert(Ptr{HHComm}, comm) : Ptr{HHComm}(C_NULL), out,
                       method.stopping_info ? pointer(stop_idx) : Ptr{Int32}(C_NULL),
                       method.stopping_info ? pointer(stop_val) : Ptr{Float64}(C_NULL),
                       method.spot_paths ? pointer(spot) : Ptr{Float64}(C_NULL))
            check(ctx, rc, "hh_lsm_american")
        end
    end
    sums = mc.allreduce([out[].sum, Float64(out[].n)])
    stopping_info = [(Int(stop_idx[p]), stop_val[p]) for p in eachindex(stop_idx)]   # :112, :163-164
    return LSMSolution(prob, method, sums[1] / sums[2], stopping_info, spot)          # :132-135
end

# ---- path-dependent payoffs (roadmap Phase 5, derivatives_pricing_roadmap.md:73-80; hh_mc_path_dependent) -------------
# Hedgehog has no Asian / barrier / digital payoff types yet. These follow VanillaOption's conventions (payoffs.jl:101-140:
# strike, expiry in ticks, call_put functor) so that solve(PricingProblem(payoff, inputs), B200MonteCarlo(...)) reads
# like the European solve. `monitor_every`: monitoring dates are every k-th step of config.steps, expiry included.
abstract type PathDependentPayoff <: Hedgehog.AbstractPayoff end
struct AsianOption{TS,TE,C<:Hedgehog.AbstractCallPut} <: PathDependentPayoff
    strike::TS
    expiry::TE
    call_put::C
    geometric::Bool
    monitor_every::Int
end
struct BarrierOption{TS,TE,C<:Hedgehog.AbstractCallPut} <: PathDependentPayoff
    strike::TS
    barrier::TS
    expiry::TE
    call_put::C
    up::Bool
    knock_out::Bool
    rebate::TS
    monitor_every::Int
end
struct DigitalOption{TS,TE,C<:Hedgehog.AbstractCallPut} <: PathDependentPayoff
    strike::TS
    expiry::TE
    call_put::C
    cash::Union{Nothing,TS}   # nothing: asset-or-nothing
    monitor_every::Int
end
hh_path_payoff(p::AsianOption) = HHPathPayoff(p.geometric ? 2 : 1, 0, p.strike, p.call_put(), 0.0, 0.0)
hh_path_payoff(p::BarrierOption) =
    HHPathPayoff(p.up ? (p.knock_out ? 3 : 4) : (p.knock_out ? 5 : 6), 0, p.strike, p.call_put(), p.barrier, p.rebate)
hh_path_payoff(p::DigitalOption) =
    p.cash === nothing ? HHPathPayoff(8, 0, p.strike, p.call_put(), 0.0, 0.0) : HHPathPayoff(7, 0, p.strike, p.call_put(), 0.0, p.cash)

function path_dependent_results(ctx, model::HHModel, sim::HHSim, payoffs::Vector{HHPathPayoff}, discount, every::Integer)
    res = Vector{HHResult}(undef, length(payoffs))
    GC.@preserve payoffs res begin
        rc = ccall((:hh_mc_path_dependent, LIB[]), Cint,
                   (Ptr{Cvoid}, Ref{HHModel}, Ref{HHSim}, Cint, Ptr{HHPathPayoff}, Cint, Cdouble, Ptr{HHResult}, Ptr{Float64}, Csize_t),
                   ctx.h, model, sim, every, pointer(payoffs), length(payoffs), discount, pointer(res), Ptr{Float64}(C_NULL), 0)
        check(ctx, rc, "hh_mc_path_dependent")
    end
    return res
end

function Hedgehog.solve(prob::PricingProblem{P,I}, method::B200MonteCarlo) where {P<:PathDependentPayoff,I<:AbstractMarketInputs}
    ctx = context()
    m = prob.market_inputs
    T = yearfrac(m.referenceDate, prob.payoff.expiry)
    model = hh_model(model_kind(method.dynamics), model_flags(method.dynamics), T, model_scalars(m, method.dynamics))
    scheme = scheme_of(method.strategy, true)                     # BlackScholesExact in its stepping form
    discount = plain(df(m.rate, prob.payoff.expiry))              # montecarlo.jl:489
    r = with_sim(method, scheme; dates_from_config = scheme == HH_SCHEME_HESTON_BK) do sim
        path_dependent_results(ctx, model, sim, [hh_path_payoff(prob.payoff)], discount, prob.payoff.monitor_every)[1]
    end
    sums = method.allreduce([r.sum, Float64(r.n)])
    return MonteCarloSolution(prob, method, discount * sums[1] / sums[2], Float64[])
end

# ---- Black-Scholes control variate for Heston vanilla prices (roadmap "Control variates using Black-Scholes") ---------------
# The kernel advances a log-GBM trajectory on the same Brownian increments next to every Heston trajectory
# (HH_PD_BS_CONTROL = 10, HH_PD_VANILLA_MINUS_BS = 11 in include/hedgehog_mc.h); the expectation of the control is
# Hedgehog's own BlackScholesAnalytic price at sigma_cv. beta === nothing: estimated from one pilot launch on another seed.
sample_var(r::HHResult) = max((r.sumsq - r.n * (r.sum / r.n)^2) / (r.n - 1), 0.0)

function solve_with_bs_control(prob::PricingProblem{VanillaOption{TS,TE,European,C,Spot},I}, method::B200MonteCarlo;
                               beta = nothing, pilot::Int = 50_000) where {TS,TE,C,I<:AbstractMarketInputs}
    method.dynamics isa HestonDynamics && method.strategy isa EulerMaruyama ||
        throw(ArgumentError("the Black-Scholes control variate runs next to HestonDynamics + EulerMaruyama"))
    ctx = context()
    m = prob.market_inputs
    T = yearfrac(m.referenceDate, prob.payoff.expiry)
    model = hh_model(HH_MODEL_HESTON, model_flags(method.dynamics), T, model_scalars(m, method.dynamics))
    K, cp = Float64(prob.payoff.strike), prob.payoff.call_put()
    discount = plain(df(m.rate, prob.payoff.expiry))
    kT = model.kappa * model.T
    w = abs(kT) > 1e-8 ? -expm1(-kT) / kT : 1 - kT / 2
    sigma_cv = sqrt(max(model.theta + (model.V0 - model.theta) * w, 1e-12))     # mean of E[V_t] over [0, T]
    control = Hedgehog.solve(PricingProblem(prob.payoff, Hedgehog.BlackScholesInputs(m.referenceDate, m.rate, m.spot, sigma_cv)),
                             Hedgehog.BlackScholesAnalytic()).price
    with_sim(method, HH_SCHEME_EM) do sim
        b = beta
        if b === nothing    # Cov(X, Y) = (Var X + Var Y - Var(X - Y)) / 2 from the three sums of one pilot launch
            pilot_key = (method.base_seed === nothing ? UInt64(method.config.seeds[1]) : method.base_seed) ⊻ 0x9E3779B97F4A7C15
            ps = HHSim(min(pilot, sim.n_paths), 0, sim.n_steps, sim.scheme, sim.vr, sim.precision, 0, 0, pilot_key,
                       C_NULL, C_NULL, 0, 0, HHBkConfig())
            px, py, pd = path_dependent_results(ctx, model, ps, [HHPathPayoff(0, 0, K, cp, 0.0, 0.0), HHPathPayoff(10, 0, K, cp, 0.0, 0.0),
                                                                 HHPathPayoff(11, 0, K, cp, 0.0, 1.0)], discount, 1)
            vy = sample_var(py)
            b = vy > 0 ? (sample_var(px) + vy - sample_var(pd)) / (2vy) : 0.0
        end
        r = path_dependent_results(ctx, model, sim, [HHPathPayoff(11, 0, K, cp, 0.0, b)], discount, 1)[1]
        sums = method.allreduce([r.sum, Float64(r.n)])
        MonteCarloSolution(prob, method, discount * sums[1] / sums[2] + b * control, Float64[])
    end
end

# ---- multi-GPU: peer mailboxes for the LSM moments (one Julia process per GPU, e.g. under MPI.jl) ---------------------------
# h = peer_export() on every rank; Allgather the 64-byte handles; peer_connect(rank, world, handles) maps the peers'
# mailboxes. B200LSM solves with world > 1 then exchange the regression moments inside the kernel over NVLink.
function peer_export()
    h = zeros(UInt8, HH_IPC_HANDLE_BYTES)
    ctx = context()
    check(ctx, ccall((:hh_peer_export, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx.h, h), "hh_peer_export")
    h
end
function peer_connect(rank::Integer, world::Integer, handles::Vector{UInt8})
    length(handles) == world * HH_IPC_HANDLE_BYTES || throw(ArgumentError("handles: world x $HH_IPC_HANDLE_BYTES bytes"))
    ctx = context()
    check(ctx, ccall((:hh_peer_connect, LIB[]), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, rank, world, handles),
          "hh_peer_connect")
end
peer_disconnect() = (ctx = context(); check(ctx, ccall((:hh_peer_disconnect, LIB[]), Cint, (Ptr{Cvoid},), ctx.h), "hh_peer_disconnect"))
peer_set_timeout(seconds::Real) =
    (ctx = context(); check(ctx, ccall((:hh_peer_set_timeout, LIB[]), Cint, (Ptr{Cvoid}, Cdouble), ctx.h, seconds), "hh_peer_set_timeout"))

end # module
