# HedgehogB200.jl — Julia host layer of libhedgehog_mc.so (B200 / sm_100a Monte Carlo pricing path).
#
# Adds methods to Hedgehog.solve for two new method types, `B200MonteCarlo` and `B200LSM`, that carry the same fields
# as Hedgehog's `MonteCarlo` (src/pricing_methods/montecarlo.jl:127-131) and `LSM`
# (src/pricing_methods/least_squares_montecarlo.jl:31-34) and return the same solution types
# (`MonteCarloSolution`, `LSMSolution`, src/solutions/pricing_solutions.jl:22-27, 78-84). Nothing numerical happens in
# Julia: scalars are extracted exactly as the reference extracts them and handed to the C ABI (include/hedgehog_mc.h)
# through `ccall`. There is no CUDA.jl and no CPU fallback: without the library or a B200 every call throws.
#
# NOTE: this image has no Julia toolchain, so this file has not been executed here; hedgehog.jl_b200/api.py is the
# line-for-line Python (ctypes) twin that the test-suite runs against the same ABI. See INTEGRATION.md.
module HedgehogB200

using Hedgehog
using Hedgehog: PricingProblem, VanillaOption, European, American, Spot, AbstractPricingMethod, AbstractMarketInputs,
                BlackScholesInputs, HestonInputs, PriceDynamics, LognormalDynamics, HestonDynamics, SimulationStrategy,
                SimulationConfig, EulerMaruyama, BlackScholesExact, HestonBroadieKaya, NoVarianceReduction, Antithetic,
                MonteCarlo, LSM, MonteCarloSolution, LSMSolution, GreekProblem, BatchGreekProblem, ForwardAD,
                SpotLens, VolLens, ZeroRateSpineLens, yearfrac, add_yearfrac, zero_rate, df, get_vol
using Libdl

export B200MonteCarlo, B200LSM, b200_library!, AsianOption, BarrierOption, DigitalOption, solve_with_bs_control

# ---- library handle -----------------------------------------------------------------------------------------------
const LIB = Ref{String}(get(ENV, "HEDGEHOG_MC_LIB", joinpath(@__DIR__, "..", "libhedgehog_mc.so")))
b200_library!(path::AbstractString) = (LIB[] = path)

const HH_OK = Cint(0)
const HH_ERR_ARG = Cint(-1)
const HH_ERR_UNSUPPORTED = Cint(-2)
const HH_MODEL_GBM, HH_MODEL_HESTON = Cint(0), Cint(1)
const HH_SCHEME_EM, HH_SCHEME_EXACT_TERMINAL, HH_SCHEME_EXACT_STEPS, HH_SCHEME_HESTON_BK = Cint(0), Cint(1), Cint(2), Cint(3)
const HH_FLAG_SPLIT_STEP, HH_FLAG_Q1_SQRT_MEAN = UInt32(1), UInt32(2)

# ---- POD mirrors of include/hedgehog_mc.h (field order and types are the ABI) -----------------------------------------
struct HHModel
    kind::Int32; flags::UInt32
    S0::Float64; r::Float64; T::Float64; sigma::Float64
    V0::Float64; kappa::Float64; theta::Float64; xi::Float64; rho::Float64
    m11::Float64; m12::Float64; m21::Float64; m22::Float64
end
struct HHBkConfig
    n_std::Int32; maxiter_newton::Int32; maxiter_bisection::Int32; max_terms::Int32
    h_fd::Float64; cf_tol::Float64; atol::Float64
end
HHBkConfig() = HHBkConfig(5, 10, 100, 4096, 1e-2, 1e-3, 1e-4)  # sample_from_cf.jl:27,50,75,110-112
struct HHSim
    n_paths::Int64; path_offset::Int64
    n_steps::Int32; scheme::Int32; vr::Int32; precision::Int32; rng_mode::Int32; reserved::Int32
    base_seed::UInt64
    seeds::Ptr{UInt64}; normals::Ptr{Float64}
    bk::HHBkConfig
end
struct HHPayoff
    strike::Float64; cp::Float64
end
struct HHResult
    sum::Float64; sumsq::Float64; n::Int64; price::Float64; std_error::Float64
    n_nonfinite::Int64; n_fallback::Int64; kernel_ms::Float64
end
struct HHTangent
    dS0::Float64; dr::Float64; dsigma::Float64; dV0::Float64; dkappa::Float64; dtheta::Float64; dxi::Float64
    dm11::Float64; dm12::Float64; dm21::Float64; dm22::Float64; ddiscount::Float64
end
struct HHLsmResult
    sum::Float64; sumsq::Float64; n::Int64; price::Float64; std_error::Float64
    n_dates_skipped::Int64; kernel_ms::Float64; path_ms::Float64; regress_ms::Float64
end

# ---- context ---------------------------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
end
const CTX = Dict{Int,Context}()

function context(device::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0")))
    get!(CTX, device) do
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
    rc == HH_ERR_UNSUPPORTED && throw(MethodError(Hedgehog.solve, (what, msg)))
    error("$what failed ($rc): $msg")
end

# ---- method types: same three fields as Hedgehog.MonteCarlo ------------------------------------------------------------
struct B200MonteCarlo{P<:PriceDynamics,S<:SimulationStrategy,C<:SimulationConfig} <: AbstractPricingMethod
    dynamics::P
    strategy::S
    config::C
    ensemble::Bool      # materialise MonteCarloSolution.ensemble on the host (8 B per trajectory D2H)
    base_seed::Union{Nothing,UInt64}  # one Philox key + trajectory index in the counter, instead of config.seeds per path
    precision::Symbol   # :f64, or :f32 = the Float32 fast mode (Heston Euler-Maruyama only, HH_PREC_F32)
end
B200MonteCarlo(d, s, c; ensemble = true, base_seed = nothing, precision = :f64) =
    B200MonteCarlo(d, s, c, ensemble, base_seed, precision)
B200MonteCarlo(m::MonteCarlo; kw...) = B200MonteCarlo(m.dynamics, m.strategy, m.config; kw...)

struct B200LSM{M<:B200MonteCarlo} <: AbstractPricingMethod
    mc_method::M
    degree::Int
end
B200LSM(d::PriceDynamics, s::SimulationStrategy, c::SimulationConfig, degree::Int; kw...) =
    B200LSM(B200MonteCarlo(d, s, c; kw...), degree)
B200LSM(m::LSM; kw...) = B200LSM(B200MonteCarlo(m.mc_method; kw...), m.degree)

# ---- scalar extraction, exactly as the reference does it -----------------------------------------------------------------
function corr_factor(rho)  # Cholesky factor of [1 rho; rho 1] (any factor gives the same law, heston.jl:18-20)
    (1.0, 0.0, rho, sqrt(1 - rho^2))
end

function hh_model(prob::PricingProblem, ::LognormalDynamics)
    m = prob.market_inputs
    T = yearfrac(m.referenceDate, prob.payoff.expiry)            # montecarlo.jl:147
    r = zero_rate(m.rate, 0.0)                                    # :150
    sigma = get_vol(m.sigma, nothing, nothing)                    # :151
    HHModel(HH_MODEL_GBM, HH_FLAG_SPLIT_STEP | HH_FLAG_Q1_SQRT_MEAN, m.spot, r, T, sigma, 0, 0, 0, 0, 0, 1, 0, 0, 1)
end

function hh_model(prob::PricingProblem, ::HestonDynamics)
    m = prob.market_inputs
    T = yearfrac(m.referenceDate, prob.payoff.expiry)            # montecarlo.jl:197
    r = zero_rate(m.rate, 0.0)                                    # :200
    m11, m12, m21, m22 = corr_factor(m.ρ)
    HHModel(HH_MODEL_HESTON, HH_FLAG_SPLIT_STEP, m.spot, r, T, 0.0, m.V0, m.κ, m.θ, m.σ, m.ρ, m11, m12, m21, m22)  # :201
end

scheme_of(::EulerMaruyama, for_lsm) = HH_SCHEME_EM
scheme_of(::BlackScholesExact, for_lsm) = for_lsm ? HH_SCHEME_EXACT_STEPS : HH_SCHEME_EXACT_TERMINAL
scheme_of(::HestonBroadieKaya, for_lsm) = HH_SCHEME_HESTON_BK
vr_of(::NoVarianceReduction) = Cint(0)
vr_of(::Antithetic) = Cint(1)

# `f(sim)` runs with the seed vector pinned for the duration of the ccall
function with_sim(f, method::B200MonteCarlo, scheme::Cint; dates_from_config::Bool=false)
    cfg = method.config
    exact = scheme == HH_SCHEME_EXACT_TERMINAL || scheme == HH_SCHEME_HESTON_BK
    # exact strategies ignore `steps` (montecarlo.jl:454-459); path-dependent payoffs under HestonBroadieKaya use them as
    # the number of exactly simulated dates
    steps = (exact && !dates_from_config) ? 1 : cfg.steps
    seeds = Vector{UInt64}(cfg.seeds)
    prec = method.precision === :f32 ? Cint(1) : Cint(0)          # HH_PREC_F32 / HH_PREC_F64
    if method.base_seed !== nothing || exact
        key = method.base_seed === nothing ? seeds[1] : method.base_seed   # Xoshiro(seeds[1]) :456 -> ONE stream
        sim = HHSim(cfg.trajectories, 0, steps, scheme, vr_of(cfg.variance_reduction), prec, 0, 0, key, C_NULL, C_NULL, HHBkConfig())
        return f(sim)
    end
    GC.@preserve seeds begin
        sim = HHSim(cfg.trajectories, 0, steps, scheme, vr_of(cfg.variance_reduction), prec, 0, 0, 0,
                    pointer(seeds), C_NULL, HHBkConfig())           # remake(prob; seed = seeds[i]) :331
        f(sim)
    end
end

# ---- solve: European Monte Carlo (montecarlo.jl:478-493) ---------------------------------------------------------------
function Hedgehog.solve(prob::PricingProblem{VanillaOption{TS,TE,European,C,Spot},I},
                        method::B200MonteCarlo) where {TS,TE,C,I<:AbstractMarketInputs}
    ctx = context()
    model = hh_model(prob, method.dynamics)
    scheme = scheme_of(method.strategy, false)
    payoff = HHPayoff(prob.payoff.strike, prob.payoff.call_put())
    discount = df(prob.market_inputs.rate, prob.payoff.expiry)    # :489
    N = method.config.trajectories
    anti = method.config.variance_reduction isa Antithetic
    terminal = method.ensemble ? Vector{Float64}(undef, anti ? 2N : N) : Float64[]
    res = Ref(HHResult(0, 0, 0, 0, 0, 0, 0, 0))
    with_sim(method, scheme) do sim
        GC.@preserve terminal begin
            rc = ccall((:hh_mc_european, LIB[]), Cint,
                       (Ptr{Cvoid}, Ref{HHModel}, Ref{HHSim}, Ref{HHPayoff}, Cint, Cdouble, Ref{HHResult}, Ptr{Float64}, Csize_t),
                       ctx.h, model, sim, payoff, 1, discount, res,
                       method.ensemble ? pointer(terminal) : Ptr{Float64}(C_NULL), length(terminal))
            check(ctx, rc, "hh_mc_european")
        end
    end
    ensemble = anti ? (terminal[1:N], terminal[N+1:end]) : terminal   # final_sample :398-402
    return MonteCarloSolution(prob, method, res[].price, ensemble)    # :492
end

# ---- solve: American LSM (least_squares_montecarlo.jl:99-136) ------------------------------------------------------------
function Hedgehog.solve(prob::PricingProblem{VanillaOption{TS,TE,American,C,S},I},
                        method::B200LSM) where {TS,TE,C,S,I<:AbstractMarketInputs}
    ctx = context()
    mc = method.mc_method
    model = hh_model(prob, mc.dynamics)
    scheme = scheme_of(mc.strategy, true)
    payoff = HHPayoff(prob.payoff.strike, prob.payoff.call_put())
    m = prob.market_inputs
    T = yearfrac(m.referenceDate, prob.payoff.expiry)                       # :104
    nsteps = mc.config.steps
    step_discount = df(m.rate, add_yearfrac(m.referenceDate, T / nsteps))   # :110
    ncols = mc.config.trajectories * (mc.config.variance_reduction isa Antithetic ? 2 : 1)
    stop_idx = Vector{Int32}(undef, ncols)
    stop_val = Vector{Float64}(undef, ncols)
    spot = Matrix{Float64}(undef, nsteps + 1, ncols)                        # column = trajectory, :50
    out = Ref(HHLsmResult(0, 0, 0, 0, 0, 0, 0, 0, 0))
    with_sim(mc, scheme; dates_from_config = scheme == HH_SCHEME_HESTON_BK) do sim
        GC.@preserve stop_idx stop_val spot begin
            rc = ccall((:hh_lsm_american, LIB[]), Cint,
                       (Ptr{Cvoid}, Ref{HHModel}, Ref{HHSim}, Ref{HHPayoff}, Cint, Cdouble, Ptr{Cvoid}, Ref{HHLsmResult},
                        Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
                       ctx.h, model, sim, payoff, method.degree, step_discount, C_NULL, out,
                       pointer(stop_idx), pointer(stop_val), pointer(spot))
            check(ctx, rc, "hh_lsm_american")
        end
    end
    stopping_info = [(Int(stop_idx[p]), stop_val[p]) for p in 1:ncols]      # :112, :163-164
    return LSMSolution(prob, method, out[].price, stopping_info, spot)       # :135
end

# ---- path-dependent payoffs (roadmap Phase 5, derivatives_pricing_roadmap.md:73-80; hh_mc_path_dependent) -------------
# Hedgehog has no Asian / barrier / digital payoff types yet. These follow VanillaOption's conventions (payoffs.jl:101-140:
# strike, expiry in ticks, call_put functor) so that solve(PricingProblem(payoff, inputs), B200MonteCarlo(...)) reads
# like the European solve. `monitor_every`: monitoring dates are every k-th step of config.steps, expiry included.
struct HHPathPayoff
    kind::Int32; reserved::Int32; strike::Float64; cp::Float64; barrier::Float64; amount::Float64
end
abstract type PathDependentPayoff <: Hedgehog.AbstractPayoff end
struct AsianOption{TS,TE,C<:Hedgehog.AbstractCallPut} <: PathDependentPayoff
    strike::TS; expiry::TE; call_put::C; geometric::Bool; monitor_every::Int
end
struct BarrierOption{TS,TE,C<:Hedgehog.AbstractCallPut} <: PathDependentPayoff
    strike::TS; barrier::TS; expiry::TE; call_put::C; up::Bool; knock_out::Bool; rebate::TS; monitor_every::Int
end
struct DigitalOption{TS,TE,C<:Hedgehog.AbstractCallPut} <: PathDependentPayoff
    strike::TS; expiry::TE; call_put::C; cash::Union{Nothing,TS}; monitor_every::Int   # cash === nothing: asset-or-nothing
end
hh_path_payoff(p::AsianOption) = HHPathPayoff(p.geometric ? 2 : 1, 0, p.strike, p.call_put(), 0.0, 0.0)   # kind 9: arithmetic - geometric (control variate)
hh_path_payoff(p::BarrierOption) =
    HHPathPayoff(p.up ? (p.knock_out ? 3 : 4) : (p.knock_out ? 5 : 6), 0, p.strike, p.call_put(), p.barrier, p.rebate)
hh_path_payoff(p::DigitalOption) =
    p.cash === nothing ? HHPathPayoff(8, 0, p.strike, p.call_put(), 0.0, 0.0) : HHPathPayoff(7, 0, p.strike, p.call_put(), 0.0, p.cash)

function Hedgehog.solve(prob::PricingProblem{P,I}, method::B200MonteCarlo) where {P<:PathDependentPayoff,I<:AbstractMarketInputs}
    ctx = context()
    model = hh_model(prob, method.dynamics)
    scheme = scheme_of(method.strategy, true)                     # BlackScholesExact in its stepping form
    payoff = hh_path_payoff(prob.payoff)
    discount = df(prob.market_inputs.rate, prob.payoff.expiry)    # montecarlo.jl:489
    res = Ref(HHResult(0, 0, 0, 0, 0, 0, 0, 0))
    with_sim(method, scheme; dates_from_config = scheme == HH_SCHEME_HESTON_BK) do sim
        rc = ccall((:hh_mc_path_dependent, LIB[]), Cint,
                   (Ptr{Cvoid}, Ref{HHModel}, Ref{HHSim}, Cint, Ref{HHPathPayoff}, Cint, Cdouble, Ref{HHResult}, Ptr{Float64}, Csize_t),
                   ctx.h, model, sim, prob.payoff.monitor_every, payoff, 1, discount, res, Ptr{Float64}(C_NULL), 0)
        check(ctx, rc, "hh_mc_path_dependent")
    end
    return MonteCarloSolution(prob, method, res[].price, Float64[])
end

# ---- Black-Scholes control variate for Heston vanilla prices (roadmap "Control variates using Black-Scholes") ---------------
# The kernel advances a log-GBM trajectory on the same Brownian increments next to every Heston trajectory
# (HH_PD_BS_CONTROL = 10, HH_PD_VANILLA_MINUS_BS = 11 in include/hedgehog_mc.h); the expectation of the control is
# Hedgehog's own BlackScholesAnalytic price at sigma_cv. beta === nothing: estimated from one pilot launch on another seed.
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
sample_var(r::HHResult) = max((r.sumsq - r.n * (r.sum / r.n)^2) / (r.n - 1), 0.0)

function solve_with_bs_control(prob::PricingProblem{VanillaOption{TS,TE,European,C,Spot},I}, method::B200MonteCarlo;
                               beta = nothing, pilot::Int = 50_000) where {TS,TE,C,I<:AbstractMarketInputs}
    method.dynamics isa HestonDynamics && method.strategy isa EulerMaruyama ||
        throw(ArgumentError("the Black-Scholes control variate runs next to HestonDynamics + EulerMaruyama"))
    ctx = context()
    model = hh_model(prob, method.dynamics)
    K, cp = prob.payoff.strike, prob.payoff.call_put()
    m = prob.market_inputs
    discount = df(m.rate, prob.payoff.expiry)
    kT = model.kappa * model.T
    w = abs(kT) > 1e-8 ? -expm1(-kT) / kT : 1 - kT / 2
    sigma_cv = sqrt(max(model.theta + (model.V0 - model.theta) * w, 1e-12))     # mean of E[V_t] over [0, T]
    control = Hedgehog.solve(PricingProblem(prob.payoff, Hedgehog.BlackScholesInputs(m.referenceDate, m.rate, m.spot, sigma_cv)),
                             Hedgehog.BlackScholesAnalytic()).price
    with_sim(method, HH_SCHEME_EM) do sim
        b = beta
        if b === nothing    # Cov(X, Y) = (Var X + Var Y - Var(X - Y)) / 2 from the three sums of one pilot launch
            ps = HHSim(min(pilot, sim.n_paths), 0, sim.n_steps, sim.scheme, sim.vr, sim.precision, 0, 0,
                       sim.base_seed ⊻ 0x9E3779B97F4A7C15, C_NULL, C_NULL, HHBkConfig())
            px, py, pd = path_dependent_results(ctx, model, ps, [HHPathPayoff(0, 0, K, cp, 0.0, 0.0), HHPathPayoff(10, 0, K, cp, 0.0, 0.0),
                                                                 HHPathPayoff(11, 0, K, cp, 0.0, 1.0)], discount, 1)
            vy = sample_var(py)
            b = vy > 0 ? (sample_var(px) + vy - sample_var(pd)) / (2vy) : 0.0
        end
        r = path_dependent_results(ctx, model, sim, [HHPathPayoff(11, 0, K, cp, 0.0, b)], discount, 1)[1]
        MonteCarloSolution(prob, method, r.price + b * control, Float64[])
    end
end

# ---- Greeks: every ForwardAD lens is one tangent direction of the SAME simulation (greeks_problem.jl:249-262, 559-568) ---
function tangent_of(prob::PricingProblem, lens)
    m = prob.market_inputs
    z = zeros(12)
    if lens isa SpotLens
        z[1] = 1.0
    elseif lens isa ZeroRateSpineLens      # the rate moves the drift AND the discount factor
        z[2] = 1.0
        z[12] = -yearfrac(m.rate.reference_date, prob.payoff.expiry) * df(m.rate, prob.payoff.expiry)
    elseif lens isa VolLens
        z[3] = 1.0
    else
        name = string(lens)                # Accessors optics: (@optic _.market_inputs.κ) etc.
        if occursin("V0", name); z[4] = 1.0
        elseif occursin("κ", name); z[5] = 1.0
        elseif occursin("θ", name); z[6] = 1.0
        elseif occursin("σ", name); z[7] = 1.0
        elseif occursin("ρ", name)          # d(Cholesky factor)/d rho
            z[10] = 1.0; z[11] = -m.ρ / sqrt(1 - m.ρ^2)
        elseif occursin("spot", name); z[1] = 1.0
        else
            throw(ArgumentError("no tangent rule for lens $lens"))
        end
    end
    HHTangent(z...)
end

function forward_ad(prob::PricingProblem, lenses, method::B200MonteCarlo)
    ctx = context()
    model = hh_model(prob, method.dynamics)
    scheme = scheme_of(method.strategy, false)
    payoff = HHPayoff(prob.payoff.strike, prob.payoff.call_put())
    discount = df(prob.market_inputs.rate, prob.payoff.expiry)
    greeks = Vector{Float64}(undef, length(lenses))
    for chunk in Iterators.partition(eachindex(lenses), 8)            # up to 8 directions per launch
        tans = [tangent_of(prob, lenses[i]) for i in chunk]
        res = Ref(HHResult(0, 0, 0, 0, 0, 0, 0, 0))
        out = Vector{Float64}(undef, length(tans))
        with_sim(method, scheme) do sim
            GC.@preserve tans out begin
                rc = ccall((:hh_mc_european_tangent, LIB[]), Cint,
                           (Ptr{Cvoid}, Ref{HHModel}, Ptr{HHTangent}, Cint, Ref{HHSim}, Ref{HHPayoff}, Cint, Cdouble,
                            Ref{HHResult}, Ptr{Float64}, Ptr{Float64}),
                           ctx.h, model, pointer(tans), length(tans), sim, payoff, 1, discount, res, pointer(out), C_NULL)
                check(ctx, rc, "hh_mc_european_tangent")
            end
        end
        greeks[collect(chunk)] .= out
    end
    greeks
end

Hedgehog.solve(gprob::GreekProblem, ::ForwardAD, method::B200MonteCarlo) =
    Hedgehog.GreekResult(forward_ad(gprob.pricing_problem, [gprob.wrt], method)[1])

function Hedgehog.solve(gprob::BatchGreekProblem, ::ForwardAD, method::B200MonteCarlo)
    g = forward_ad(gprob.pricing_problem, collect(gprob.lenses), method)
    Dict(lens => g[i] for (i, lens) in enumerate(gprob.lenses))      # greeks_problem.jl:559-568
end
# Multi-GPU (one Julia process per GPU, e.g. MPI.jl): after `h = peer_export()` on every rank and an Allgather of the 64-byte
# handles, `peer_connect(rank, world, handles)` maps the peers' mailboxes; B200LSM solves then pass
# hh_comm(C_NULL, C_NULL, rank, world) and the regression moments are exchanged inside the kernel over NVLink.
function peer_export()
    h = zeros(UInt8, 64)
    ctx = context()
    check(ctx, ccall((:hh_peer_export, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx.h, h), "hh_peer_export")
    h
end
function peer_connect(rank::Integer, world::Integer, handles::Vector{UInt8})
    ctx = context()
    check(ctx, ccall((:hh_peer_connect, LIB[]), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, rank, world, handles),
          "hh_peer_connect")
end

# FiniteDifference Greeks need no method here: Hedgehog's generic code re-solves with bumped inputs
# (greeks_problem.jl:279-329), and the deterministic Philox stream gives it common random numbers.

end # module
