/*
 * hedgehog_mc.h — C ABI of libhedgehog_mc.so, the B200 (sm_100a) Monte Carlo pricing path.
 *
 * This is the drop-in boundary for the hot path of aleCombi/Hedgehog.jl (reference paths are
 * relative to the reference checkout):
 *
 *   solve(::PricingProblem{VanillaOption{..European..}}, ::MonteCarlo)   src/pricing_methods/montecarlo.jl:478-493
 *   solve(::PricingProblem{VanillaOption{..American..}}, ::LSM)          src/pricing_methods/least_squares_montecarlo.jl:99-136
 *   solve(::GreekProblem / ::BatchGreekProblem, ::ForwardAD, method)     src/greeks/greeks_problem.jl:249-262,559-568
 *
 * The reference is pure Julia and has no FFI of its own; a Julia host file binds these entry
 * points with `ccall` (see INTEGRATION.md and hedgehog.jl_b200/julia/HedgehogB200.jl).
 *
 * Conventions
 *   - plain C, POD structs only, no exceptions cross the boundary;
 *   - return code 0 = OK, <0 = argument error, >0 = CUDA error class; text via hh_last_error();
 *   - the caller owns every buffer it passes; the library never keeps a caller pointer past
 *     return; device memory, streams and events live inside the opaque hh_ctx;
 *   - nullable outputs (terminal values, stopping info, spot paths) are only materialised /
 *     copied to the host when the pointer is non-null;
 *   - calls on one hh_ctx are serialised by an internal mutex; each call blocks until its
 *     results are on the host (the *_launch/_collect pair splits that for benchmarking);
 *   - there is NO CPU fallback: every entry point fails with HH_ERR_CUDA if no sm_100 GPU
 *     is usable.
 */
#ifndef HEDGEHOG_MC_H
#define HEDGEHOG_MC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HH_VERSION 200 /* 0.2.0: hh_sim carries seeds_len / normals_len; HH_RNG_PHILOX_64; hh_peer_set_timeout */

/* ---- error codes ------------------------------------------------------------------- */
#define HH_OK 0
#define HH_ERR_ARG (-1)         /* bad argument (mirrors the reference's ArgumentError, montecarlo.jl:65-66); also any
                                 * non-finite model parameter: the truncations inside the schemes would swallow a NaN */
#define HH_ERR_UNSUPPORTED (-2) /* combination the reference itself cannot run (e.g. Q5: Antithetic + BK) */
#define HH_ERR_CUDA 1           /* CUDA runtime failure, see hh_last_error */
#define HH_ERR_NOMEM 2          /* device allocation failed */
#define HH_ERR_COMM 3           /* the caller-supplied allreduce callback failed */
#define HH_ERR_PEER_TIMEOUT 4   /* a peer's contribution did not arrive within the in-kernel time limit */

/* ---- enums (int32 in the structs) ---------------------------------------------------- */
/* dynamics: LognormalDynamics / HestonDynamics                montecarlo.jl:15,22 */
#define HH_MODEL_GBM 0
#define HH_MODEL_HESTON 1

/* strategy                                                    montecarlo.jl:93,108,115 */
#define HH_SCHEME_EM 0             /* EulerMaruyama: log-space EM, montecarlo.jl:166-202 + heston.jl:7-52 */
#define HH_SCHEME_EXACT_TERMINAL 1 /* BlackScholesExact, European path: one draw from marginal_law, montecarlo.jl:293-303,454-459 */
#define HH_SCHEME_EXACT_STEPS 2    /* BlackScholesExact as NoiseProblem (LSM path generator), S-space, montecarlo.jl:140-159 */
#define HH_SCHEME_HESTON_BK 3      /* HestonBroadieKaya, heston.jl:125-300 + sample_from_cf.jl */

/* variance reduction                                           montecarlo.jl:36,43 */
#define HH_VR_NONE 0
#define HH_VR_ANTITHETIC 1
/* Quasi-random sampling of the one-draw exact sampler (roadmap "Stratified sampling / quasi-random",
 * docs/src/derivatives_pricing_roadmap.md:164; not in the reference): HH_SCHEME_EXACT_TERMINAL only. Trajectory g (GLOBAL
 * index: path_offset + i, so the points do not depend on how the job is sharded) takes the base-2 van der Corput point of g
 * under one Cranley-Patterson rotation derived from base_seed,
 *     u_g = frac(bitreverse64(g) 2^-64 + shift(base_seed)),   Z_g = Phi^-1(u_g)   (midpoints of the 2^-53 grid),
 * instead of a pseudo-random normal: an unbiased estimator whose error falls like log N / N. hh_result.std_error is still
 * the plain sample standard error: an UPPER bound here (the points are not independent). */
#define HH_VR_QUASI_RANDOM 2

#define HH_PREC_F64 0
#define HH_PREC_F32 1 /* fast mode (Heston Euler-Maruyama, in-kernel RNG): f32 state and normals from 32-bit uniforms, \
                         two steps per Philox block, f64 payoff accumulation; agrees with f64 statistically (3 sigma) */

#define HH_RNG_PHILOX 0  /* in-kernel Philox4x32-10 + Box-Muller, 52-bit uniforms: one Philox block per Heston step */
#define HH_RNG_NORMALS 1 /* parity mode: consume caller-supplied standard normals */
#define HH_RNG_PHILOX_64 2 /* opt-in fast stream (Heston Euler-Maruyama, f64, European pricing): one Philox4x32-10 block \
                              feeds TWO steps, each step taking 64 random bits — a 32-bit radius uniform \
                              u1 = 1 - (k + 1/2) 2^-32 and a 32-bit angle theta = 2 pi w 2^-32 — evaluated in f64 by the \
                              same table-driven Box-Muller. Counter stream word 2: never reuses HH_RNG_PHILOX numbers. \
                              Same law up to the 2^-32 grid (|z| <= 6.7); restated bit for bit in the oracle. \
                              hh_lsm_american accepts it for the exact GBM generator (LognormalDynamics + BlackScholesExact): \
                              one normal per step, so ONE block feeds FOUR steps. */

/* hh_model.flags */
#define HH_FLAG_SPLIT_STEP 1u   /* EM{split=true}: diffusion evaluated at K = u + dt f(u) [StochasticDiffEq default] */
#define HH_FLAG_Q1_SQRT_MEAN 2u /* marginal_law puts sqrt(alpha) in the mean (montecarlo.jl:302); off = alpha */

typedef struct hh_ctx hh_ctx;

/* Model scalars, extracted on the host exactly as the reference does
 * (T: montecarlo.jl:147; r = zero_rate(rate, 0.0): :150; sigma: :151; Heston fields: :201). */
typedef struct hh_model {
  int32_t kind;  /* HH_MODEL_* */
  uint32_t flags; /* HH_FLAG_* */
  double S0;     /* spot */
  double r;      /* flat zero rate used in the drift */
  double T;      /* yearfrac(referenceDate, expiry), ACT/365 */
  double sigma;  /* GBM volatility (FlatVolSurface) */
  double V0, kappa, theta, xi, rho; /* Heston: initial var, mean reversion, long-run var, vol-of-vol, correlation */
  /* Factor M of the Brownian correlation, M M^T = [1 rho; rho 1]:  dW = sqrt(dt) * M * (Z1, Z2)^T.
   * Any factor gives the same law; it only fixes how parity-mode normals map to increments
   * (CorrelatedWienerProcess, heston.jl:18-20). */
  double m11, m12, m21, m22;
} hh_model;

/* Broadie-Kaya tolerances; defaults are the reference's keyword defaults
 * (sample_from_cf.jl:27 n=5, :50 h=1e-2, :75 cf_tol=1e-3, :110-112 atol=1e-4, 10, 100). */
typedef struct hh_bk_config {
  int32_t n_std;            /* n in h = pi/(mean + n*sd) */
  int32_t maxiter_newton;   /* secant evaluations */
  int32_t maxiter_bisection;
  int32_t max_terms;        /* safety cap on the Fourier series length (reference: 10^9) */
  double h_fd;              /* finite-difference step for the CF moments (moments_from_cf h = 1e-2). > 0: where the
                             * rounding noise of that second difference exceeds 2 % of the variance (short horizons,
                             * low vol of vol: sigma * tau < ~0.01) the variance is re-read at a step scaled to the
                             * law; < 0: plain finite differences at |h_fd| always, as the reference computes them */
  double cf_tol;
  double atol;
} hh_bk_config;

/* SimulationConfig (montecarlo.jl:58-79) + execution knobs. */
typedef struct hh_sim {
  int64_t n_paths;     /* trajectories simulated by THIS call (pairs when antithetic) */
  int64_t path_offset; /* global index of the first local trajectory (multi-GPU shards); 0 on one GPU */
  int32_t n_steps;     /* config.steps; dt = T / n_steps (montecarlo.jl:349) */
  int32_t scheme;      /* HH_SCHEME_* */
  int32_t vr;          /* HH_VR_* */
  int32_t precision;   /* HH_PREC_* */
  int32_t rng_mode;    /* HH_RNG_* */
  int32_t reserved;    /* hh_lsm_american only: bit length of the JOB's trajectory count when this call simulates one shard
                        * of it (0: this call is the whole job). Every rank of a job must pass the same value: it fixes the
                        * interval of the regression's Chebyshev variable, and the ranks' moment sums are added. */
  uint64_t base_seed;  /* Philox key when seeds == NULL; counter carries the global path index */
  const uint64_t *seeds;  /* host, nullable: one key per local trajectory (config.seeds), len >= n_paths */
  const double *normals;  /* host, parity mode only: Z[path][step][component] contiguous */
  uint64_t seeds_len;     /* elements behind `seeds`; checked: >= n_paths (montecarlo.jl:65-66 ArgumentError) */
  uint64_t normals_len;   /* elements behind `normals`; checked: >= n_paths * n_steps * components (1 GBM, 2 Heston; \
                             n_steps counts as 1 for HH_SCHEME_EXACT_TERMINAL) */
  hh_bk_config bk;
} hh_sim;

/* VanillaOption payoff max(cp*(S-K),0)                          payoffs.jl:154-156 */
typedef struct hh_payoff {
  double strike;
  double cp; /* +1 call, -1 put */
} hh_payoff;

typedef struct hh_result {
  double sum;    /* sum over local trajectories of the (pair-averaged) payoff */
  double sumsq;  /* sum of squares of the same */
  int64_t n;     /* trajectories contributing */
  double price;  /* discount * sum / n   (montecarlo.jl:490) */
  double std_error; /* discount * sample std / sqrt(n); not in the reference */
  int64_t n_nonfinite; /* terminal values that were NaN/Inf */
  int64_t n_fallback;  /* BK inversions that fell back (the reference @warns, sample_from_cf.jl:125,131) */
  double kernel_ms;    /* device time of the kernels of this call (CUDA events on the ctx stream) */
} hh_result;

/* One tangent direction = d(model)/d(parameter) plus d(discount)/d(parameter). */
typedef struct hh_tangent {
  double dS0, dr, dsigma, dV0, dkappa, dtheta, dxi;
  double dm11, dm12, dm21, dm22; /* derivative of the correlation factor (rho sensitivity) */
  double ddiscount;
} hh_tangent;

typedef struct hh_lsm_result {
  double sum, sumsq; /* of discount^tau * value over local columns */
  int64_t n;         /* local columns (2*n_paths when antithetic) */
  double price;      /* mean(discount^tau * value)   least_squares_montecarlo.jl:132-133 */
  double std_error;
  int64_t n_dates_skipped; /* dates with no in-the-money path (:122) */
  double kernel_ms;
  double path_ms; /* path generation + store */
  double regress_ms; /* backward induction */
} hh_lsm_result;

/* Sum-allreduce of `count` doubles living in DEVICE memory, in place, ordered after work already
 * enqueued on `cuda_stream`. Supplied by a multi-GPU host (torch.distributed / NCCL); NULL on one GPU. */
typedef int (*hh_allreduce_fn)(void *user, void *dev_ptr, size_t count, void *cuda_stream);
typedef struct hh_comm {
  hh_allreduce_fn allreduce_sum_f64;
  void *user;
  int32_t rank, world;
} hh_comm;

/* ---- peer mailboxes: the multi-GPU exchange of the LSM moments WITHOUT a collective library -------------
 * One process per GPU. Each context owns a small device "mailbox"; hh_peer_export returns its CUDA IPC handle, the
 * host exchanges the handles (any transport: MPI, torch.distributed, files) and hh_peer_connect maps every peer's
 * mailbox into this process. hh_lsm_american then exchanges the per-date regression moments inside the tail of the
 * pass kernel: the last block of each rank stores its 3*degree+3 sums into every peer's mailbox over NVLink
 * (st.global to the mapped peer pointer, release flag), waits for the peers' flags and adds the contributions in
 * rank order — so every rank fits the same polynomial, bit for bit, with no NCCL call and no host round trip.
 * Select it by passing an hh_comm with allreduce_sum_f64 == NULL and world > 1. */
#define HH_IPC_HANDLE_BYTES 64
#define HH_MAX_PEERS 16
int hh_peer_export(hh_ctx *ctx, unsigned char handle[HH_IPC_HANDLE_BYTES]);
int hh_peer_connect(hh_ctx *ctx, int rank, int world, const unsigned char *handles /* world x HH_IPC_HANDLE_BYTES */);
int hh_peer_disconnect(hh_ctx *ctx);
/* In-kernel limit on one wait for a peer's moments (default 30 s, or the environment variable HH_PEER_TIMEOUT_S at
 * hh_create). A rank that gives up marks every rank's mailbox, so ALL ranks return HH_ERR_PEER_TIMEOUT from the same
 * hh_lsm_american call; the connection is dead afterwards (hh_peer_connect again). Hosts should still open a solve with
 * a barrier: ranks must not be further apart than this limit when they enter hh_lsm_american. */
int hh_peer_set_timeout(hh_ctx *ctx, double seconds);

/* ---- lifetime --------------------------------------------------------------------------- */
int hh_version(void);
int hh_create(hh_ctx **out, int device);
int hh_destroy(hh_ctx *ctx);
const char *hh_last_error(hh_ctx *ctx); /* ctx may be NULL: last error of hh_create on this thread */
/* Run on a caller-owned CUDA stream (e.g. torch's current stream); NULL restores the ctx stream. */
int hh_set_stream(hh_ctx *ctx, void *cuda_stream);
int hh_device_info(hh_ctx *ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, size_t *total_mem);
void hh_default_bk_config(hh_bk_config *out);
/* Measured FP64 peak of this GPU: a register-resident DFMA-chain microbenchmark (2 FLOP per DFMA).
 * MEASURED_PEAKS.json holds no FP64 figure, so bench.py's roofline denominator comes from here. */
int hh_bench_fp64_peak(hh_ctx *ctx, double *tflops, double *ms);
/* Ablation of the headline kernel (heston_fast2_kernel, the instantiation hh_mc_european launches for a large
 * base-seed job): the SAME kernel with one part of the step removed, timed on n_paths x n_steps.
 *   part 0: the full step (both streams: rng_mode HH_RNG_PHILOX or HH_RNG_PHILOX_64)
 *   part 1: no Philox (the random words are replaced by a two-instruction counter hash): table-driven Box-Muller + SDE step
 *   part 2: Philox only (the words are XOR-folded into the state; no FP64 work)
 * Results are not prices; `ms` is the device time of the one launch. bench.py reports the split. */
int hh_bench_heston_ablation(hh_ctx *ctx, int64_t n_paths, int n_steps, int rng_mode, int part, double *ms);
/* Debug aid (the stand-in for compute-sanitizer's memcheck / initcheck where that tool is unavailable): with the
 * environment variable HH_DEBUG_GUARDS=1 set before the first hh_create, every device scratch buffer of a context carries
 * a 4 KB guard band on each side and is pre-filled with 0xFF bytes (NaN as f64). *violations = number of guard bytes that
 * were overwritten so far (an out-of-bounds write), or -1 when the variable is not set. Synchronises the device. */
int hh_debug_check_guards(hh_ctx *ctx, int64_t *violations);

/* ---- European Monte Carlo: solve(::PricingProblem, ::MonteCarlo), montecarlo.jl:478-493 ------
 * One simulation prices `npayoffs` vanilla payoffs on the same paths (npayoffs = 1 is the
 * reference's solve; >1 is the strike grid of BasketPricingProblem, src/calibration/basket.jl:35-38).
 * `terminal`: nullable; S_T per trajectory (MonteCarloSolution.ensemble, montecarlo.jl:492),
 * length n_paths (NoVR) or 2*n_paths (antithetic: [plus | minus], montecarlo.jl:400-402). */
int hh_mc_european(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs,
                   int npayoffs, double discount, hh_result *results, double *terminal,
                   size_t terminal_len);

/* Split form: launch enqueues the kernels (inputs uploaded first), collect waits and reads back. */
int hh_mc_european_launch(hh_ctx *ctx, const hh_model *model, const hh_sim *sim,
                          const hh_payoff *payoffs, int npayoffs, int want_terminal);
int hh_mc_european_collect(hh_ctx *ctx, double discount, hh_result *results, double *terminal,
                           size_t terminal_len);

/* ---- pathwise forward-mode Greeks: ForwardDiff through solve, greeks_problem.jl:249-262 ------
 * Propagates `ntangents` dual directions through the same simulation.
 * tangent_results[k*ntangents + p] = d price_k / d direction_p (product rule with ddiscount applied);
 * tangent_stderr likewise (nullable). */
int hh_mc_european_tangent(hh_ctx *ctx, const hh_model *model, const hh_tangent *tangents,
                           int ntangents, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                           double discount, hh_result *results, double *tangent_results,
                           double *tangent_stderr);
/* Raw sums for multi-GPU reduction: out[k*(2+2*ntangents) + {0: sum payoff, 1: sum payoff^2,
 * 2+p: sum dpayoff_p, 2+ntangents+p: sum dpayoff_p^2}] over local trajectories.
 *
 * Second order in the spot from the SAME simulation (SecondOrderGreekProblem(spot, spot) + FiniteDifference(eps),
 * greeks_problem.jl:395-412: ABSOLUTE bump, and the same seeds at S0 - eps, S0, S0 + eps): every scheme here is linear in
 * S0, so the re-solved terminal spots are S_T (S0 +- eps) / S0 and the bumped payoffs are evaluated on the trajectories
 * already simulated. second_sums: nullable, [npayoffs][4] = {sum sd, sum sd^2, sum dd, sum dd^2} with, per trajectory,
 *   sd = payoff(S0 + eps) - 2 payoff(S0) + payoff(S0 - eps)      gamma = discount * mean(sd) / eps^2   (the reference's form)
 *   dd = pathwise delta at S0 + eps minus at S0 - eps            gamma = discount * mean(dd) / (2 eps) (lower variance)
 * spot_bump = eps in (0, S0); ignored when second_sums is NULL. */
int hh_mc_european_tangent_sums(hh_ctx *ctx, const hh_model *model, const hh_tangent *tangents,
                                int ntangents, const hh_sim *sim, const hh_payoff *payoffs,
                                int npayoffs, double *sums, double spot_bump, double *second_sums,
                                double *kernel_ms);

/* ---- Path-dependent payoffs on the simulation grid ------------------------------
 * Not in the reference yet: its roadmap lists them as Phase 5 (derivatives_pricing_roadmap.md:73-80: arithmetic /
 * geometric Asian, cash- and asset-or-nothing digitals, discretely monitored up/down knock-in/out barriers, "Monitoring /
 * Averaging modifiers"). The trajectories are those of hh_mc_european (same schemes, RNG streams, antithetic pairing
 * and reduce_payoffs averaging, montecarlo.jl:430-432); the kernel keeps running statistics per trajectory in registers.
 * Monitoring dates: steps monitor_every, 2 monitor_every, ..., n_steps (t = 0 excluded, expiry included); n_steps must
 * be a multiple of monitor_every. */
#define HH_PD_VANILLA 0        /* max(cp (S_T - K), 0) */
#define HH_PD_ASIAN_ARITH 1    /* max(cp (A - K), 0), A = mean of S over the monitoring dates */
#define HH_PD_ASIAN_GEOM 2     /* max(cp (G - K), 0), G = exp(mean of log S) */
#define HH_PD_UP_OUT 3         /* vanilla unless max S >= barrier on a monitoring date, then `amount` (rebate at expiry) */
#define HH_PD_UP_IN 4          /* vanilla if max S >= barrier on a monitoring date, else `amount` */
#define HH_PD_DOWN_OUT 5       /* vanilla unless min S <= barrier, then `amount` */
#define HH_PD_DOWN_IN 6        /* vanilla if min S <= barrier, else `amount` */
#define HH_PD_DIGITAL_CASH 7   /* `amount` if cp (S_T - K) > 0 */
#define HH_PD_DIGITAL_ASSET 8  /* S_T if cp (S_T - K) > 0 */
#define HH_PD_ASIAN_ARITH_MINUS_GEOM 9 /* max(cp (A - K), 0) - max(cp (G - K), 0): the arithmetic Asian with the geometric one \
                                          as a control variate (roadmap "control variates"): the host adds the \
                                          closed-form geometric price under Black-Scholes */
/* Black-Scholes control variate for HestonDynamics + Euler-Maruyama (roadmap "Control variates using Black-Scholes",
 * SURVEY N3): next to every Heston trajectory the kernel advances a log-GBM trajectory on the SAME Brownian increments
 * dW1, with the constant variance sigma_cv^2 = theta + (V0 - theta)(1 - e^(-kappa T))/(kappa T) (the mean of E[V_t]
 * over [0, T]); its terminal spot S_cv has the closed-form Black-Scholes expectation, which the host adds back. */
#define HH_PD_BS_CONTROL 10        /* max(cp (S_cv - K), 0): the control alone (pilot runs estimate its covariance) */
#define HH_PD_VANILLA_MINUS_BS 11  /* max(cp (S_T - K), 0) - amount * max(cp (S_cv - K), 0), amount = beta */
#define HH_PD_NKINDS 12
#define HH_PD_NSTATS 5         /* per column: S_T, A, G, max S, min S over the monitoring dates */
typedef struct hh_path_payoff {
  int32_t kind; /* HH_PD_* */
  int32_t reserved;
  double strike;
  double cp;      /* +1 call, -1 put */
  double barrier; /* barriers only */
  double amount;  /* barrier rebate (paid at expiry) or digital cash amount */
} hh_path_payoff;

/* model/sim as in hh_mc_european; schemes HH_SCHEME_EM (both models), HH_SCHEME_EXACT_STEPS (GBM, advanced in log
 * space: the same law and the same values to rounding) and HH_SCHEME_HESTON_BK (exact Broadie-Kaya transitions between
 * n_steps dates: no time-stepping bias at the monitoring dates; in-kernel RNG, no antithetic, like hh_mc_european).
 * npayoffs in [1, 256].
 * path_stats: nullable host buffer, HH_PD_NSTATS x ncols row-major (ncols = n_paths, doubled [plus | minus] when
 * antithetic), path_stats_len its length in doubles. */
int hh_mc_path_dependent(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, int monitor_every,
                         const hh_path_payoff *payoffs, int npayoffs, double discount, hh_result *results,
                         double *path_stats, size_t path_stats_len);

/* ---- American LSM: solve(::PricingProblem{American}, ::LSM), least_squares_montecarlo.jl:99-136 --
 * stop_idx/stop_val: nullable, stopping_info[(tau, value)] per column (:112,:163-164);
 * spot_paths: nullable, (n_steps+1) x ncols, column-major like the reference Matrix (:50);
 * comm: nullable; when set, the per-date regression moments are sum-allreduced across ranks.
 * sim->scheme: HH_SCHEME_EXACT_STEPS (LognormalDynamics + BlackScholesExact, the configuration the reference tests),
 * HH_SCHEME_EM (log-GBM / log-Heston Euler-Maruyama) where every saved date holds S = exp(x): the reference takes
 * the saved component raw there (:53) and regresses on log-prices; or HH_SCHEME_HESTON_BK: n_steps exercise dates
 * simulated exactly (Bermudan exercise without time-stepping bias). HH_SCHEME_EXACT_TERMINAL saves no dates:
 * HH_ERR_UNSUPPORTED. */
int hh_lsm_american(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoff,
                    int degree, double step_discount, const hh_comm *comm, hh_lsm_result *out,
                    int32_t *stop_idx, double *stop_val, double *spot_paths);

/* ---- Broadie-Kaya deterministic pieces (parity probes; heston.jl:184-212, sample_from_cf.jl:50-96) --
 * Evaluate Phi(a_j) of the integrated variance for n independent (V0, VT) pairs on the GPU.
 * a: [n][na]; out_re/out_im: [n][na]; the angle is unwrapped along j as the reference does. */
int hh_bk_chf(hh_ctx *ctx, const hh_model *model, double tau, const double *V0, const double *VT,
              int n, const double *a, int na, double *out_re, double *out_im);
/* log I_nu(z) for complex z, real order nu > -1 (SpecialFunctions.besseli, heston.jl:173,207). */
int hh_bk_log_besseli(hh_ctx *ctx, double nu, const double *z_re, const double *z_im, int n,
                      double *out_re, double *out_im);
/* The table-driven elementary functions of the Broadie-Kaya kernels (csrc/hh_bessel.cuh), for accuracy tests:
 * kind 0: out_a = exp(x); 1: out_a = sin(x), out_b = cos(x); 2: out_a = log(x); 3: out_a = atan2(y, x). y and out_b
 * may be NULL for the kinds that do not use them. */
int hh_bk_elementary(hh_ctx *ctx, int kind, const double *x, const double *y, int n, double *out_a, double *out_b);
/* sample_from_cf (sample_from_cf.jl:27-41) for n independent (V0, VT, u) triples, u = the uniform the reference draws
 * at :29. out8[i] = {x = sampled integral of V, mean, variance (moments_from_cf :50-64), h (:37), J = number of series
 * terms (:84-93), status (0 root inside [0, max_guess], 1 secant accepted without a bracket, 2 fell back to max_guess),
 * F(x) - u (0 when the last Newton step was below 1e-9 of the bracket: F is not re-evaluated at the returned x),
 * number of CDF evaluations}. cfg NULL = reference defaults. */
int hh_bk_integral(hh_ctx *ctx, const hh_model *model, double tau, const hh_bk_config *cfg, const double *V0,
                   const double *VT, const double *u, int n, double *out8);
/* sample_V_T (heston.jl:125-133): VT[i] = c * NoncentralChisq(d, lambda(V0[i])), one draw per i from Philox key `seed`. */
int hh_bk_variance(hh_ctx *ctx, const hh_model *model, double tau, const double *V0, int n, uint64_t seed,
                   double *VT);
/* Statistics of the last Broadie-Kaya run on this context: out5 = {inversions that fell back to max_guess,
 * sum of series lengths J, sum of CDF evaluations, transitions, inversions accepted without a bracket}. */
int hh_bk_last_stats(hh_ctx *ctx, double *out5);

#ifdef __cplusplus
}
#endif
#endif /* HEDGEHOG_MC_H */
