/*
 * hh_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic of Hedgehog.jl's Monte Carlo pricing path, used only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker for the CUDA library. Nothing under hedgehog.jl_b200/ may link or call it.
 *
 * PARITY STATUS: "parity unpinned" at the per-path level. The reference is pure Julia, cannot run
 * in this container (no julia), and its tests hold no golden vectors or stored paths — only
 * statistical tolerances against analytic / Carr-Madan / CRR values. The oracle
 * is pinned against (i) every deterministic known answer the reference's tests hold (Black-Scholes
 * 7.9655/16.6994/2.3101/9.8237, CRR 0.25225758542934945 / 0.07409148128021317, payoff, df, ACT/365),
 * (ii) the reference's own statistical agreement tests re-run with their parameters and tolerances,
 * and (iii) the Random123 Philox4x32-10 known-answer vectors. Third-party behaviours that are not
 * in the reference tree (StochasticDiffEq EM split step, DiffEqNoiseProcess increments) are switches.
 *
 * The entry points mirror include/hedgehog_mc.h one to one (prefix hho_) so that a test compares
 * struct against struct on the same inputs.
 */
#ifndef HH_ORACLE_H
#define HH_ORACLE_H

#include "../include/hedgehog_mc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Random123 Philox4x32-10 (Salmon et al. 2011). */
void hho_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The library's native-RNG convention: one Philox block -> one Box-Muller pair. */
void hho_normal_pair(uint64_t key, uint64_t idx, uint32_t block, uint32_t stream, double *z1, double *z2);
/* HH_RNG_PHILOX_64: the pair of Heston step `step` from half a Philox block (32-bit radius uniform, 32-bit angle). */
void hho_normal_pair64(uint64_t key, uint64_t idx, uint32_t step, double *z1, double *z2);
/* Fill Z[path][step][comp] with the native stream's normals (so parity mode can replay it). */
void hho_fill_normals(const hh_model *model, const hh_sim *sim, double *Z);

int hho_threads(void);
int hho_threads_used(void); /* measured inside a parallel region */
void hho_set_threads(int n);

/* solve(::PricingProblem, ::MonteCarlo)  montecarlo.jl:478-493 */
int hho_mc_european(const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                    double discount, hh_result *results, double *terminal, size_t terminal_len);
/* variance at expiry per trajectory (Heston only; diagnostic) */
int hho_heston_em_terminal_v(const hh_model *model, const hh_sim *sim, double *v_terminal);

/* ForwardDiff through solve  greeks_problem.jl:249-262 — hand-derived tangent recursions. */
int hho_mc_european_tangent_sums(const hh_model *model, const hh_tangent *tangents, int ntangents,
                                 const hh_sim *sim, const hh_payoff *payoffs, int npayoffs, double *sums);

/* SecondOrderGreekProblem(spot, spot) + FiniteDifference(eps): three solves at S0 - eps, S0, S0 + eps on the same seeds
 * (greeks_problem.jl:395-412), restated literally; second_sums[k][4] as in hh_mc_european_tangent_sums. */
int hho_mc_european_second_sums(const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                                double spot_bump, double *second_sums);

/* solve(::PricingProblem{American}, ::LSM)  least_squares_montecarlo.jl:99-136.
 * beta_out: nullable, [n_steps+1][degree+1] raw-monomial coefficients per date (0 where skipped). */
int hho_lsm_american(const hh_model *model, const hh_sim *sim, const hh_payoff *payoff, int degree,
                     double step_discount, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                     double *spot_paths, double *beta_out);
/* Backward induction on a caller-supplied grid G[date][col] (date-major), same outputs. */
int hho_lsm_backward(const double *grid, int64_t ncols, int n_steps, const hh_payoff *payoff, int degree,
                     double step_discount, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                     double *beta_out);

#ifdef __cplusplus
}
#endif
#endif
