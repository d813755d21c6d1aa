/*
 * hh_oracle.c — CPU ORACLE (test infrastructure, NOT product code). See hh_oracle.h.
 *
 * Restates, in plain C, the arithmetic of the reference's Monte Carlo path. Citations are
 * relative to the reference checkout (aleCombi/Hedgehog.jl v0.1.1):
 *   montecarlo.jl  = src/pricing_methods/montecarlo.jl
 *   heston.jl      = src/distributions/heston.jl
 *   lsm.jl         = src/pricing_methods/least_squares_montecarlo.jl
 *   greeks.jl      = src/greeks/greeks_problem.jl
 * Third-party behaviour that is not in the reference tree is marked [upstream] and follows
 * SURVEY.md Appendix A (StochasticDiffEq EM{split=true}, DiffEqNoiseProcess GBM increment).
 */
#include "hh_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 0;

int hho_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}
/* threads an OpenMP parallel region really gets (what bench.py prints as `cores`) */
int hho_threads_used(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
#endif
  return n;
}
void hho_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

/* ------------------------------------------------------------------------------------------
 * RNG: Philox4x32-10 + Box-Muller. The reference draws from Xoroshiro/Xoshiro streams inside
 * third-party packages (montecarlo.jl:331,456); counter-based Philox is this build's native
 * stream (north star), restated here so that GPU and oracle can be compared path by path.
 * ---------------------------------------------------------------------------------------- */
void hho_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* sin(pi t), cos(pi t) for t in [0,2) with exact range reduction (mirrors sincospi on the GPU). */
static void sincospi_ref(double t, double *s, double *c) {
  /* q = nearest multiple of 1/2, f = t - q/2 in [-1/4, 1/4] : all exact in binary64 */
  double q = nearbyint(2.0 * t);
  double f = t - 0.5 * q;
  double sf = sin(M_PI * f), cf = cos(M_PI * f);
  int iq = ((int)q) & 3;
  switch (iq) {
    case 0: *s = sf;  *c = cf;  break;
    case 1: *s = cf;  *c = -sf; break;
    case 2: *s = -sf; *c = -cf; break;
    default: *s = -cf; *c = sf; break;
  }
}

void hho_normal_pair(uint64_t key, uint64_t idx, uint32_t block, uint32_t stream, double *z1, double *z2) {
  uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), block, stream};
  uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
  uint32_t w[4];
  hho_philox4x32_10(ctr, k, w);
  /* bits -> uniforms, the library's convention (hedgehog.jl_b200/csrc/hh_fastnormal.cuh):
   *   u1 = 1 - n1 2^-52, n1 = (w1 & 0xFFFFF) << 32 | (w0 | 1)   in [2^-52, 1 - 2^-52]
   *   u2 = n2 2^-52,     n2 = (w3 & 0xFFFFF) << 32 | w2         in [0, 1)                */
  uint64_t n1 = ((uint64_t)(w[1] & 0xFFFFFu) << 32) | (w[0] | 1u);
  uint64_t n2 = ((uint64_t)(w[3] & 0xFFFFFu) << 32) | w[2];
  double u1 = 1.0 - (double)n1 * 0x1.0p-52; /* exact */
  double u2 = (double)n2 * 0x1.0p-52;       /* exact */
  double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi_ref(2.0 * u2, &s, &c);
  *z1 = r * c;
  *z2 = r * s;
}

/* HH_RNG_PHILOX_64 (include/hedgehog_mc.h; hedgehog.jl_b200/csrc/hh_fastnormal.cuh): the Box-Muller pair of Heston
 * step `step` from 64 random bits — words (0, 1) of Philox block step/2 (counter stream word 2) when step is even, words
 * (2, 3) when it is odd:
 *   angle   n2 = (wa & 0xFFFFF) << 32 | wa                          theta = 2 pi n2 2^-52
 *   radius  n1 = (wb & 0xFFFFF) << 32 | (wb & 0xFFF00000) | 0x80000  u1 = 1 - n1 2^-52 = 1 - (rotl(wb, 12) + 1/2) 2^-32 */
void hho_normal_pair64(uint64_t key, uint64_t idx, uint32_t step, double *z1, double *z2) {
  uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), step >> 1, 2u};
  uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
  uint32_t w[4];
  hho_philox4x32_10(ctr, k, w);
  const uint32_t wa = w[(step & 1u) * 2], wb = w[(step & 1u) * 2 + 1];
  uint64_t n1 = ((uint64_t)(wb & 0xFFFFFu) << 32) | (uint64_t)((wb & 0xFFF00000u) | 0x80000u);
  uint64_t n2 = ((uint64_t)(wa & 0xFFFFFu) << 32) | wa;
  double u1 = 1.0 - (double)n1 * 0x1.0p-52; /* exact */
  double u2 = (double)n2 * 0x1.0p-52;       /* exact */
  double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi_ref(2.0 * u2, &s, &c);
  *z1 = r * c;
  *z2 = r * s;
}

/* key / counter-index of local trajectory i (hedgehog_mc.h: hh_sim.seeds / base_seed) */
static inline void path_stream(const hh_sim *sim, int64_t i, uint64_t *key, uint64_t *idx) {
  if (sim->seeds) { *key = sim->seeds[i]; *idx = 0; }
  else { *key = sim->base_seed; *idx = (uint64_t)(sim->path_offset + i); }
}

static int ncomp_of(const hh_model *m, const hh_sim *s) {
  (void)s;
  return m->kind == HH_MODEL_HESTON ? 2 : 1;
}
static int nsteps_of(const hh_sim *s) { return s->scheme == HH_SCHEME_EXACT_TERMINAL ? 1 : s->n_steps; }

/* Normal draws of trajectory i at step n. GBM: step n uses component (n&1) of block n>>1;
 * Heston: both components of block n. Parity mode reads Z[path][step][comp]. */
static inline void draw(const hh_model *m, const hh_sim *sim, int64_t i, int n, uint64_t key, uint64_t idx,
                        double *z1, double *z2) {
  int nc = ncomp_of(m, sim);
  if (sim->rng_mode == HH_RNG_NORMALS) {
    const double *z = sim->normals + ((size_t)i * (size_t)nsteps_of(sim) + (size_t)n) * (size_t)nc;
    *z1 = z[0];
    *z2 = nc == 2 ? z[1] : 0.0;
    return;
  }
  if (nc == 2 && sim->rng_mode == HH_RNG_PHILOX_64) {
    hho_normal_pair64(key, idx, (uint32_t)n, z1, z2);
  } else if (nc == 2) {
    hho_normal_pair(key, idx, (uint32_t)n, 0u, z1, z2);
  } else if (sim->rng_mode == HH_RNG_PHILOX_64) {
    /* the exact GBM generator of LSM under the 64-bit stream: one Philox block per FOUR steps — step n takes component
     * n & 1 of the pair built from 64 bits, hho_normal_pair64(key, idx, n >> 1) (hedgehog.jl_b200/csrc/hh_lsm.cu) */
    double a, b;
    hho_normal_pair64(key, idx, (uint32_t)(n >> 1), &a, &b);
    *z1 = (n & 1) ? b : a;
    *z2 = 0.0;
  } else {
    double a, b;
    hho_normal_pair(key, idx, (uint32_t)(n >> 1), 0u, &a, &b);
    *z1 = (n & 1) ? b : a;
    *z2 = 0.0;
  }
}

void hho_fill_normals(const hh_model *model, const hh_sim *sim, double *Z) {
  int nc = ncomp_of(model, sim), ns = nsteps_of(sim);
  hh_sim s = *sim;
  if (s.rng_mode == HH_RNG_NORMALS) s.rng_mode = HH_RNG_PHILOX;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < sim->n_paths; ++i) {
    uint64_t key, idx;
    path_stream(&s, i, &key, &idx);
    for (int n = 0; n < ns; ++n) {
      double z1, z2;
      draw(model, &s, i, n, key, idx, &z1, &z2);
      double *z = Z + ((size_t)i * ns + n) * nc;
      z[0] = z1;
      if (nc == 2) z[1] = z2;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * One trajectory -> terminal spot (plus, minus). `grid`, when non-null, receives every saved
 * state of the S-space generator (LSM), stride `gstride` between dates.
 * ---------------------------------------------------------------------------------------- */
typedef struct { double Sp, Sm, vp, vm, cvp, cvm; } terminal_t; /* cv*: terminal spot of the Black-Scholes control (Heston EM f64) */

/* sigma_cv^2 of the Black-Scholes control variate (include/hedgehog_mc.h HH_PD_BS_CONTROL): mean of E[V_t] over [0, T] */
static double bs_control_variance(const hh_model *m) {
  double kT = m->kappa * m->T;
  double w = fabs(kT) > 1e-8 ? -expm1(-kT) / kT : 1.0 - 0.5 * kT;
  double v = m->theta + (m->V0 - m->theta) * w;
  return v > 1e-12 ? v : 1e-12;
}

static terminal_t simulate_one(const hh_model *m, const hh_sim *sim, int64_t i, double *grid_p, double *grid_m,
                               size_t gstride) {
  terminal_t out = {0, 0, 0, 0, 0, 0};
  uint64_t key, idx;
  path_stream(sim, i, &key, &idx);
  const int anti = sim->vr == HH_VR_ANTITHETIC;
  double z1, z2;

  if (sim->scheme == HH_SCHEME_EXACT_TERMINAL) {
    /* marginal_law + final_sample: montecarlo.jl:293-303, 384-390. Q1: sqrt(alpha) in the mean. */
    double alpha = m->T;
    double c = (m->flags & HH_FLAG_Q1_SQRT_MEAN) ? sqrt(alpha) : alpha;
    double mu = log(m->S0) + (m->r - m->sigma * m->sigma / 2) * c;
    double sd = m->sigma * sqrt(alpha);
    draw(m, sim, i, 0, key, idx, &z1, &z2);
    double X = mu + sd * z1;
    out.Sp = exp(X);
    out.Sm = anti ? exp(2 * mu - X) : 0.0;
    return out;
  }

  const int M = sim->n_steps;
  const double dt = m->T / M; /* montecarlo.jl:349,367 */
  const double sqdt = sqrt(dt);

  if (m->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_EM) {
    /* LogGBMProblem heston.jl:33-52 ; EM: K = x + dt f ; x' = K + g dW [upstream] */
    double drift = m->r - 0.5 * (m->sigma * m->sigma);
    double xp = log(m->S0), xm = xp;
    /* grid (LSM): the corrected S-space extraction exp(x) of the log state, not the raw component the
     * reference takes (least_squares_montecarlo.jl:53, Q7) */
    if (grid_p) grid_p[0] = m->S0;
    if (grid_m) grid_m[0] = m->S0;
    for (int n = 0; n < M; ++n) {
      draw(m, sim, i, n, key, idx, &z1, &z2);
      double dW = sqdt * z1;
      xp = (xp + dt * drift) + m->sigma * dW;
      if (grid_p) grid_p[(size_t)(n + 1) * gstride] = exp(xp);
      if (anti) {
        xm = (xm + dt * drift) + m->sigma * (-dW); /* NoiseGrid(t, -W) montecarlo.jl:258 */
        if (grid_m) grid_m[(size_t)(n + 1) * gstride] = exp(xm);
      }
    }
    out.Sp = exp(xp); /* final_sample montecarlo.jl:398 */
    out.Sm = anti ? exp(xm) : 0.0;
    return out;
  }

  if (m->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_EXACT_STEPS) {
    /* GeometricBrownianMotionProcess [upstream]: S += S (exp((r - s^2/2) dt + s sqrt(dt) Z) - 1);
     * antithetic = same seeds, sigma -> -sigma (montecarlo.jl:270-284). State is S itself. */
    double sg = m->sigma;
    double drift = (m->r - 0.5 * (sg * sg)) * dt;
    double Sp = m->S0, Sm = m->S0;
    if (grid_p) grid_p[0] = Sp;
    if (grid_m) grid_m[0] = Sm;
    for (int n = 0; n < M; ++n) {
      draw(m, sim, i, n, key, idx, &z1, &z2);
      double e = sg * sqdt * z1;
      Sp = Sp + Sp * (exp(drift + e) - 1.0);
      if (grid_p) grid_p[(size_t)(n + 1) * gstride] = Sp;
      if (anti) {
        Sm = Sm + Sm * (exp(drift - e) - 1.0);
        if (grid_m) grid_m[(size_t)(n + 1) * gstride] = Sm;
      }
    }
    out.Sp = Sp;
    out.Sm = anti ? Sm : 0.0;
    return out;
  }

  if (m->kind == HH_MODEL_HESTON && sim->scheme == HH_SCHEME_EM && sim->precision == HH_PREC_F32) {
    /* Float32 fast mode (new-build convention, include/hedgehog_mc.h HH_PREC_F32): the same scheme as below in
     * binary32, 32-bit uniforms, two steps per Philox block (stream word 1), state y = log(S/S0) - r t.
     * The GPU uses MUFU approximations of lg2 / sin / cos / sqrt, so agreement is to ~1e-4 per path, not bitwise. */
    const int split = (m->flags & HH_FLAG_SPLIT_STEP) != 0;
    const float nhdt = (float)(-0.5 * dt), nkdt = (float)(-(m->kappa * dt)), ktdt = (float)(m->kappa * m->theta * dt);
    const float a11 = (float)(sqdt * m->m11), a12 = (float)(sqdt * m->m12);
    const float b21 = (float)(m->xi * (sqdt * m->m21)), b22 = (float)(m->xi * (sqdt * m->m22));
    float yp = 0.0f, vp = (float)m->V0, ym = 0.0f, vm = vp;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int n = 0; n < M; ++n) {
      if ((n & 1) == 0) {
        uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(n >> 1), 1u};
        uint32_t k2[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
        hho_philox4x32_10(ctr, k2, w);
      }
      const uint32_t wa = w[(n & 1) * 2], wb = w[(n & 1) * 2 + 1];
      union { uint32_t u; float f; } fa, fb;
      fa.u = (wa >> 9) | 0x3F800000u;
      fb.u = (wb >> 9) | 0x3F800000u;
      const float R2 = -1.3862943611198906f * log2f(2.0f - fa.f);
      const float th = fmaf(fb.f, 6.283185307179586f, -9.42477796076938f);
      const float c = cosf(th), sn = sinf(th);
      const float cc1 = fmaf(a12, sn, a11 * c), cc2 = fmaf(b22, sn, b21 * c);
      {
        float vplus = fmaxf(vp, 0.0f);
        float K1 = fmaf(nhdt, vplus, yp), K2 = fmaf(nkdt, vplus, vp + ktdt);
        float sr = sqrtf((split ? fmaxf(K2, 0.0f) : vplus) * R2);
        yp = fmaf(sr, cc1, K1);
        vp = fmaf(sr, cc2, K2);
      }
      if (anti) {
        float vplus = fmaxf(vm, 0.0f);
        float K1 = fmaf(nhdt, vplus, ym), K2 = fmaf(nkdt, vplus, vm + ktdt);
        float sr = sqrtf((split ? fmaxf(K2, 0.0f) : vplus) * R2);
        ym = fmaf(-sr, cc1, K1);
        vm = fmaf(-sr, cc2, K2);
      }
    }
    const double shift = log(m->S0) + (double)M * (m->r * dt);
    out.Sp = exp((double)yp + shift);
    out.vp = vp;
    out.Sm = anti ? exp((double)ym + shift) : 0.0;
    out.vm = vm;
    return out;
  }

  if (m->kind == HH_MODEL_HESTON && sim->scheme == HH_SCHEME_EM) {
    /* LogHestonProblem heston.jl:7-31 : full truncation, diagonal noise, correlated Wiener;
     * EM{split=true} [upstream]: K = u + dt f(u); u' = K + g(K) .* dW. */
    const int split = (m->flags & HH_FLAG_SPLIT_STEP) != 0;
    const double a11 = sqdt * m->m11, a12 = sqdt * m->m12, a21 = sqdt * m->m21, a22 = sqdt * m->m22;
    double xp = log(m->S0), vp = m->V0, xm = xp, vm = vp;
    const double cv_sigma = sqrt(bs_control_variance(m)), cv_drift = (m->r - 0.5 * cv_sigma * cv_sigma) * dt;
    double xbp = xp, xbm = xp; /* log-GBM control on the same dW1 */
    if (grid_p) grid_p[0] = m->S0; /* LSM grid: S = exp(x), see the GBM EM branch */
    if (grid_m) grid_m[0] = m->S0;
    for (int n = 0; n < M; ++n) {
      if (n > 0 && grid_p) grid_p[(size_t)n * gstride] = exp(xp);
      if (n > 0 && grid_m) grid_m[(size_t)n * gstride] = exp(xm);
      draw(m, sim, i, n, key, idx, &z1, &z2);
      double dW1 = a11 * z1 + a12 * z2;
      double dW2 = a21 * z1 + a22 * z2;
      {
        double vplus = fmax(vp, 0.0);
        double K1 = xp + dt * (m->r - 0.5 * vplus);
        double K2 = vp + dt * (m->kappa * (m->theta - vplus));
        double s = sqrt(fmax(split ? K2 : vp, 0.0));
        xp = K1 + s * dW1;
        vp = K2 + (m->xi * s) * dW2;
      }
      if (anti) {
        double vplus = fmax(vm, 0.0);
        double K1 = xm + dt * (m->r - 0.5 * vplus);
        double K2 = vm + dt * (m->kappa * (m->theta - vplus));
        double s = sqrt(fmax(split ? K2 : vm, 0.0));
        xm = K1 + s * (-dW1);
        vm = K2 + (m->xi * s) * (-dW2);
      }
      xbp = (xbp + cv_drift) + cv_sigma * dW1;
      xbm = (xbm + cv_drift) + cv_sigma * (-dW1);
    }
    out.cvp = exp(xbp);
    out.cvm = exp(xbm);
    out.Sp = exp(xp);
    out.vp = vp;
    out.Sm = anti ? exp(xm) : 0.0;
    out.vm = vm;
    if (grid_p) grid_p[(size_t)M * gstride] = out.Sp;
    if (grid_m) grid_m[(size_t)M * gstride] = out.Sm;
    return out;
  }
  out.Sp = NAN;
  return out;
}

static int check_args(const hh_model *m, const hh_sim *sim) {
  if (!m || !sim || sim->n_paths <= 0) return HH_ERR_ARG;
  if (sim->scheme != HH_SCHEME_EXACT_TERMINAL && sim->n_steps <= 0) return HH_ERR_ARG;
  if (sim->rng_mode == HH_RNG_NORMALS && !sim->normals) return HH_ERR_ARG;
  if (sim->rng_mode == HH_RNG_PHILOX_64 && sim->precision == HH_PREC_F64 &&
      !(m->kind == HH_MODEL_HESTON && sim->scheme == HH_SCHEME_EM) &&
      !(m->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_EXACT_STEPS)) /* the latter: the LSM generator's stream */
    return HH_ERR_UNSUPPORTED;
  if (sim->rng_mode == HH_RNG_PHILOX_64 && sim->precision != HH_PREC_F64) return HH_ERR_UNSUPPORTED;
  if (sim->scheme == HH_SCHEME_HESTON_BK) return HH_ERR_UNSUPPORTED; /* BK oracle lives in oracle/bk_ref.py (scipy AMOS) */
  if (m->kind == HH_MODEL_HESTON && sim->scheme != HH_SCHEME_EM) return HH_ERR_ARG;
  if (m->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_HESTON_BK) return HH_ERR_ARG;
  return HH_OK;
}

static inline double payoff_of(const hh_payoff *p, double S) { /* payoffs.jl:154-156 */
  return fmax(p->cp * (S - p->strike), 0.0);
}

int hho_mc_european(const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                    double discount, hh_result *results, double *terminal, size_t terminal_len) {
  int rc = check_args(model, sim);
  if (rc) return rc;
  if (npayoffs < 0 || (npayoffs > 0 && (!payoffs || !results))) return HH_ERR_ARG;
  const int64_t N = sim->n_paths;
  const int anti = sim->vr == HH_VR_ANTITHETIC;
  if (terminal && terminal_len < (size_t)(anti ? 2 * N : N)) return HH_ERR_ARG;

  long double *sum = calloc((size_t)npayoffs + 1, sizeof(long double));
  long double *sumsq = calloc((size_t)npayoffs + 1, sizeof(long double));
  int64_t nonfinite = 0;

#pragma omp parallel
  {
    long double *ls = calloc((size_t)npayoffs + 1, sizeof(long double));
    long double *lq = calloc((size_t)npayoffs + 1, sizeof(long double));
    int64_t lnf = 0;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
      terminal_t t = simulate_one(model, sim, i, NULL, NULL, 0);
      if (!isfinite(t.Sp) || (anti && !isfinite(t.Sm))) lnf++;
      if (terminal) {
        terminal[i] = t.Sp;
        if (anti) terminal[N + i] = t.Sm; /* (plus, minus) tuple montecarlo.jl:400-402 */
      }
      for (int k = 0; k < npayoffs; ++k) {
        double p = payoff_of(&payoffs[k], t.Sp);
        if (anti) p = (p + payoff_of(&payoffs[k], t.Sm)) / 2; /* reduce_payoffs montecarlo.jl:430-432 */
        ls[k] += p;
        lq[k] += (long double)p * p;
      }
    }
#pragma omp critical
    {
      for (int k = 0; k < npayoffs; ++k) { sum[k] += ls[k]; sumsq[k] += lq[k]; }
      nonfinite += lnf;
    }
    free(ls);
    free(lq);
  }
  for (int k = 0; k < npayoffs; ++k) {
    hh_result *r = &results[k];
    memset(r, 0, sizeof(*r));
    r->sum = (double)sum[k];
    r->sumsq = (double)sumsq[k];
    r->n = N;
    long double mean = sum[k] / N;
    r->price = (double)(discount * mean); /* montecarlo.jl:489-490 */
    long double var = N > 1 ? (sumsq[k] - N * mean * mean) / (N - 1) : 0;
    if (var < 0) var = 0;
    r->std_error = (double)(discount * sqrtl(var / N));
    r->n_nonfinite = nonfinite;
  }
  free(sum);
  free(sumsq);
  return HH_OK;
}

/* ------------------------------------------------------------------------------------------
 * Path-dependent payoffs on the simulation grid (include/hedgehog_mc.h hh_mc_path_dependent; the reference's
 * roadmap Phase 5, derivatives_pricing_roadmap.md:73-80 — not implemented there yet). The statistics are taken in
 * S-SPACE from the saved spots of each trajectory (the grid the LSM extension fills), independently of the
 * kernel's log-space bookkeeping: A = mean S, G = exp(mean log S), max S, min S over the monitoring dates.
 * ---------------------------------------------------------------------------------------- */
static double pd_payoff_of(const hh_path_payoff *c, const double *st, double Scv) {
  const double ST = st[0], A = st[1], G = st[2], mx = st[3], mn = st[4];
  const double vanilla = fmax(c->cp * (ST - c->strike), 0.0);
  switch (c->kind) {
    case HH_PD_ASIAN_ARITH: return fmax(c->cp * (A - c->strike), 0.0);
    case HH_PD_ASIAN_GEOM: return fmax(c->cp * (G - c->strike), 0.0);
    case HH_PD_ASIAN_ARITH_MINUS_GEOM: return fmax(c->cp * (A - c->strike), 0.0) - fmax(c->cp * (G - c->strike), 0.0);
    case HH_PD_UP_OUT: return mx >= c->barrier ? c->amount : vanilla;
    case HH_PD_UP_IN: return mx >= c->barrier ? vanilla : c->amount;
    case HH_PD_DOWN_OUT: return mn <= c->barrier ? c->amount : vanilla;
    case HH_PD_DOWN_IN: return mn <= c->barrier ? vanilla : c->amount;
    case HH_PD_DIGITAL_CASH: return c->cp * (ST - c->strike) > 0.0 ? c->amount : 0.0;
    case HH_PD_DIGITAL_ASSET: return c->cp * (ST - c->strike) > 0.0 ? ST : 0.0;
    case HH_PD_BS_CONTROL: return fmax(c->cp * (Scv - c->strike), 0.0);
    case HH_PD_VANILLA_MINUS_BS: return vanilla - c->amount * fmax(c->cp * (Scv - c->strike), 0.0);
    default: return vanilla;
  }
}

static void pd_stats_of(const double *spots, int M, int every, double *st) {
  double sum = 0.0, sumlog = 0.0, mx = -INFINITY, mn = INFINITY;
  int m = 0;
  for (int t = every; t <= M; t += every, ++m) {
    sum += spots[t];
    sumlog += log(spots[t]);
    mx = fmax(mx, spots[t]);
    mn = fmin(mn, spots[t]);
  }
  st[0] = spots[M];
  st[1] = sum / m;
  st[2] = exp(sumlog / m);
  st[3] = mx;
  st[4] = mn;
}

int hho_mc_path_dependent(const hh_model *model, const hh_sim *sim, int monitor_every, const hh_path_payoff *payoffs,
                          int npayoffs, double discount, hh_result *results, double *path_stats) {
  int rc = check_args(model, sim);
  if (rc) return rc;
  if (npayoffs < 1 || !payoffs || !results) return HH_ERR_ARG;
  if (!(sim->scheme == HH_SCHEME_EM || (model->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_EXACT_STEPS)))
    return HH_ERR_UNSUPPORTED;
  if (sim->precision != HH_PREC_F64) return HH_ERR_UNSUPPORTED;
  const int M = sim->n_steps;
  if (monitor_every < 1 || M % monitor_every != 0) return HH_ERR_ARG;
  for (int k = 0; k < npayoffs; ++k) {
    if (payoffs[k].kind < 0 || payoffs[k].kind >= HH_PD_NKINDS) return HH_ERR_ARG;
    if ((payoffs[k].kind == HH_PD_BS_CONTROL || payoffs[k].kind == HH_PD_VANILLA_MINUS_BS) &&
        !(model->kind == HH_MODEL_HESTON && sim->scheme == HH_SCHEME_EM))
      return HH_ERR_ARG;
  }
  const int64_t N = sim->n_paths;
  const int anti = sim->vr == HH_VR_ANTITHETIC;
  const int64_t ncols = anti ? 2 * N : N;
  long double *sum = calloc((size_t)npayoffs, sizeof(long double));
  long double *sumsq = calloc((size_t)npayoffs, sizeof(long double));
  int64_t nonfinite = 0;
#pragma omp parallel
  {
    long double *ls = calloc((size_t)npayoffs, sizeof(long double));
    long double *lq = calloc((size_t)npayoffs, sizeof(long double));
    double *gp = malloc(sizeof(double) * (size_t)(M + 1) * 2);
    double *gm = gp + (M + 1);
    int64_t lnf = 0;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
      terminal_t tt = simulate_one(model, sim, i, gp, anti ? gm : NULL, 1);
      double sp[HH_PD_NSTATS], sm[HH_PD_NSTATS];
      pd_stats_of(gp, M, monitor_every, sp);
      if (anti) pd_stats_of(gm, M, monitor_every, sm);
      if (!isfinite(sp[0]) || (anti && !isfinite(sm[0]))) lnf++;
      if (path_stats)
        for (int q = 0; q < HH_PD_NSTATS; ++q) {
          path_stats[(size_t)q * ncols + i] = sp[q];
          if (anti) path_stats[(size_t)q * ncols + N + i] = sm[q];
        }
      for (int k = 0; k < npayoffs; ++k) {
        double p = pd_payoff_of(&payoffs[k], sp, tt.cvp);
        if (anti) p = (p + pd_payoff_of(&payoffs[k], sm, tt.cvm)) / 2; /* reduce_payoffs montecarlo.jl:430-432 */
        ls[k] += p;
        lq[k] += (long double)p * p;
      }
    }
#pragma omp critical
    {
      for (int k = 0; k < npayoffs; ++k) { sum[k] += ls[k]; sumsq[k] += lq[k]; }
      nonfinite += lnf;
    }
    free(ls);
    free(lq);
    free(gp);
  }
  for (int k = 0; k < npayoffs; ++k) {
    hh_result *r = &results[k];
    memset(r, 0, sizeof(*r));
    r->sum = (double)sum[k];
    r->sumsq = (double)sumsq[k];
    r->n = N;
    long double mean = sum[k] / N;
    r->price = (double)(discount * mean);
    long double var = N > 1 ? (sumsq[k] - N * mean * mean) / (N - 1) : 0;
    if (var < 0) var = 0;
    r->std_error = (double)(discount * sqrtl(var / N));
    r->n_nonfinite = nonfinite;
  }
  free(sum);
  free(sumsq);
  return HH_OK;
}

int hho_heston_em_terminal_v(const hh_model *model, const hh_sim *sim, double *v_terminal) {
  int rc = check_args(model, sim);
  if (rc) return rc;
  if (model->kind != HH_MODEL_HESTON || sim->scheme != HH_SCHEME_EM) return HH_ERR_ARG;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < sim->n_paths; ++i) {
    terminal_t t = simulate_one(model, sim, i, NULL, NULL, 0);
    v_terminal[i] = t.vp;
    if (sim->vr == HH_VR_ANTITHETIC) v_terminal[sim->n_paths + i] = t.vm;
  }
  return HH_OK;
}

/* ------------------------------------------------------------------------------------------
 * Tangents. The reference pushes one ForwardDiff.Dual through the entire solve
 * (greeks.jl:249-262); these are the same derivative rules written out by hand:
 *   max(x,0) passes the tangent iff x > 0 (else constant zero, which also keeps sqrt at 0 finite),
 *   payoff tangent = cp * 1{cp (S-K) > 0} * dS,  S = exp(x) => dS = S dx.
 * ---------------------------------------------------------------------------------------- */
typedef struct { double S, dS; } dual_t;

static void simulate_one_tangent(const hh_model *m, const hh_tangent *tg, const hh_sim *sim, int64_t i, double sign,
                                 dual_t *out) {
  uint64_t key, idx;
  path_stream(sim, i, &key, &idx);
  double z1, z2;
  if (sim->scheme == HH_SCHEME_EXACT_TERMINAL) {
    double alpha = m->T;
    double c = (m->flags & HH_FLAG_Q1_SQRT_MEAN) ? sqrt(alpha) : alpha;
    double sa = sqrt(alpha);
    double mu = log(m->S0) + (m->r - m->sigma * m->sigma / 2) * c;
    double dmu = tg->dS0 / m->S0 + (tg->dr - m->sigma * tg->dsigma) * c;
    double sd = m->sigma * sa, dsd = tg->dsigma * sa;
    draw(m, sim, i, 0, key, idx, &z1, &z2);
    double X = mu + sd * z1, dX = dmu + dsd * z1;
    if (sign < 0) { /* exp(2 mean(law) - X)  montecarlo.jl:387 */
      X = 2 * mu - X;
      dX = 2 * dmu - dX;
    }
    out->S = exp(X);
    out->dS = out->S * dX;
    return;
  }
  const int M = sim->n_steps;
  const double dt = m->T / M, sqdt = sqrt(dt);
  if (m->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_EM) {
    double x = log(m->S0), dx = tg->dS0 / m->S0;
    double drift = m->r - 0.5 * (m->sigma * m->sigma), ddrift = tg->dr - m->sigma * tg->dsigma;
    for (int n = 0; n < M; ++n) {
      draw(m, sim, i, n, key, idx, &z1, &z2);
      double dW = sign * sqdt * z1;
      x = (x + dt * drift) + m->sigma * dW;
      dx = (dx + dt * ddrift) + tg->dsigma * dW;
    }
    out->S = exp(x);
    out->dS = out->S * dx;
    return;
  }
  if (m->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_EXACT_STEPS) {
    double sg = sign * m->sigma, dsg = sign * tg->dsigma; /* antithetic: sigma -> -sigma */
    double S = m->S0, dS = tg->dS0;
    double drift = (m->r - 0.5 * (sg * sg)) * dt, ddrift = (tg->dr - sg * dsg) * dt;
    for (int n = 0; n < M; ++n) {
      draw(m, sim, i, n, key, idx, &z1, &z2);
      double y = drift + sg * sqdt * z1, dy = ddrift + dsg * sqdt * z1;
      double g = exp(y) - 1.0, dg = exp(y) * dy;
      double Sn = S + S * g;
      dS = dS + dS * g + S * dg;
      S = Sn;
    }
    out->S = S;
    out->dS = dS;
    return;
  }
  /* Heston EM */
  const int split = (m->flags & HH_FLAG_SPLIT_STEP) != 0;
  const double a11 = sqdt * m->m11, a12 = sqdt * m->m12, a21 = sqdt * m->m21, a22 = sqdt * m->m22;
  const double da11 = sqdt * tg->dm11, da12 = sqdt * tg->dm12, da21 = sqdt * tg->dm21, da22 = sqdt * tg->dm22;
  double x = log(m->S0), dx = tg->dS0 / m->S0, v = m->V0, dv = tg->dV0;
  for (int n = 0; n < M; ++n) {
    draw(m, sim, i, n, key, idx, &z1, &z2);
    double dW1 = sign * (a11 * z1 + a12 * z2), ddW1 = sign * (da11 * z1 + da12 * z2);
    double dW2 = sign * (a21 * z1 + a22 * z2), ddW2 = sign * (da21 * z1 + da22 * z2);
    double vplus = fmax(v, 0.0), dvplus = v > 0 ? dv : 0.0;
    double K1 = x + dt * (m->r - 0.5 * vplus);
    double dK1 = dx + dt * (tg->dr - 0.5 * dvplus);
    double K2 = v + dt * (m->kappa * (m->theta - vplus));
    double dK2 = dv + dt * (tg->dkappa * (m->theta - vplus) + m->kappa * (tg->dtheta - dvplus));
    double w = split ? K2 : v, dw = split ? dK2 : dv;
    double s = sqrt(fmax(w, 0.0));
    double ds = w > 0 ? dw / (2 * s) : 0.0;
    x = K1 + s * dW1;
    dx = dK1 + ds * dW1 + s * ddW1;
    v = K2 + (m->xi * s) * dW2;
    dv = dK2 + (tg->dxi * s + m->xi * ds) * dW2 + (m->xi * s) * ddW2;
  }
  out->S = exp(x);
  out->dS = out->S * dx;
}

int hho_mc_european_tangent_sums(const hh_model *model, const hh_tangent *tangents, int ntangents,
                                 const hh_sim *sim, const hh_payoff *payoffs, int npayoffs, double *sums) {
  int rc = check_args(model, sim);
  if (rc) return rc;
  if (ntangents <= 0 || npayoffs <= 0 || !tangents || !payoffs || !sums) return HH_ERR_ARG;
  const int64_t N = sim->n_paths;
  const int anti = sim->vr == HH_VR_ANTITHETIC;
  const int stride = 2 + 2 * ntangents;
  const size_t tot = (size_t)npayoffs * stride;
  long double *acc = calloc(tot, sizeof(long double));
#pragma omp parallel
  {
    long double *la = calloc(tot, sizeof(long double));
    dual_t *dp = malloc(sizeof(dual_t) * (size_t)ntangents), *dm = malloc(sizeof(dual_t) * (size_t)ntangents);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
      for (int p = 0; p < ntangents; ++p) {
        simulate_one_tangent(model, &tangents[p], sim, i, +1.0, &dp[p]);
        if (anti) simulate_one_tangent(model, &tangents[p], sim, i, -1.0, &dm[p]);
      }
      for (int k = 0; k < npayoffs; ++k) {
        const hh_payoff *po = &payoffs[k];
        double pay = payoff_of(po, dp[0].S);
        double indp = po->cp * (dp[0].S - po->strike) > 0 ? po->cp : 0.0;
        double indm = 0.0;
        if (anti) {
          pay = (pay + payoff_of(po, dm[0].S)) / 2;
          indm = po->cp * (dm[0].S - po->strike) > 0 ? po->cp : 0.0;
        }
        long double *a = la + (size_t)k * stride;
        a[0] += pay;
        a[1] += (long double)pay * pay;
        for (int p = 0; p < ntangents; ++p) {
          double d = indp * dp[p].dS;
          if (anti) d = (d + indm * dm[p].dS) / 2;
          a[2 + p] += d;
          a[2 + ntangents + p] += (long double)d * d;
        }
      }
    }
#pragma omp critical
    for (size_t j = 0; j < tot; ++j) acc[j] += la[j];
    free(la);
    free(dp);
    free(dm);
  }
  for (size_t j = 0; j < tot; ++j) sums[j] = (double)acc[j];
  free(acc);
  return HH_OK;
}

/* Second order in the spot: SecondOrderGreekProblem(spot, spot) with FiniteDifference(eps), greeks.jl:395-412 — THREE
 * solves at S0 - eps, S0, S0 + eps (absolute bump) on the same seeds, restated literally: every trajectory is simulated
 * three times. second_sums[k][4] = {sum sd, sum sd^2, sum dd, sum dd^2} over trajectories (pair-averaged when antithetic),
 *   sd = payoff(S0 + eps) - 2 payoff(S0) + payoff(S0 - eps),
 *   dd = cp 1{..} S_T / S0 evaluated at S0 + eps minus the same at S0 - eps (the pathwise delta of each bumped solve). */
int hho_mc_european_second_sums(const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                                double spot_bump, double *second_sums) {
  int rc = check_args(model, sim);
  if (rc) return rc;
  if (npayoffs <= 0 || !payoffs || !second_sums || !(spot_bump > 0.0) || !(spot_bump < model->S0)) return HH_ERR_ARG;
  const int64_t N = sim->n_paths;
  const int anti = sim->vr == HH_VR_ANTITHETIC;
  hh_model mu = *model, md = *model;
  mu.S0 = model->S0 + spot_bump;
  md.S0 = model->S0 - spot_bump;
  const size_t tot = (size_t)npayoffs * 4;
  long double *acc = calloc(tot, sizeof(long double));
#pragma omp parallel
  {
    long double *la = calloc(tot, sizeof(long double));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
      terminal_t t0 = simulate_one(model, sim, i, NULL, NULL, 0);
      terminal_t tu = simulate_one(&mu, sim, i, NULL, NULL, 0);
      terminal_t td = simulate_one(&md, sim, i, NULL, NULL, 0);
      for (int k = 0; k < npayoffs; ++k) {
        const hh_payoff *po = &payoffs[k];
        double sd = payoff_of(po, tu.Sp) - 2.0 * payoff_of(po, t0.Sp) + payoff_of(po, td.Sp);
        double dd = (po->cp * (tu.Sp - po->strike) > 0 ? po->cp * tu.Sp / mu.S0 : 0.0) -
                    (po->cp * (td.Sp - po->strike) > 0 ? po->cp * td.Sp / md.S0 : 0.0);
        if (anti) {
          double sdm = payoff_of(po, tu.Sm) - 2.0 * payoff_of(po, t0.Sm) + payoff_of(po, td.Sm);
          double ddm = (po->cp * (tu.Sm - po->strike) > 0 ? po->cp * tu.Sm / mu.S0 : 0.0) -
                       (po->cp * (td.Sm - po->strike) > 0 ? po->cp * td.Sm / md.S0 : 0.0);
          sd = (sd + sdm) / 2;
          dd = (dd + ddm) / 2;
        }
        long double *a = la + (size_t)k * 4;
        a[0] += sd;
        a[1] += (long double)sd * sd;
        a[2] += dd;
        a[3] += (long double)dd * dd;
      }
    }
#pragma omp critical
    for (size_t j = 0; j < tot; ++j) acc[j] += la[j];
    free(la);
  }
  for (size_t j = 0; j < tot; ++j) second_sums[j] = (double)acc[j];
  free(acc);
  return HH_OK;
}

/* ------------------------------------------------------------------------------------------
 * Longstaff-Schwartz: lsm.jl:99-165. Regression = Polynomials.fit(x, y, degree) [upstream:
 * monomial Vandermonde least squares solved by QR]; here Householder QR on the raw Vandermonde.
 * ---------------------------------------------------------------------------------------- */
static void lstsq_householder(double *A, double *b, int64_t m, int n, double *x) {
  /* A: column-major m x n (overwritten), b: m (overwritten). Minimises |A x - b|. */
  for (int k = 0; k < n; ++k) {
    double *ak = A + (size_t)k * m;
    long double nrm2 = 0;
    for (int64_t i = k; i < m; ++i) nrm2 += (long double)ak[i] * ak[i];
    double nrm = (double)sqrtl(nrm2);
    if (nrm == 0.0) continue;
    double alpha = ak[k] > 0 ? -nrm : nrm;
    double v0 = ak[k] - alpha;
    /* v = (v0, a[k+1..]) ; beta = 2 / (v'v) */
    long double vtv = (long double)v0 * v0 + (nrm2 - (long double)ak[k] * ak[k]);
    if (vtv <= 0) continue;
    double beta = (double)(2.0L / vtv);
    for (int j = k + 1; j <= n; ++j) {
      double *aj = j < n ? A + (size_t)j * m : b;
      long double dot = (long double)v0 * aj[k];
      for (int64_t i = k + 1; i < m; ++i) dot += (long double)ak[i] * aj[i];
      double f = beta * (double)dot;
      aj[k] -= f * v0;
      for (int64_t i = k + 1; i < m; ++i) aj[i] -= f * ak[i];
    }
    ak[k] = alpha;
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[(size_t)j * m + k] * x[j];
    double rkk = A[(size_t)k * m + k];
    x[k] = rkk != 0.0 ? s / rkk : 0.0;
  }
}

int hho_lsm_backward(const double *grid, int64_t ncols, int n_steps, const hh_payoff *payoff, int degree,
                     double step_discount, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                     double *beta_out) {
  if (!grid || ncols <= 0 || n_steps <= 0 || degree < 0 || !payoff || !out) return HH_ERR_ARG;
  const int M = n_steps, nb = degree + 1;
  const double D = step_discount;
  int32_t *tau = stop_idx ? stop_idx : malloc(sizeof(int32_t) * (size_t)ncols);
  double *val = stop_val ? stop_val : malloc(sizeof(double) * (size_t)ncols);
  double *A = malloc(sizeof(double) * (size_t)ncols * nb);
  double *y = malloc(sizeof(double) * (size_t)ncols);
  int64_t *itm = malloc(sizeof(int64_t) * (size_t)ncols);
  double beta[32];
  int64_t skipped = 0;
  if (beta_out) memset(beta_out, 0, sizeof(double) * (size_t)(M + 1) * nb);

  const double *GM = grid + (size_t)M * ncols;
  for (int64_t p = 0; p < ncols; ++p) { /* lsm.jl:112 */
    tau[p] = M;
    val[p] = payoff_of(payoff, GM[p]);
  }
  for (int t = M - 1; t >= 1; --t) { /* lsm.jl:114-115, i = t+1 */
    const double *Gt = grid + (size_t)t * ncols;
    int64_t n = 0;
    for (int64_t p = 0; p < ncols; ++p)
      if (payoff_of(payoff, Gt[p]) > 0) itm[n++] = p; /* lsm.jl:120-121 */
    if (n == 0) { skipped++; continue; }             /* lsm.jl:122 */
    for (int64_t q = 0; q < n; ++q) {
      int64_t p = itm[q];
      double xq = Gt[p], pw = 1.0;
      for (int k = 0; k < nb; ++k) { A[(size_t)k * n + q] = pw; pw *= xq; }
      y[q] = pow(D, (double)(tau[p] - t)) * val[p]; /* lsm.jl:117-118 */
    }
    lstsq_householder(A, y, n, nb, beta); /* lsm.jl:126 */
    if (beta_out) memcpy(beta_out + (size_t)t * nb, beta, sizeof(double) * nb);
    for (int64_t q = 0; q < n; ++q) {
      int64_t p = itm[q];
      double xq = Gt[p];
      double c = beta[nb - 1];
      for (int k = nb - 2; k >= 0; --k) c = c * xq + beta[k]; /* poly.(x) lsm.jl:127 */
      double e = payoff_of(payoff, xq);
      if (e > c) { tau[p] = t; val[p] = e; } /* strict, lsm.jl:163-164 */
    }
  }
  long double s = 0, sq = 0;
  for (int64_t p = 0; p < ncols; ++p) { /* lsm.jl:132-133 */
    double d = pow(D, (double)tau[p]) * val[p];
    s += d;
    sq += (long double)d * d;
  }
  memset(out, 0, sizeof(*out));
  out->sum = (double)s;
  out->sumsq = (double)sq;
  out->n = ncols;
  out->price = (double)(s / ncols);
  long double var = ncols > 1 ? (sq - s * s / ncols) / (ncols - 1) : 0;
  out->std_error = (double)sqrtl((var > 0 ? var : 0) / ncols);
  out->n_dates_skipped = skipped;
  if (!stop_idx) free(tau);
  if (!stop_val) free(val);
  free(A);
  free(y);
  free(itm);
  return HH_OK;
}

int hho_lsm_american(const hh_model *model, const hh_sim *sim, const hh_payoff *payoff, int degree,
                     double step_discount, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                     double *spot_paths, double *beta_out) {
  int rc = check_args(model, sim);
  if (rc) return rc;
  /* Q7: LSM reads component 1 of the saved state as the spot (lsm.jl:53) — only the S-space
   * BlackScholesExact generator is meaningful in the reference. The log-space Euler-Maruyama schemes are
   * accepted here with the corrected extraction S = exp(x), binary64 only. */
  if (!(model->kind == HH_MODEL_GBM && sim->scheme == HH_SCHEME_EXACT_STEPS) && sim->scheme != HH_SCHEME_EM)
    return HH_ERR_UNSUPPORTED;
  if (sim->precision != HH_PREC_F64) return HH_ERR_UNSUPPORTED;
  const int64_t N = sim->n_paths;
  const int anti = sim->vr == HH_VR_ANTITHETIC;
  const int64_t ncols = anti ? 2 * N : N;
  const int M = sim->n_steps;
  double *grid = malloc(sizeof(double) * (size_t)(M + 1) * ncols);
  if (!grid) return HH_ERR_NOMEM;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i)
    simulate_one(model, sim, i, grid + i, anti ? grid + N + i : NULL, (size_t)ncols);
  rc = hho_lsm_backward(grid, ncols, M, payoff, degree, step_discount, out, stop_idx, stop_val, beta_out);
  if (spot_paths) /* Matrix (nsteps+1) x ncols, column-major: lsm.jl:50 */
    for (int64_t p = 0; p < ncols; ++p)
      for (int t = 0; t <= M; ++t) spot_paths[(size_t)p * (M + 1) + t] = grid[(size_t)t * ncols + p];
  free(grid);
  return rc;
}
