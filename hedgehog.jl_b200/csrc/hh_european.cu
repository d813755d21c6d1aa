// hh_european.cu — European Monte Carlo: path simulation + payoff + reduction in ONE kernel,
// for prices (T = double) and for pathwise forward-mode Greeks (T = Dual<P>).
//
// Replaces, for solve(::PricingProblem, ::MonteCarlo) (reference src/pricing_methods/montecarlo.jl:478-493):
//   sde_problem / simulate_paths (:140-231, :342-375)  -> in-register stepping, Philox normals in-kernel
//   final_sample (:384-402)                             -> exp of the log state at expiry
//   reduce_payoffs + mean (:428-432, :490)              -> fused block reduction of sum / sum of squares
// and, for solve(::GreekProblem, ::ForwardAD, method) (src/greeks/greeks_problem.jl:249-262), the Dual that
// ForwardDiff pushes through all of the above.
//
// Layout: one trajectory (or antithetic pair) per thread, state in registers for all steps; no HBM
// traffic in the step loop. After a batch of 256 trajectories the block evaluates every payoff on the
// batch from shared memory ("payoff transpose": thread = (strike, path group)), so up to 256 strikes are
// priced on the same paths with register accumulators and a fixed summation order.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "hh_ctx.h"
#include "hh_fastnormal.cuh"
#include "hh_paths.cuh"

namespace hh {

constexpr int kThreads = 256;
constexpr int kMaxTan = 8;

enum : int { K_GBM_EM = 0, K_GBM_TERMINAL = 1, K_GBM_STEPS = 2, K_HESTON_EM = 3 };

struct TangentPack {  // d(parameter)/d(direction p), zero-padded to the kernel's P
  double x0[kMaxTan], S0[kMaxTan], dt_drift[kMaxTan], sigma[kMaxTan], sig_sqdt[kMaxTan], mu[kMaxTan], sd[kMaxTan];
  double v0[kMaxTan], r[kMaxTan], kappa[kMaxTan], theta[kMaxTan], xi[kMaxTan];
  double a11[kMaxTan], a12[kMaxTan], a21[kMaxTan], a22[kMaxTan];
};

struct EuroArgs {
  int64_t n, path_offset;
  uint64_t base_seed;
  const uint64_t *seeds;
  const double *normals;
  double *terminal;
  const hh_payoff *payoffs;
  double *partials;  // [grid][npay][nacc]
  int npay, kp_log2, n_steps, split, parity, f32;
  PathParams<double> p;
  HestonFolded f;
  PhiloxRoundKeys rk;  // round keys of base_seed (uniform across threads when seeds == NULL)
  // bit patterns handed over as ARGUMENTS so that (x & mask) | pattern stays one LOP3 with a uniform-register operand
  // (as literals the compiler splits it into an AND-immediate and an OR-immediate)
  uint32_t one_hi, magic_hi;  // 0x3FF00000 (high word of 1.0), 0x43300000 (high word of 2^52)
  uint32_t f32_one;           // 0x3F800000 (1.0f)
  uint32_t lo_fill;           // 0x00080000: the half-ulp bit behind the 32 random mantissa bits (HH_RNG_PHILOX_64)
  int rng64;                  // HH_RNG_PHILOX_64
  // Second order in the spot on the SAME trajectories (SecondOrderGreekProblem(spot, spot), greeks_problem.jl:395-412: absolute
  // bump eps, common random numbers). Every scheme here is linear in S0 (log-space state, or S-space steps that multiply S),
  // so a re-solve at S0 +- eps with the same seeds has terminal spots S_T (S0 +- eps) / S0: the bumped payoffs are evaluated
  // in the payoff stage of the one simulation. Per trajectory (pair-averaged when antithetic):
  //   sd = pay(S_T up) - 2 pay(S_T) + pay(S_T dn)                         -> gamma_fd = D mean(sd) / eps^2   (the reference's form)
  //   dd = delta_path(S0 + eps) - delta_path(S0 - eps),  delta_path = cp 1{cp (S - K) > 0} S / S0   -> gamma_pw = D mean(dd) / (2 eps)
  int g_on;
  double g_up, g_dn, g_iu, g_id;  // (S0 + eps) / S0, (S0 - eps) / S0, 1 / (S0 + eps), 1 / (S0 - eps)
  // HH_VR_QUASI_RANDOM (K_GBM_TERMINAL): randomised van der Corput points instead of Philox normals
  int qmc;
  uint64_t qmc_shift;  // Cranley-Patterson rotation on the 2^-64 grid, from base_seed (qmc_shift_of)
};

// splitmix64 (Steele, Lea & Flood 2014) of the seed: the rotation of the quasi-random points. Restated in oracle/oracle.py.
__host__ __device__ inline uint64_t qmc_shift_of(uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// second-order contributions of one terminal spot at one strike
__device__ __forceinline__ void gamma_terms(const EuroArgs &a, double sp, double ep, double cp, double strike, double &s2, double &dd) {
  const double su = sp * a.g_up, sl = sp * a.g_dn;
  const double eu = cp * (su - strike), el = cp * (sl - strike);
  s2 = fmax(eu, 0.0) - 2.0 * fmax(ep, 0.0) + fmax(el, 0.0);
  dd = (eu > 0.0 ? cp * su * a.g_iu : 0.0) - (el > 0.0 ? cp * sl * a.g_id : 0.0);
}

// number of accumulators per (block, payoff): sum, sumsq, nonfinite, then per tangent: dsum, dsumsq, and — in the tangent
// kernels — the four second-order sums of the spot bump (see GammaArgs)
constexpr int kNGamma = 4;
template <class T>
__host__ __device__ constexpr int nacc_of() { return 3 + 2 * num_traits<T>::ntan + (num_traits<T>::ntan > 0 ? kNGamma : 0); }

template <class T>
__device__ __forceinline__ T lift(double v, const double *dv) {
  if constexpr (num_traits<T>::ntan == 0) {
    return v;
  } else {
    T r;
    r.v = v;
#pragma unroll
    for (int i = 0; i < num_traits<T>::ntan; ++i) r.d[i] = dv[i];
    return r;
  }
}

template <class T>
__device__ __forceinline__ T zero() {
  if constexpr (num_traits<T>::ntan == 0) {
    return 0.0;
  } else {
    T r;
    r.v = 0.0;
#pragma unroll
    for (int i = 0; i < num_traits<T>::ntan; ++i) r.d[i] = 0.0;
    return r;
  }
}

template <class T>
__device__ __forceinline__ PathParams<T> lift_params(const PathParams<double> &p, const TangentPack *t) {
  if constexpr (num_traits<T>::ntan == 0) {
    return p;
  } else {
    PathParams<T> q;
    q.dt = p.dt;
    q.sqdt = p.sqdt;
    q.x0 = lift<T>(p.x0, t->x0);
    q.S0 = lift<T>(p.S0, t->S0);
    q.dt_drift = lift<T>(p.dt_drift, t->dt_drift);
    q.sigma = lift<T>(p.sigma, t->sigma);
    q.sig_sqdt = lift<T>(p.sig_sqdt, t->sig_sqdt);
    q.mu = lift<T>(p.mu, t->mu);
    q.sd = lift<T>(p.sd, t->sd);
    q.v0 = lift<T>(p.v0, t->v0);
    q.r = lift<T>(p.r, t->r);
    q.kappa = lift<T>(p.kappa, t->kappa);
    q.theta = lift<T>(p.theta, t->theta);
    q.xi = lift<T>(p.xi, t->xi);
    q.a11 = lift<T>(p.a11, t->a11);
    q.a12 = lift<T>(p.a12, t->a12);
    q.a21 = lift<T>(p.a21, t->a21);
    q.a22 = lift<T>(p.a22, t->a22);
    return q;
  }
}

// Where a trajectory's standard normals come from: in-kernel Philox (native) or the caller's buffer (parity).
struct NormalSource {
  uint64_t key, idx;
  const double *z;
  const FastNormalTables *tb;
  __device__ __forceinline__ NormalSource(const EuroArgs &a, const FastNormalTables *tables, bool parity, int64_t i,
                                          int per_path) {
    z = nullptr;
    tb = tables;
    key = idx = 0;
    if (parity) {
      z = a.normals + (size_t)i * (size_t)per_path;
    } else if (a.seeds) {
      key = a.seeds[i];
    } else {
      key = a.base_seed;
      idx = (uint64_t)(a.path_offset + i);
    }
  }
  // both normals of Philox block n (parity: elements 2n, 2n+1 of this trajectory's slice)
  __device__ __forceinline__ void pair(bool parity, int n, double &z1, double &z2) const {
    if (parity) {
      z1 = z[2 * n];
      z2 = z[2 * n + 1];
    } else {
      const u32x4 w = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)n, 0u, (uint32_t)key,
                                    (uint32_t)(key >> 32));
      fast_normal_pair(tb, w.x, w.y, w.z, w.w, z1, z2);
    }
  }
};

// ---- one trajectory (or antithetic pair) -> terminal spot(s) in number type T ---------------------------
template <int KIND, class T, bool ANTI, bool FAST>
__device__ __forceinline__ void simulate(const EuroArgs &a, const PathParams<T> &p, const FastNormalTables *tb,
                                         bool parity, bool split, int64_t i, T &Sp, T &Sm) {
  const int M = a.n_steps;
  if constexpr (KIND == K_GBM_TERMINAL) {
    // marginal_law + final_sample, montecarlo.jl:293-303, 384-390
    NormalSource src(a, tb, parity, i, 1);
    double z1, z2;
    if (parity) {
      z1 = src.z[0];
    } else if (a.qmc) {
      // u = frac(van der Corput_2(g) + shift) on the 2^-64 grid (the addition wraps), taken at the midpoints of the
      // 2^-53 grid so that 0 < u < 1; g = global trajectory index
      const uint64_t pt = __brevll((uint64_t)(a.path_offset + i)) + a.qmc_shift;
      z1 = normcdfinv(((double)(pt >> 11) + 0.5) * 0x1p-53);
    } else {
      src.pair(false, 0, z1, z2);
    }
    const T X = fma_(p.sd, z1, p.mu);
    Sp = exp_(X);
    if (ANTI) Sm = exp_(p.mu * 2.0 - X);
  } else if constexpr (KIND == K_GBM_EM || KIND == K_GBM_STEPS) {
    NormalSource src(a, tb, parity, i, M);
    T sp = KIND == K_GBM_EM ? p.x0 : p.S0;
    T sm = sp;
#pragma unroll 1
    for (int n = 0; n < M; n += 2) {
      double za, zb;
      if (parity) {
        za = src.z[n];
        zb = n + 1 < M ? src.z[n + 1] : 0.0;
      } else {
        src.pair(false, n >> 1, za, zb);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (n + h < M) {
          const double z = h ? zb : za;
          if constexpr (KIND == K_GBM_EM) {
            const double dW = p.sqdt * z;
            gbm_em_step(p, sp, dW);
            if (ANTI) gbm_em_step(p, sm, -dW);  // NoiseGrid(t, -W) montecarlo.jl:258
          } else {
            gbm_exact_step(p, sp, z, 1.0);
            if (ANTI) gbm_exact_step(p, sm, z, -1.0);  // sigma -> -sigma, same normals (montecarlo.jl:270-284)
          }
        }
      }
    }
    if constexpr (KIND == K_GBM_EM) {
      Sp = exp_(sp);  // final_sample montecarlo.jl:398
      if (ANTI) Sm = exp_(sm);
    } else {
      Sp = sp;
      if (ANTI) Sm = sm;
    }
  } else {
    NormalSource src(a, tb, parity, i, 2 * M);
    T xp = p.x0, vp = p.v0, xm = p.x0, vm = p.v0;
    if constexpr (FAST) {
      // native-RNG pricing: constants pre-folded on the host, branch-free sqrt (hh_fastnormal.cuh)
#pragma unroll 1
      for (int n = 0; n < M; ++n) {
        double z1, z2;
        src.pair(false, n, z1, z2);
        const double dW1 = fma(p.a12, z2, p.a11 * z1);
        const double dW2 = fma(a.f.b22, z2, a.f.b21 * z1);  // xi * dW2
        heston_em_step_fast(a.f, split, xp, vp, dW1, dW2);
        if (ANTI) heston_em_step_fast(a.f, split, xm, vm, -dW1, -dW2);
      }
    } else {
#pragma unroll 1
      for (int n = 0; n < M; ++n) {
        double z1, z2;
        src.pair(parity, n, z1, z2);
        const T dW1 = fma_(p.a12, z2, p.a11 * z1);
        const T dW2 = fma_(p.a22, z2, p.a21 * z1);
        heston_em_step(p, split, xp, vp, dW1, dW2);
        if (ANTI) heston_em_step(p, split, xm, vm, -dW1, -dW2);
      }
    }
    Sp = exp_(xp);
    if (ANTI) Sm = exp_(xm);
  }
}

// PARITY / SPLIT: 0 or 1 = compile-time, 2 = read from the arguments (used by the tangent kernels to
// keep the number of instantiations down).
template <int KIND, class T, bool ANTI, int PARITY, int SPLIT>
__global__ void __launch_bounds__(kThreads) european_kernel(const EuroArgs a, const TangentPack *__restrict__ tpack) {
  constexpr int NT = num_traits<T>::ntan;
  constexpr int NACC = nacc_of<T>();
  constexpr int NV = 1 + NT;                 // values per terminal state: S, dS_1..dS_NT
  constexpr int NSIDE = ANTI ? 2 : 1;
  constexpr int STAGE = NV * NSIDE * kThreads;
  constexpr int RCH = NACC < 19 ? NACC : 19;  // accumulators reduced per pass (keeps the static shared memory under 48 KB)
  constexpr int RED = RCH * kThreads;
  __shared__ double smem[STAGE > RED ? STAGE : RED];

  __shared__ FastNormalTables s_tables;  // Box-Muller lookup tables (native-RNG mode)
  if (PARITY != 1) {
    load_fast_tables(&s_tables);
    __syncthreads();
  }
  constexpr bool FAST = KIND == K_HESTON_EM && NT == 0 && PARITY == 0;
  const bool parity = PARITY == 2 ? a.parity != 0 : PARITY == 1;
  const bool split = SPLIT == 2 ? a.split != 0 : SPLIT == 1;
  const PathParams<T> p = lift_params<T>(a.p, tpack);

  const int tid = threadIdx.x;
  const int KP = 1 << a.kp_log2;        // payoffs padded to a power of two <= 256
  const int k = tid & (KP - 1);         // my strike
  const int g = tid >> a.kp_log2;       // my path group
  const int G = kThreads >> a.kp_log2;  // number of path groups
  double strike = 0.0, cp = 0.0;
  if (k < a.npay) {
    strike = a.payoffs[k].strike;
    cp = a.payoffs[k].cp;
  }
  double acc[NACC];
#pragma unroll
  for (int c = 0; c < NACC; ++c) acc[c] = 0.0;

  for (int64_t base = (int64_t)blockIdx.x * kThreads; base < a.n; base += (int64_t)gridDim.x * kThreads) {
    const int64_t i = base + tid;
    T Sp = zero<T>(), Sm = zero<T>();
    if (i < a.n) {
      simulate<KIND, T, ANTI, FAST>(a, p, &s_tables, parity, split, i, Sp, Sm);
      if (a.terminal) {  // MonteCarloSolution.ensemble: (plus | minus), montecarlo.jl:400-402, 492
        a.terminal[i] = value(Sp);
        if (ANTI) a.terminal[a.n + i] = value(Sm);
      }
    }
    smem[tid] = value(Sp);
#pragma unroll
    for (int q = 0; q < NT; ++q) smem[(1 + q) * kThreads + tid] = tangent(Sp, q);
    if (ANTI) {
      smem[NV * kThreads + tid] = value(Sm);
#pragma unroll
      for (int q = 0; q < NT; ++q) smem[(NV + 1 + q) * kThreads + tid] = tangent(Sm, q);
    }
    __syncthreads();
    const int64_t rem = a.n - base;
    const int nvalid = rem < kThreads ? (int)rem : kThreads;
    if (k < a.npay) {
      for (int j = g; j < nvalid; j += G) {
        const double sp = smem[j];
        const double ep = cp * (sp - strike);
        double pay = fmax(ep, 0.0);  // payoffs.jl:154-156
        const double ip = ep > 0.0 ? cp : 0.0;
        bool bad = !isfinite(sp);
        double im = 0.0;
        if (ANTI) {
          const double sm = smem[NV * kThreads + j];
          const double em = cp * (sm - strike);
          pay = 0.5 * (pay + fmax(em, 0.0));  // reduce_payoffs montecarlo.jl:430-432
          im = em > 0.0 ? cp : 0.0;
          bad = bad || !isfinite(sm);
        }
        acc[0] += pay;
        acc[1] = fma(pay, pay, acc[1]);
        if (k == 0 && bad) acc[2] += 1.0;
#pragma unroll
        for (int q = 0; q < NT; ++q) {
          double dpay = ip * smem[(1 + q) * kThreads + j];  // cp 1{cp(S-K)>0} dS
          if (ANTI) dpay = 0.5 * (dpay + im * smem[(NV + 1 + q) * kThreads + j]);
          acc[3 + q] += dpay;
          acc[3 + NT + q] = fma(dpay, dpay, acc[3 + NT + q]);
        }
        if constexpr (NT > 0) {
          if (a.g_on) {
            double s2, dd;
            gamma_terms(a, sp, ep, cp, strike, s2, dd);
            if (ANTI) {
              const double sm = smem[NV * kThreads + j];
              double s2m, ddm;
              gamma_terms(a, sm, cp * (sm - strike), cp, strike, s2m, ddm);
              s2 = 0.5 * (s2 + s2m);
              dd = 0.5 * (dd + ddm);
            }
            acc[3 + 2 * NT + 0] += s2;
            acc[3 + 2 * NT + 1] = fma(s2, s2, acc[3 + 2 * NT + 1]);
            acc[3 + 2 * NT + 2] += dd;
            acc[3 + 2 * NT + 3] = fma(dd, dd, acc[3 + 2 * NT + 3]);
          }
        }
      }
    }
    __syncthreads();
  }

  // fixed-order reduction over the path groups that share a strike, RCH accumulators per pass
#pragma unroll
  for (int c0 = 0; c0 < NACC; c0 += RCH) {
#pragma unroll
    for (int c = 0; c < RCH; ++c)
      if (c0 + c < NACC) smem[c * kThreads + tid] = acc[c0 + c < NACC ? c0 + c : 0];
    __syncthreads();
    if (tid < a.npay) {
      double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * NACC;
      for (int c = 0; c < RCH && c0 + c < NACC; ++c) {
        double t = 0.0;
        for (int gg = 0; gg < G; ++gg) t += smem[c * kThreads + (gg << a.kp_log2) + tid];
        out[c0 + c] = t;
      }
    }
    __syncthreads();
  }
}

#ifdef HH_TUNING  // the first version of the headline kernel (tables not replicated): kept for A/B timing only
// ---- the headline kernel: Heston Euler-Maruyama, Float64, native RNG (config C2) -----------------------------
// Same trajectory arithmetic and payoff transpose as european_kernel, specialised for throughput:
//   UKEY  the Philox round keys are uniform (base_seed mode) and come from the kernel arguments (constant bank),
//         which removes 20 registers of key schedule per thread;
//   ILP   trajectories advanced per thread in lock step (independent dependency chains for the FP64 pipe).
template <bool ANTI, bool SPLIT, bool UKEY, int ILP, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) heston_fast_kernel(const EuroArgs a) {
  constexpr int NSIDE = ANTI ? 2 : 1;
  constexpr int NACC = 3;
  constexpr int STAGE = NSIDE * kThreads;
  constexpr int RED = NACC * kThreads;
  __shared__ double smem[STAGE > RED ? STAGE : RED];
  __shared__ FastNormalTables s_tables;
  // phase table: {P1, Q1}, {P2, Q2} per table angle, P = m . (cos, sin), Q = m . (-sin, cos) for the rows
  // m1 = sqrt(dt) (m11, m12) and m2 = xi sqrt(dt) (m21, m22) of the Brownian factor
  __shared__ double2 s_phase[2 * tables::kTrigN];
  load_fast_tables(&s_tables);
  __syncthreads();
  for (int j = threadIdx.x; j < tables::kTrigN; j += kThreads) {
    const double2 cs = s_tables.trig_tab[j];
    s_phase[2 * j] = make_double2(fma(a.p.a12, cs.y, a.p.a11 * cs.x), fma(a.p.a12, cs.x, -(a.p.a11 * cs.y)));
    s_phase[2 * j + 1] = make_double2(fma(a.f.b22, cs.y, a.f.b21 * cs.x), fma(a.f.b22, cs.x, -(a.f.b21 * cs.y)));
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const int KP = 1 << a.kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> a.kp_log2;
  const int G = kThreads >> a.kp_log2;
  double strike = 0.0, cp = 0.0;
  if (k < a.npay) {
    strike = a.payoffs[k].strike;
    cp = a.payoffs[k].cp;
  }
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  const int M = a.n_steps;
  constexpr int64_t kBatch = (int64_t)kThreads * ILP;

  for (int64_t base = (int64_t)blockIdx.x * kBatch; base < a.n; base += (int64_t)gridDim.x * kBatch) {
    double xp[ILP], vp[ILP], xm[ILP], vm[ILP];
    uint32_t c0[ILP], c1[ILP];
    PhiloxRoundKeys rk[UKEY ? 1 : ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t i = base + (int64_t)j * kThreads + tid;
      const int64_t ic = i < a.n ? i : a.n - 1;  // tail lanes repeat the last trajectory; never accumulated
      xp[j] = xm[j] = a.p.x0;
      vp[j] = vm[j] = a.p.v0;
      if (UKEY) {
        const uint64_t idx = (uint64_t)(a.path_offset + ic);
        c0[j] = (uint32_t)idx;
        c1[j] = (uint32_t)(idx >> 32);
      } else {
        c0[j] = c1[j] = 0u;
        rk[UKEY ? 0 : j] = philox_round_keys(a.seeds[ic]);
      }
    }
#pragma unroll 1
    for (int n = 0; n < M; ++n) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        const u32x4 w = philox4x32_10_rk(c0[j], c1[j], (uint32_t)n, 0u, UKEY ? a.rk : rk[UKEY ? 0 : j]);
        // Box-Muller folded into the step: with (z1, z2) = rad (cos th, sin th),
        //   dW1 = rad c1,  xi dW2 = rad c2,  c_i = P_i cos(delta) + Q_i sin(delta)   (phase table: rotation x correlation)
        //   s dW = sqrt(K2+ R2) c        (ONE square root for diffusion and radius)
        const double R2 = fast_neg2log(&s_tables, w.x, w.y);
        double sn, cm;
        const uint32_t jt = fast_angle(w.z, w.w, sn, cm);
        const double2 pq1 = s_phase[2 * jt], pq2 = s_phase[2 * jt + 1];
        const double cc1 = fma(pq1.x, cm, fma(pq1.y, sn, pq1.x));
        const double cc2 = fma(pq2.x, cm, fma(pq2.y, sn, pq2.x));
        heston_em_step_folded(a.f, SPLIT, xp[j], vp[j], R2, cc1, cc2);
        if (ANTI) heston_em_step_folded(a.f, SPLIT, xm[j], vm[j], R2, -cc1, -cc2);
      }
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t sub = base + (int64_t)j * kThreads;
      if (sub >= a.n) break;
      const int64_t i = sub + tid;
      const double Sp = exp(xp[j]);
      const double Sm = ANTI ? exp(xm[j]) : 0.0;
      if (a.terminal && i < a.n) {
        a.terminal[i] = Sp;
        if (ANTI) a.terminal[a.n + i] = Sm;
      }
      smem[tid] = Sp;
      if (ANTI) smem[kThreads + tid] = Sm;
      __syncthreads();
      const int64_t rem = a.n - sub;
      const int nvalid = rem < kThreads ? (int)rem : kThreads;
      if (k < a.npay) {
        for (int q = g; q < nvalid; q += G) {
          const double sp = smem[q];
          double pay = fmax(cp * (sp - strike), 0.0);  // payoffs.jl:154-156
          bool bad = !isfinite(sp);
          if (ANTI) {
            const double sm = smem[kThreads + q];
            pay = 0.5 * (pay + fmax(cp * (sm - strike), 0.0));  // reduce_payoffs montecarlo.jl:430-432
            bad = bad || !isfinite(sm);
          }
          acc0 += pay;
          acc1 = fma(pay, pay, acc1);
          if (k == 0 && bad) acc2 += 1.0;
        }
      }
      __syncthreads();
    }
  }
  smem[tid] = acc0;
  smem[kThreads + tid] = acc1;
  smem[2 * kThreads + tid] = acc2;
  __syncthreads();
  if (tid < a.npay) {
    double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * NACC;
    for (int c = 0; c < NACC; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += smem[c * kThreads + (gg << a.kp_log2) + tid];
      out[c] = t;
    }
  }
}

#endif  // HH_TUNING

// ---- v2 of the headline kernel: lane-replicated tables in dynamic shared memory ---------------------------------
// Same trajectory arithmetic as heston_fast_kernel up to rounding (the drift r dt is added once at expiry, the clamps
// use one integer max each), but every table read of the step loop is bank-conflict free (hh_fastnormal.cuh, "v2").
// Dynamic shared memory: [payoff staging / reduction | log table x8 | phase table x8 | exponent table].
__host__ __device__ constexpr int fast2_stage_bytes(int threads) { return 3 * threads * 8; }
__host__ __device__ constexpr int fast2_smem_bytes(int threads) { return fast2_stage_bytes(threads) + kLogRepBytes + kPhaseRepBytes + kExp2Bytes; }

// One Euler-Maruyama step of one trajectory (and its antithetic partner) from the random words of the step:
// (w0..w3) = one Philox block under HH_RNG_PHILOX, (w0, w1) = half a block under HH_RNG_PHILOX_64 (R64).
template <bool ANTI, bool SPLIT, bool R64>
__device__ __forceinline__ void fast2_step(const EuroArgs &a, const char *__restrict__ log_lane,
                                           const char *__restrict__ exp_biased, const char *__restrict__ phase_lane,
                                           uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, double &xp, double &vp,
                                           double &xm, double &vm) {
  double sn, cs;
  const double R2 = R64 ? fast_neg2log_32(log_lane, exp_biased, w1, a.one_hi, a.lo_fill)
                        : fast_neg2log_v2(log_lane, exp_biased, w0, w1, a.one_hi);
  const uint32_t poff = R64 ? fast_angle_v2(w0, w0, a.magic_hi, sn, cs) : fast_angle_v2(w2, w3, a.magic_hi, sn, cs);
  const double2 pq1 = *reinterpret_cast<const double2 *>(phase_lane + poff);
  const double2 pq2 = *reinterpret_cast<const double2 *>(phase_lane + poff + kRep * 16);
  const double cc1 = fma(pq1.x, cs, pq1.y * sn);
  const double cc2 = fma(pq2.x, cs, pq2.y * sn);
  {
    const double vplus = max0_hi(vp);
    const double K1 = fma(a.f.neg_half_dt, vplus, xp);
    const double K2 = fma(a.f.neg_kdt, vplus, vp + a.f.ktdt);
    const double sr = fast_sqrt_pos5(max_tiny_hi((SPLIT ? K2 : vplus) * R2));
    xp = fma(sr, cc1, K1);
    vp = fma(sr, cc2, K2);
  }
  if (ANTI) {
    const double vplus = max0_hi(vm);
    const double K1 = fma(a.f.neg_half_dt, vplus, xm);
    const double K2 = fma(a.f.neg_kdt, vplus, vm + a.f.ktdt);
    const double sr = fast_sqrt_pos5(max_tiny_hi((SPLIT ? K2 : vplus) * R2));
    xm = fma(-sr, cc1, K1);
    vm = fma(-sr, cc2, K2);
  }
}

// R64: the HH_RNG_PHILOX_64 stream (one Philox block per two steps, hh_fastnormal.cuh).
// ABL: ablation for hh_bench_heston_ablation — 0 the product kernel; 1 Philox replaced by a counter hash (what the
// table-driven Box-Muller and the step cost alone); 2 Philox only (its words XOR-folded into the state, no FP64 work).
template <bool ANTI, bool SPLIT, bool UKEY, int ILP, int THREADS, int MINB, bool R64 = false, int ABL = 0>
__global__ void __launch_bounds__(THREADS, MINB) heston_fast2_kernel(const EuroArgs a) {
  constexpr int NACC = 3;
  constexpr int kThreads = THREADS;  // shadows the file-scope block size
  constexpr int kFast2StageBytes = fast2_stage_bytes(THREADS);
  extern __shared__ __align__(16) unsigned char dsm[];
  double *smem = reinterpret_cast<double *>(dsm);
  char *s_log = reinterpret_cast<char *>(dsm) + kFast2StageBytes;
  char *s_phase = s_log + kLogRepBytes;
  double *s_exp = reinterpret_cast<double *>(s_phase + kPhaseRepBytes);
  const int tid = threadIdx.x;
  for (int e = tid; e < tables::kLog2Buckets * kRep; e += kThreads)
    reinterpret_cast<double2 *>(s_log)[e] = g_fast_tables2.log_tab[e / kRep];
  for (int e = tid; e < tables::kTrigN * kRep; e += kThreads) {
    // phase table: {P1, Q1}, {P2, Q2} per table angle, P = m . (cos, sin), Q = m . (-sin, cos) for the rows
    // m1 = sqrt(dt) (m11, m12) and m2 = xi sqrt(dt) (m21, m22) of the Brownian factor
    const int j = e / kRep, q = e % kRep;
    const double2 cs = g_fast_tables2.trig_tab[j];
    double2 *dst = reinterpret_cast<double2 *>(s_phase) + (size_t)j * 2 * kRep + q;
    dst[0] = make_double2(fma(a.p.a12, cs.y, a.p.a11 * cs.x), fma(a.p.a12, cs.x, -(a.p.a11 * cs.y)));
    dst[kRep] = make_double2(fma(a.f.b22, cs.y, a.f.b21 * cs.x), fma(a.f.b22, cs.x, -(a.f.b21 * cs.y)));
  }
  for (int e = tid; e < tables::kExp2N; e += kThreads) s_exp[e] = g_fast_tables2.exp_tab[e];
  __syncthreads();
  const char *log_lane = s_log + (tid & (kRep - 1)) * 16;
  const char *phase_lane = s_phase + (tid & (kRep - 1)) * 16;
  const char *exp_biased = reinterpret_cast<const char *>(s_exp) - tables::kExp2Bias * 8;

  const int KP = 1 << a.kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> a.kp_log2;
  const int G = kThreads >> a.kp_log2;
  double strike = 0.0, cp = 0.0;
  if (k < a.npay) {
    strike = a.payoffs[k].strike;
    cp = a.payoffs[k].cp;
  }
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  const int M = a.n_steps;
  const double drift_total = (double)M * a.f.rdt;
  constexpr int64_t kBatch = (int64_t)kThreads * ILP;

  for (int64_t base = (int64_t)blockIdx.x * kBatch; base < a.n; base += (int64_t)gridDim.x * kBatch) {
    double xp[ILP], vp[ILP], xm[ILP], vm[ILP];
    uint32_t c0[ILP], c1[ILP];
    PhiloxRoundKeys rk[UKEY ? 1 : ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t i = base + (int64_t)j * kThreads + tid;
      const int64_t ic = i < a.n ? i : a.n - 1;  // tail lanes repeat the last trajectory; never accumulated
      xp[j] = xm[j] = a.p.x0;
      vp[j] = vm[j] = a.p.v0;
      if (UKEY) {
        const uint64_t idx = (uint64_t)(a.path_offset + ic);
        c0[j] = (uint32_t)idx;
        c1[j] = (uint32_t)(idx >> 32);
      } else {
        c0[j] = c1[j] = 0u;
        rk[UKEY ? 0 : j] = philox_round_keys(a.seeds[ic]);
      }
    }
    uint32_t fold[ILP];  // ABL == 2 only
#pragma unroll
    for (int j = 0; j < ILP; ++j) fold[j] = 0u;
    constexpr int kStepsPerBlock = R64 ? 2 : 1;
    constexpr uint32_t kStreamWord = R64 ? 2u : 0u;  // f32 fast mode: 1
#pragma unroll 1
    for (int n = 0; n < M; n += kStepsPerBlock) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        u32x4 w;
        if (ABL == 1) {  // no Philox: a counter hash, so that the table indices still vary
          const uint32_t t = (c0[j] + (uint32_t)n) * 0x9E3779B9u;
          w.x = t;
          w.y = (t >> 7) ^ (uint32_t)n;
          w.z = t ^ 0x5555AAAAu;
          w.w = ~t;
        } else {
          w = philox4x32_10_rk(c0[j], c1[j], (uint32_t)(R64 ? n >> 1 : n), kStreamWord, UKEY ? a.rk : rk[UKEY ? 0 : j]);
        }
        if (ABL == 2) {
          fold[j] ^= w.x ^ w.y ^ w.z ^ w.w;
          continue;
        }
        // Box-Muller folded into the step: with (z1, z2) = rad (cos th, sin th),
        //   dW1 = rad c1,  xi dW2 = rad c2,  c_i = P_i cos(delta) + Q_i sin(delta)   (phase table: rotation x correlation)
        //   s dW = sqrt(K2+ R2) c        (ONE square root for diffusion and radius)
        fast2_step<ANTI, SPLIT, R64>(a, log_lane, exp_biased, phase_lane, w.x, w.y, w.z, w.w, xp[j], vp[j], xm[j], vm[j]);
        if (R64 && n + 1 < M)
          fast2_step<ANTI, SPLIT, true>(a, log_lane, exp_biased, phase_lane, w.z, w.w, 0u, 0u, xp[j], vp[j], xm[j], vm[j]);
      }
    }
    if (ABL == 2) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) xp[j] = __hiloint2double((int)(fold[j] & 0x000FFFFFu) | 0x3FF00000, (int)fold[j]);
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t sub = base + (int64_t)j * kThreads;
      if (sub >= a.n) break;
      const int64_t i = sub + tid;
      const double Sp = exp(xp[j] + drift_total);
      const double Sm = ANTI ? exp(xm[j] + drift_total) : 0.0;
      if (a.terminal && i < a.n) {
        a.terminal[i] = Sp;
        if (ANTI) a.terminal[a.n + i] = Sm;
      }
      smem[tid] = Sp;
      if (ANTI) smem[kThreads + tid] = Sm;
      __syncthreads();
      const int64_t rem = a.n - sub;
      const int nvalid = rem < kThreads ? (int)rem : kThreads;
      if (k < a.npay) {
        for (int q = g; q < nvalid; q += G) {
          const double sp = smem[q];
          double pay = fmax(cp * (sp - strike), 0.0);  // payoffs.jl:154-156
          bool bad = !isfinite(sp);
          if (ANTI) {
            const double sm = smem[kThreads + q];
            pay = 0.5 * (pay + fmax(cp * (sm - strike), 0.0));  // reduce_payoffs montecarlo.jl:430-432
            bad = bad || !isfinite(sm);
          }
          acc0 += pay;
          acc1 = fma(pay, pay, acc1);
          if (k == 0 && bad) acc2 += 1.0;
        }
      }
      __syncthreads();
    }
  }
  smem[tid] = acc0;
  smem[kThreads + tid] = acc1;
  smem[2 * kThreads + tid] = acc2;
  __syncthreads();
  if (tid < a.npay) {
    double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * NACC;
    for (int c = 0; c < NACC; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += smem[c * kThreads + (gg << a.kp_log2) + tid];
      out[c] = t;
    }
  }
}

// ---- LognormalDynamics pricing kernels (configs C1-like at scale): GBM Euler-Maruyama and exact steps ---------------
// The generic european_kernel carries the number-type template and the v1 tables; for plain prices with the in-kernel
// RNG this kernel uses the lane-replicated v2 tables, the uniform round keys, one Philox block per TWO steps (a GBM step
// needs one normal), the table-driven exp(y) - 1 for the exact step, and ILP 2. Same trajectories as the generic
// kernel and the oracle up to rounding (the log-space drift is added once at expiry).
__host__ __device__ constexpr int gbm_fast_smem(bool anti, int threads, bool steps) {
  return 3 * threads * 8 * (anti ? 2 : 1) + kLogRepBytes + kTrigRepBytes + kExp2Bytes + (steps ? kExpm1TabBytes : 0);
}

template <int KIND, bool ANTI, bool UKEY, int ILP, int THREADS>
__global__ void __launch_bounds__(THREADS) gbm_fast_kernel(const EuroArgs a) {
  static_assert(KIND == K_GBM_EM || KIND == K_GBM_STEPS, "terminal sampling stays in the generic kernel");
  constexpr int kThreads = THREADS;  // shadows the file-scope block size
  constexpr int NACC = 3;
  constexpr int NSIDE = ANTI ? 2 : 1;
  extern __shared__ __align__(16) unsigned char dsm[];
  double *smem = reinterpret_cast<double *>(dsm);  // 3 * kThreads * NSIDE doubles: staging, then the final reduction
  char *s_log = reinterpret_cast<char *>(dsm) + 3 * kThreads * 8 * NSIDE;
  char *s_trig = s_log + kLogRepBytes;
  double *s_e2 = reinterpret_cast<double *>(s_trig + kTrigRepBytes);
  double2 *s_exp = reinterpret_cast<double2 *>(reinterpret_cast<char *>(s_e2) + kExp2Bytes);
  const int tid = threadIdx.x;
  for (int e = tid; e < tables::kLog2Buckets * kRep; e += kThreads)
    reinterpret_cast<double2 *>(s_log)[e] = g_fast_tables2.log_tab[e / kRep];
  for (int e = tid; e < tables::kTrigN * kRep; e += kThreads)
    reinterpret_cast<double2 *>(s_trig)[e] = g_fast_tables2.trig_tab[e / kRep];
  for (int e = tid; e < tables::kExp2N; e += kThreads) s_e2[e] = g_fast_tables2.exp_tab[e];
  if (KIND == K_GBM_STEPS) fill_expm1_table(s_exp);
  __syncthreads();
  const char *log_lane = s_log + (tid & (kRep - 1)) * 16;
  const char *trig_lane = s_trig + (tid & (kRep - 1)) * 16;
  const char *exp_biased = reinterpret_cast<const char *>(s_e2) - tables::kExp2Bias * 8;
  const int KP = 1 << a.kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> a.kp_log2;
  const int G = kThreads >> a.kp_log2;
  double strike = 0.0, cp = 0.0;
  if (k < a.npay) {
    strike = a.payoffs[k].strike;
    cp = a.payoffs[k].cp;
  }
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  const int M = a.n_steps;
  const double sig = a.p.sig_sqdt, drift = a.p.dt_drift;
  const double drift_total = (double)M * drift;
  constexpr int64_t kBatch = (int64_t)kThreads * ILP;

  for (int64_t base = (int64_t)blockIdx.x * kBatch; base < a.n; base += (int64_t)gridDim.x * kBatch) {
    double sp[ILP], sm[ILP];
    uint32_t c0[ILP], c1[ILP];
    PhiloxRoundKeys rk[UKEY ? 1 : ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t i = base + (int64_t)j * kThreads + tid;
      const int64_t ic = i < a.n ? i : a.n - 1;
      sp[j] = sm[j] = KIND == K_GBM_EM ? a.p.x0 : a.p.S0;
      if (UKEY) {
        const uint64_t idx = (uint64_t)(a.path_offset + ic);
        c0[j] = (uint32_t)idx;
        c1[j] = (uint32_t)(idx >> 32);
      } else {
        c0[j] = c1[j] = 0u;
        rk[UKEY ? 0 : j] = philox_round_keys(a.seeds[ic]);
      }
    }
#pragma unroll 1
    for (int n = 0; n < M; n += 2) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        const u32x4 w = philox4x32_10_rk(c0[j], c1[j], (uint32_t)(n >> 1), 0u, UKEY ? a.rk : rk[UKEY ? 0 : j]);
        double za, zb;
        fast_normal_pair_v2(log_lane, exp_biased, trig_lane, w.x, w.y, w.z, w.w, a.one_hi, a.magic_hi, za, zb);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (n + h < M) {
            const double z = h ? zb : za;
            if (KIND == K_GBM_EM) {  // heston.jl:33-52 under EM; the drift is added at expiry
              sp[j] = fma(sig, z, sp[j]);
              if (ANTI) sm[j] = fma(-sig, z, sm[j]);  // NoiseGrid(t, -W) montecarlo.jl:258
            } else {  // S += S (exp((r - s^2/2) dt + s sqrt(dt) Z) - 1), antithetic: sigma -> -sigma (montecarlo.jl:270-284)
              sp[j] = fma(sp[j], fast_expm1_small(s_exp, fma(sig, z, drift)), sp[j]);
              if (ANTI) sm[j] = fma(sm[j], fast_expm1_small(s_exp, fma(-sig, z, drift)), sm[j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t sub = base + (int64_t)j * kThreads;
      if (sub >= a.n) break;
      const int64_t i = sub + tid;
      const double Sp = KIND == K_GBM_EM ? exp(sp[j] + drift_total) : sp[j];  // final_sample montecarlo.jl:398
      const double Sm = ANTI ? (KIND == K_GBM_EM ? exp(sm[j] + drift_total) : sm[j]) : 0.0;
      if (a.terminal && i < a.n) {
        a.terminal[i] = Sp;
        if (ANTI) a.terminal[a.n + i] = Sm;
      }
      smem[tid] = Sp;
      if (ANTI) smem[kThreads + tid] = Sm;
      __syncthreads();
      const int64_t rem = a.n - sub;
      const int nvalid = rem < kThreads ? (int)rem : kThreads;
      if (k < a.npay) {
        for (int q = g; q < nvalid; q += G) {
          const double s_ = smem[q];
          double pay = fmax(cp * (s_ - strike), 0.0);  // payoffs.jl:154-156
          bool bad = !isfinite(s_);
          if (ANTI) {
            const double t_ = smem[kThreads + q];
            pay = 0.5 * (pay + fmax(cp * (t_ - strike), 0.0));  // reduce_payoffs montecarlo.jl:430-432
            bad = bad || !isfinite(t_);
          }
          acc0 += pay;
          acc1 = fma(pay, pay, acc1);
          if (k == 0 && bad) acc2 += 1.0;
        }
      }
      __syncthreads();
    }
  }
  smem[tid] = acc0;
  smem[kThreads + tid] = acc1;
  smem[2 * kThreads + tid] = acc2;
  __syncthreads();
  if (tid < a.npay) {
    double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * NACC;
    for (int c = 0; c < NACC; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += smem[c * kThreads + (gg << a.kp_log2) + tid];
      out[c] = t;
    }
  }
}

// ---- Float32 fast mode of the headline kernel (config C2, HH_PREC_F32) ----------------------------------------
// FP32 state and normals, FP64 accumulation of the payoff sums. One Philox4x32-10 block feeds TWO steps (32-bit
// uniforms): step n uses words (0, 1) of block n/2 when n is even and words (2, 3) when n is odd; the stream word of
// the counter is 1, so the f32 mode never reuses the f64 mode's numbers. Mapping (restated in oracle/hh_oracle.c):
//   u1 = 2 - float{0x3F800000 | (wa >> 9)}            in [2^-23, 1]      R2 = -2 ln(u1)            (MUFU.LG2)
//   th = 2 pi (float{0x3F800000 | (wb >> 9)} - 1.5)   in [-pi, pi)       (cos, sin)(th)            (MUFU.COS/SIN)
// The state is y = log(S / S0) - r t (starts at 0, so the FP32 rounding of the running sum stays ~1e-8 per step);
// the drift and log S0 are added in FP64 at expiry. One square root per step, as in the FP64 kernel.
struct HestonF32 {
  float neg_half_dt, neg_kdt, ktdt, v0, a11, a12, b21, b22;
};

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool SPLIT>
__device__ __forceinline__ void heston_f32_step(const HestonF32 &h, uint32_t wa, uint32_t wb, uint32_t one_bits, float &yp,
                                                float &vp, float &ym, float &vm, bool anti) {
  const float fa = __uint_as_float((wa >> 9) | one_bits);
  const float fb = __uint_as_float((wb >> 9) | one_bits);
  const float R2 = -1.3862943611198906f * lg2_approx(2.0f - fa);  // -2 ln 2 * log2(u1)
  const float th = fmaf(fb, 6.283185307179586f, -9.42477796076938f);
  const float c = __cosf(th), s = __sinf(th);
  const float cc1 = fmaf(h.a12, s, h.a11 * c);
  const float cc2 = fmaf(h.b22, s, h.b21 * c);
  {
    const float vplus = fmaxf(vp, 0.0f);
    const float K1 = fmaf(h.neg_half_dt, vplus, yp);
    const float K2 = fmaf(h.neg_kdt, vplus, vp + h.ktdt);
    const float sr = sqrt_approx((SPLIT ? fmaxf(K2, 0.0f) : vplus) * R2);
    yp = fmaf(sr, cc1, K1);
    vp = fmaf(sr, cc2, K2);
  }
  if (anti) {
    const float vplus = fmaxf(vm, 0.0f);
    const float K1 = fmaf(h.neg_half_dt, vplus, ym);
    const float K2 = fmaf(h.neg_kdt, vplus, vm + h.ktdt);
    const float sr = sqrt_approx((SPLIT ? fmaxf(K2, 0.0f) : vplus) * R2);
    ym = fmaf(-sr, cc1, K1);
    vm = fmaf(-sr, cc2, K2);
  }
}

template <bool ANTI, bool SPLIT, bool UKEY, int ILP, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) heston_f32_kernel(const EuroArgs a, const HestonF32 h) {
  constexpr int NSIDE = ANTI ? 2 : 1;
  constexpr int NACC = 3;
  constexpr int STAGE = NSIDE * kThreads;
  constexpr int RED = NACC * kThreads;
  __shared__ double smem[STAGE > RED ? STAGE : RED];
  const int tid = threadIdx.x;
  const int KP = 1 << a.kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> a.kp_log2;
  const int G = kThreads >> a.kp_log2;
  double strike = 0.0, cp = 0.0;
  if (k < a.npay) {
    strike = a.payoffs[k].strike;
    cp = a.payoffs[k].cp;
  }
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  const int M = a.n_steps;
  const double x_shift = a.p.x0 + (double)M * a.f.rdt;
  constexpr int64_t kBatch = (int64_t)kThreads * ILP;

  for (int64_t base = (int64_t)blockIdx.x * kBatch; base < a.n; base += (int64_t)gridDim.x * kBatch) {
    float yp[ILP], vp[ILP], ym[ILP], vm[ILP];
    uint32_t c0[ILP], c1[ILP];
    PhiloxRoundKeys rk[UKEY ? 1 : ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t i = base + (int64_t)j * kThreads + tid;
      const int64_t ic = i < a.n ? i : a.n - 1;
      yp[j] = ym[j] = 0.0f;
      vp[j] = vm[j] = h.v0;
      if (UKEY) {
        const uint64_t idx = (uint64_t)(a.path_offset + ic);
        c0[j] = (uint32_t)idx;
        c1[j] = (uint32_t)(idx >> 32);
      } else {
        c0[j] = c1[j] = 0u;
        rk[UKEY ? 0 : j] = philox_round_keys(a.seeds[ic]);
      }
    }
#pragma unroll 1
    for (int n = 0; n < M; n += 2) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        const u32x4 w = philox4x32_10_rk(c0[j], c1[j], (uint32_t)(n >> 1), 1u, UKEY ? a.rk : rk[UKEY ? 0 : j]);
        heston_f32_step<SPLIT>(h, w.x, w.y, a.f32_one, yp[j], vp[j], ym[j], vm[j], ANTI);
        if (n + 1 < M) heston_f32_step<SPLIT>(h, w.z, w.w, a.f32_one, yp[j], vp[j], ym[j], vm[j], ANTI);
      }
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int64_t sub = base + (int64_t)j * kThreads;
      if (sub >= a.n) break;
      const int64_t i = sub + tid;
      const double Sp = exp((double)yp[j] + x_shift);
      const double Sm = ANTI ? exp((double)ym[j] + x_shift) : 0.0;
      if (a.terminal && i < a.n) {
        a.terminal[i] = Sp;
        if (ANTI) a.terminal[a.n + i] = Sm;
      }
      smem[tid] = Sp;
      if (ANTI) smem[kThreads + tid] = Sm;
      __syncthreads();
      const int64_t rem = a.n - sub;
      const int nvalid = rem < kThreads ? (int)rem : kThreads;
      if (k < a.npay) {
        for (int q = g; q < nvalid; q += G) {
          const double sp = smem[q];
          double pay = fmax(cp * (sp - strike), 0.0);  // payoffs.jl:154-156
          bool bad = !isfinite(sp);
          if (ANTI) {
            const double sm = smem[kThreads + q];
            pay = 0.5 * (pay + fmax(cp * (sm - strike), 0.0));  // reduce_payoffs montecarlo.jl:430-432
            bad = bad || !isfinite(sm);
          }
          acc0 += pay;
          acc1 = fma(pay, pay, acc1);
          if (k == 0 && bad) acc2 += 1.0;
        }
      }
      __syncthreads();
    }
  }
  smem[tid] = acc0;
  smem[kThreads + tid] = acc1;
  smem[2 * kThreads + tid] = acc2;
  __syncthreads();
  if (tid < a.npay) {
    double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * NACC;
    for (int c = 0; c < NACC; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += smem[c * kThreads + (gg << a.kp_log2) + tid];
      out[c] = t;
    }
  }
}

// ---- Heston Euler-Maruyama with pathwise tangents, specialised (config C5) ---------------------------------------
// european_kernel<K_HESTON_EM, Dual<8>> pushes a full dual number through every operation: 255 registers, 8 warps per
// SM, and most of the arithmetic multiplies structural zeros. Here the tangent recursion of the step is written out:
//     dK2_p = dv_p (1 - kappa dt 1{v>0}) + A_p - B_p v+          A_p = dt (dkappa_p theta + kappa dtheta_p), B_p = dt dkappa_p
//     ds_p  = 1{arg>0} d(arg)_p / (2 s)                          arg = K2 (split step) or v
//     dx_p' = dx_p - (dt/2) 1{v>0} dv_p + ds_p W1 + s (da11_p z1 + da12_p z2)        (+ dt dr_p, added once at expiry)
//     dv_p' = dK2_p + ds_p W2 + s (db21_p z1 + db22_p z2)         W2 = xi dW2, db2x_p = dxi_p a2x + xi da2x_p
// with the same rules as ForwardDiff (max(x,0) and sqrt pass a tangent iff their argument is positive): 10 FP64
// instructions per direction and step, every per-direction constant a constant-bank operand. Directions that never touch
// the variance (only dS0 and dr non-zero: delta, rho) are "trivial": dx_p is the constant dS0_p/S0 + T dr_p, no per-step
// work. Kernel order of the directions: NF full ones first, then the trivial ones; the host permutes the sums back.
struct HestonTanConsts {
  double dx0[kMaxTan], dv0[kMaxTan], drift[kMaxTan];  // dS0/S0, dV0, T dr
  double A[kMaxTan], B[kMaxTan], da11[kMaxTan], da12[kMaxTan], db21[kMaxTan], db22[kMaxTan];
};

// 1 / sqrt(d) to full double precision: MUFU.RSQ64H seed, two Newton steps
__device__ __forceinline__ double rsqrt_full(double d) {
  double y = rsqrt_seed(d);
  double e = fma(-d * y, y, 1.0);
  y = fma(y * e, fma(e, 0.375, 0.5), y);
  e = fma(-d * y, y, 1.0);
  return fma(y * e, 0.5, y);
}

template <bool SPLIT, int NF>
__device__ __forceinline__ void heston_tangent_step(const HestonFolded &f, const HestonTanConsts &c, double one_m_kdt,
                                                    double &x, double &v, double *dx, double *dv, double z1, double z2,
                                                    double W1, double W2) {
  const bool ind = __double2hiint(v) > 0;  // v > 0 (hi word test: exact unless v is a positive denormal)
  const double vplus = max0_hi(v);
  const double K1 = fma(f.neg_half_dt, vplus, x);
  const double K2 = fma(f.neg_kdt, vplus, v + f.ktdt);
  const double arg = SPLIT ? K2 : v;
  const bool pos = __double2hiint(arg) > 0;
  const double y = rsqrt_full(max_tiny_hi(arg));
  const double s = pos ? arg * y : 0.0;     // sqrt(max(arg, 0))
  const double hs = pos ? 0.5 * y : 0.0;    // d sqrt / d arg
  const double cind = ind ? one_m_kdt : 1.0;
  const double chalf = ind ? f.neg_half_dt : 0.0;
  const double sz1 = s * z1, sz2 = s * z2;
#pragma unroll
  for (int p = 0; p < NF; ++p) {
    const double aP = fma(dv[p], cind, fma(-c.B[p], vplus, c.A[p]));
    const double g = (SPLIT ? aP : (ind ? dv[p] : 0.0)) * hs;
    dx[p] = fma(c.da11[p], sz1, fma(c.da12[p], sz2, fma(g, W1, fma(dv[p], chalf, dx[p]))));
    dv[p] = fma(c.db21[p], sz1, fma(c.db22[p], sz2, fma(g, W2, aP)));
  }
  x = fma(s, W1, K1);
  v = fma(s, W2, K2);
}

template <bool ANTI, bool SPLIT, int NF, int P>
#ifndef HH_TAN_MINB
#define HH_TAN_MINB 2  // resident blocks the register allocation is bounded for (A/B: -DHH_TAN_MINB=1)
#endif
__global__ void __launch_bounds__(kThreads, HH_TAN_MINB) heston_tangent_kernel(const EuroArgs a, const HestonTanConsts c) {
  constexpr int NACC = 3 + 2 * P + kNGamma;
  constexpr int NV = 1 + P;
  constexpr int NSIDE = ANTI ? 2 : 1;
  constexpr int STAGE = NV * NSIDE * kThreads;
  constexpr int RED = NACC * kThreads;
  constexpr int STAGE_DOUBLES = STAGE > RED ? STAGE : RED;
  // dynamic shared memory: [payoff staging / reduction | log table x8 | trig table x8 | exponent table]
  extern __shared__ __align__(16) unsigned char dsm[];
  double *smem = reinterpret_cast<double *>(dsm);
  char *s_log = reinterpret_cast<char *>(dsm) + STAGE_DOUBLES * 8;
  char *s_trig = s_log + kLogRepBytes;
  double *s_e2 = reinterpret_cast<double *>(s_trig + kTrigRepBytes);
  const int tid = threadIdx.x;
  for (int e = tid; e < tables::kLog2Buckets * kRep; e += kThreads)
    reinterpret_cast<double2 *>(s_log)[e] = g_fast_tables2.log_tab[e / kRep];
  for (int e = tid; e < tables::kTrigN * kRep; e += kThreads)
    reinterpret_cast<double2 *>(s_trig)[e] = g_fast_tables2.trig_tab[e / kRep];
  for (int e = tid; e < tables::kExp2N; e += kThreads) s_e2[e] = g_fast_tables2.exp_tab[e];
  __syncthreads();
  const char *log_lane = s_log + (tid & (kRep - 1)) * 16;
  const char *trig_lane = s_trig + (tid & (kRep - 1)) * 16;
  const char *exp_biased = reinterpret_cast<const char *>(s_e2) - tables::kExp2Bias * 8;
  const int KP = 1 << a.kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> a.kp_log2;
  const int G = kThreads >> a.kp_log2;
  double strike = 0.0, cp = 0.0;
  if (k < a.npay) {
    strike = a.payoffs[k].strike;
    cp = a.payoffs[k].cp;
  }
  double acc[NACC];
#pragma unroll
  for (int q = 0; q < NACC; ++q) acc[q] = 0.0;
  const int M = a.n_steps;
  const double drift_total = (double)M * a.f.rdt;
  const double one_m_kdt = 1.0 + a.f.neg_kdt;

  for (int64_t base = (int64_t)blockIdx.x * kThreads; base < a.n; base += (int64_t)gridDim.x * kThreads) {
    const int64_t i = base + tid;
    const int64_t ic = i < a.n ? i : a.n - 1;
    uint64_t key = a.base_seed, idx = (uint64_t)(a.path_offset + ic);
    if (a.seeds) {
      key = a.seeds[ic];
      idx = 0;
    }
    double xp = a.p.x0, vp = a.p.v0, xm = a.p.x0, vm = a.p.v0;
    double dxp[NF > 0 ? NF : 1], dvp[NF > 0 ? NF : 1], dxm[NF > 0 ? NF : 1], dvm[NF > 0 ? NF : 1];
#pragma unroll
    for (int p = 0; p < NF; ++p) {
      dxp[p] = dxm[p] = c.dx0[p];
      dvp[p] = dvm[p] = c.dv0[p];
    }
#pragma unroll 1
    for (int n = 0; n < M; ++n) {
      const u32x4 w = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)n, 0u, (uint32_t)key, (uint32_t)(key >> 32));
      double z1, z2;
      fast_normal_pair_v2(log_lane, exp_biased, trig_lane, w.x, w.y, w.z, w.w, a.one_hi, a.magic_hi, z1, z2);
      const double W1 = fma(a.p.a12, z2, a.p.a11 * z1);
      const double W2 = fma(a.f.b22, z2, a.f.b21 * z1);  // xi dW2
      heston_tangent_step<SPLIT, NF>(a.f, c, one_m_kdt, xp, vp, dxp, dvp, z1, z2, W1, W2);
      if (ANTI) heston_tangent_step<SPLIT, NF>(a.f, c, one_m_kdt, xm, vm, dxm, dvm, -z1, -z2, -W1, -W2);
    }
    const double Sp = exp(xp + drift_total);
    const double Sm = ANTI ? exp(xm + drift_total) : 0.0;
    if (a.terminal && i < a.n) {
      a.terminal[i] = Sp;
      if (ANTI) a.terminal[a.n + i] = Sm;
    }
    smem[tid] = Sp;
    if (ANTI) smem[NV * kThreads + tid] = Sm;
#pragma unroll
    for (int q = 0; q < P; ++q) {  // dS_q = S d(log S)_q
      const double tp_ = q < NF ? dxp[q < NF ? q : 0] + c.drift[q] : c.dx0[q] + c.drift[q];
      smem[(1 + q) * kThreads + tid] = Sp * tp_;
      if (ANTI) {
        const double tm_ = q < NF ? dxm[q < NF ? q : 0] + c.drift[q] : c.dx0[q] + c.drift[q];
        smem[(NV + 1 + q) * kThreads + tid] = Sm * tm_;
      }
    }
    __syncthreads();
    const int64_t rem = a.n - base;
    const int nvalid = rem < kThreads ? (int)rem : kThreads;
    if (k < a.npay) {
      for (int j = g; j < nvalid; j += G) {
        const double sp = smem[j];
        const double ep = cp * (sp - strike);
        double pay = fmax(ep, 0.0);  // payoffs.jl:154-156
        const double ip = ep > 0.0 ? cp : 0.0;
        bool bad = !isfinite(sp);
        double im = 0.0;
        if (ANTI) {
          const double sm = smem[NV * kThreads + j];
          const double em = cp * (sm - strike);
          pay = 0.5 * (pay + fmax(em, 0.0));  // reduce_payoffs montecarlo.jl:430-432
          im = em > 0.0 ? cp : 0.0;
          bad = bad || !isfinite(sm);
        }
        acc[0] += pay;
        acc[1] = fma(pay, pay, acc[1]);
        if (k == 0 && bad) acc[2] += 1.0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
          double dpay = ip * smem[(1 + q) * kThreads + j];  // cp 1{cp(S-K)>0} dS
          if (ANTI) dpay = 0.5 * (dpay + im * smem[(NV + 1 + q) * kThreads + j]);
          acc[3 + q] += dpay;
          acc[3 + P + q] = fma(dpay, dpay, acc[3 + P + q]);
        }
        if (a.g_on) {  // second order in the spot from the same trajectories (GammaArgs in EuroArgs)
          double s2, dd;
          gamma_terms(a, sp, ep, cp, strike, s2, dd);
          if (ANTI) {
            const double sm = smem[NV * kThreads + j];
            double s2m, ddm;
            gamma_terms(a, sm, cp * (sm - strike), cp, strike, s2m, ddm);
            s2 = 0.5 * (s2 + s2m);
            dd = 0.5 * (dd + ddm);
          }
          acc[3 + 2 * P + 0] += s2;
          acc[3 + 2 * P + 1] = fma(s2, s2, acc[3 + 2 * P + 1]);
          acc[3 + 2 * P + 2] += dd;
          acc[3 + 2 * P + 3] = fma(dd, dd, acc[3 + 2 * P + 3]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < NACC; ++q) smem[q * kThreads + tid] = acc[q];
  __syncthreads();
  if (tid < a.npay) {
    double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * NACC;
    for (int q = 0; q < NACC; ++q) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += smem[q * kThreads + (gg << a.kp_log2) + tid];
      out[q] = t;
    }
  }
}

// Payoff sums from terminal spots that another kernel produced (Broadie-Kaya): the same payoff transpose as above.
__global__ void __launch_bounds__(kThreads) terminal_payoff_kernel(const double *__restrict__ terminal, int64_t n,
                                                                   const hh_payoff *__restrict__ payoffs, int npay,
                                                                   int kp_log2, double *partials) {
  constexpr int NACC = 3;
  __shared__ double smem[NACC * kThreads];
  const int tid = threadIdx.x;
  const int KP = 1 << kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> kp_log2;
  const int G = kThreads >> kp_log2;
  double strike = 0.0, cp = 0.0;
  if (k < npay) {
    strike = payoffs[k].strike;
    cp = payoffs[k].cp;
  }
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  for (int64_t base = (int64_t)blockIdx.x * kThreads; base < n; base += (int64_t)gridDim.x * kThreads) {
    const int64_t i = base + tid;
    smem[tid] = i < n ? terminal[i] : 0.0;
    __syncthreads();
    const int64_t rem = n - base;
    const int nvalid = rem < kThreads ? (int)rem : kThreads;
    if (k < npay) {
      for (int q = g; q < nvalid; q += G) {
        const double sp = smem[q];
        const double pay = fmax(cp * (sp - strike), 0.0);  // payoffs.jl:154-156
        acc0 += pay;
        acc1 = fma(pay, pay, acc1);
        if (k == 0 && !isfinite(sp)) acc2 += 1.0;
      }
    }
    __syncthreads();
  }
  smem[tid] = acc0;
  smem[kThreads + tid] = acc1;
  smem[2 * kThreads + tid] = acc2;
  __syncthreads();
  if (tid < npay) {
    double *out = partials + ((size_t)blockIdx.x * npay + tid) * NACC;
    for (int c = 0; c < NACC; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += smem[c * kThreads + (gg << kp_log2) + tid];
      out[c] = t;
    }
  }
}

// Sum the per-block partials in a fixed order: one block per payoff.
__global__ void __launch_bounds__(kThreads) finalize_kernel(const double *partials, int nblocks, int npay, int nacc,
                                                            double *out) {
  __shared__ double scratch[kThreads / 32];
  const int k = blockIdx.x;
  for (int c = 0; c < nacc; ++c) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += kThreads) v += partials[((size_t)b * npay + k) * nacc + c];
    const double t = block_sum<kThreads>(v, scratch);
    if (threadIdx.x == 0) out[(size_t)k * nacc + c] = t;
  }
}

// ---- host side --------------------------------------------------------------------------------------

int validate_model_sim(hh_ctx *ctx, const hh_model *m, const hh_sim *s) {
  if (!m || !s) return ctx->fail(HH_ERR_ARG, "model/sim is NULL");
  if (s->n_paths <= 0) return ctx->fail(HH_ERR_ARG, "n_paths must be positive (got %lld)", (long long)s->n_paths);
  if (s->scheme != HH_SCHEME_EXACT_TERMINAL && s->n_steps <= 0)
    return ctx->fail(HH_ERR_ARG, "n_steps must be positive (got %d)", s->n_steps);
  if (s->vr != HH_VR_NONE && s->vr != HH_VR_ANTITHETIC && s->vr != HH_VR_QUASI_RANDOM)
    return ctx->fail(HH_ERR_ARG, "unknown variance reduction %d", s->vr);
  if (s->vr == HH_VR_QUASI_RANDOM && !(m->kind == HH_MODEL_GBM && s->scheme == HH_SCHEME_EXACT_TERMINAL &&
                                       s->rng_mode == HH_RNG_PHILOX && s->precision == HH_PREC_F64 && !s->seeds))
    return ctx->fail(HH_ERR_UNSUPPORTED, "HH_VR_QUASI_RANDOM is defined for the one-draw exact sampler (LognormalDynamics + "
                                         "BlackScholesExact, European), in-kernel stream keyed by base_seed, f64");
  if (s->rng_mode != HH_RNG_NORMALS && s->rng_mode != HH_RNG_PHILOX && s->rng_mode != HH_RNG_PHILOX_64)
    return ctx->fail(HH_ERR_ARG, "unknown rng_mode %d", s->rng_mode);
  if (s->rng_mode == HH_RNG_NORMALS && !s->normals)
    return ctx->fail(HH_ERR_ARG, "rng_mode = HH_RNG_NORMALS needs a normals buffer");
  if (s->rng_mode == HH_RNG_PHILOX_64 && s->precision == HH_PREC_F64 && m->kind == HH_MODEL_GBM &&
      s->scheme == HH_SCHEME_EXACT_STEPS) {
    // the LSM path generator's form of the stream (hh_lsm.cu); every other entry point rejects it below
  } else if (s->rng_mode == HH_RNG_PHILOX_64 &&
             !(m->kind == HH_MODEL_HESTON && s->scheme == HH_SCHEME_EM && s->precision == HH_PREC_F64))
    return ctx->fail(HH_ERR_UNSUPPORTED, "HH_RNG_PHILOX_64 is defined for HestonDynamics + EulerMaruyama in f64 (European "
                                         "pricing) and for the exact GBM generator of hh_lsm_american");
  // the lengths behind the caller's pointers (the copies below read exactly this many elements)
  if (s->rng_mode == HH_RNG_NORMALS) {
    const uint64_t ncomp = m->kind == HH_MODEL_HESTON ? 2 : 1;
    const uint64_t nst = s->scheme == HH_SCHEME_EXACT_TERMINAL ? 1 : (uint64_t)(s->n_steps > 0 ? s->n_steps : 0);
    const uint64_t need = (uint64_t)s->n_paths * nst * ncomp;
    if (s->normals_len < need)
      return ctx->fail(HH_ERR_ARG, "normals buffer too short: %llu elements for n_paths x n_steps x components = %llu",
                       (unsigned long long)s->normals_len, (unsigned long long)need);
  } else if (s->seeds && s->seeds_len < (uint64_t)s->n_paths) {  // montecarlo.jl:65-66
    return ctx->fail(HH_ERR_ARG, "Number of seeds (%llu) must be >= number of trajectories (%lld).",
                     (unsigned long long)s->seeds_len, (long long)s->n_paths);
  }
  if (m->kind == HH_MODEL_GBM) {
    if (s->scheme != HH_SCHEME_EM && s->scheme != HH_SCHEME_EXACT_TERMINAL && s->scheme != HH_SCHEME_EXACT_STEPS)
      return ctx->fail(HH_ERR_ARG, "scheme %d is not defined for LognormalDynamics", s->scheme);
  } else if (m->kind == HH_MODEL_HESTON) {
    if (s->scheme != HH_SCHEME_EM && s->scheme != HH_SCHEME_HESTON_BK)
      return ctx->fail(HH_ERR_ARG, "scheme %d is not defined for HestonDynamics", s->scheme);
    // Q5: Antithetic + HestonBroadieKaya calls mean(::LogHestonDistribution), which does not exist (montecarlo.jl:387)
    if (s->scheme == HH_SCHEME_HESTON_BK && s->vr == HH_VR_ANTITHETIC)
      return ctx->fail(HH_ERR_UNSUPPORTED, "Antithetic + HestonBroadieKaya is a MethodError in the reference (Q5)");
  } else {
    return ctx->fail(HH_ERR_ARG, "unknown model kind %d", m->kind);
  }
  if (!(m->S0 > 0.0) || !(m->S0 < 1e300)) return ctx->fail(HH_ERR_ARG, "spot must be positive");
  if (!(m->T > 0.0) || !(m->T < 1e300)) return ctx->fail(HH_ERR_ARG, "time to expiry must be positive");
  // A NaN parameter does not always surface as a NaN price: the full-truncation max(v, 0) and the clamps of the table-driven
  // functions swallow it (a NaN kappa priced a call at 2.62 with n_nonfinite = 0). Non-finite parameters are an argument error.
  if (!(fabs(m->r) < 1e300)) return ctx->fail(HH_ERR_ARG, "the rate must be finite");
  if (m->kind == HH_MODEL_GBM) {
    if (!(fabs(m->sigma) < 1e300)) return ctx->fail(HH_ERR_ARG, "the volatility must be finite");
  } else {
    const double hp[] = {m->V0, m->kappa, m->theta, m->xi, m->rho, m->m11, m->m12, m->m21, m->m22};
    const char *hn[] = {"V0", "kappa", "theta", "xi", "rho", "m11", "m12", "m21", "m22"};
    for (int k = 0; k < 9; ++k)
      if (!(fabs(hp[k]) < 1e300)) return ctx->fail(HH_ERR_ARG, "Heston parameter %s must be finite", hn[k]);
  }
  return HH_OK;
}

template <int KIND, class T, bool ANTI, int PARITY, int SPLIT>
static cudaError_t launch_one(const EuroArgs &a, const TangentPack *tp, int sm_count, cudaStream_t st, int *nblocks,
                              bool query_only) {
  auto kern = european_kernel<KIND, T, ANTI, PARITY, SPLIT>;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int64_t batches = (a.n + kThreads - 1) / kThreads;
  int64_t grid = (int64_t)sm_count * occ;  // persistent: one wave, grid-stride over batches
  if (grid > batches) grid = batches;
  *nblocks = (int)grid;
  if (query_only) return cudaSuccess;
  kern<<<(unsigned)grid, kThreads, 0, st>>>(a, tp);
  return cudaGetLastError();
}

// plain pricing: PARITY and SPLIT are compile-time
template <int KIND, bool ANTI>
static cudaError_t launch_plain(const EuroArgs &a, int sm_count, cudaStream_t st, int *nb, bool q) {
  if (KIND == K_HESTON_EM) {
    if (a.parity) {
      if (a.split) return launch_one<KIND, double, ANTI, 1, 1>(a, nullptr, sm_count, st, nb, q);
      return launch_one<KIND, double, ANTI, 1, 0>(a, nullptr, sm_count, st, nb, q);
    }
    if (a.split) return launch_one<KIND, double, ANTI, 0, 1>(a, nullptr, sm_count, st, nb, q);
    return launch_one<KIND, double, ANTI, 0, 0>(a, nullptr, sm_count, st, nb, q);
  }
  if (a.parity) return launch_one<KIND, double, ANTI, 1, 0>(a, nullptr, sm_count, st, nb, q);
  return launch_one<KIND, double, ANTI, 0, 0>(a, nullptr, sm_count, st, nb, q);
}

template <int KIND, int P>
static cudaError_t launch_tan(const EuroArgs &a, const TangentPack *tp, bool anti, int sm_count, cudaStream_t st,
                              int *nb, bool q) {
  if (anti) return launch_one<KIND, Dual<P>, true, 2, 2>(a, tp, sm_count, st, nb, q);
  return launch_one<KIND, Dual<P>, false, 2, 2>(a, tp, sm_count, st, nb, q);
}

template <int KIND>
static cudaError_t launch_kind(const EuroArgs &a, const TangentPack *tp, int P, bool anti, int sm_count,
                               cudaStream_t st, int *nb, bool q) {
  switch (P) {
    case 0: return anti ? launch_plain<KIND, true>(a, sm_count, st, nb, q) : launch_plain<KIND, false>(a, sm_count, st, nb, q);
    case 1: return launch_tan<KIND, 1>(a, tp, anti, sm_count, st, nb, q);
    case 2: return launch_tan<KIND, 2>(a, tp, anti, sm_count, st, nb, q);
    case 4: return launch_tan<KIND, 4>(a, tp, anti, sm_count, st, nb, q);
    default: return launch_tan<KIND, 8>(a, tp, anti, sm_count, st, nb, q);
  }
}

#ifdef HH_TUNING
template <bool ANTI, bool SPLIT, bool UKEY, int ILP, int MINB>
static cudaError_t launch_fast_one(const EuroArgs &a, int sm_count, cudaStream_t st, int *nblocks, bool query_only) {
  auto kern = heston_fast_kernel<ANTI, SPLIT, UKEY, ILP, MINB>;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int64_t batch = (int64_t)kThreads * ILP;
  const int64_t batches = (a.n + batch - 1) / batch;
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > batches) grid = batches;
  *nblocks = (int)grid;
  if (query_only) return cudaSuccess;
  kern<<<(unsigned)grid, kThreads, 0, st>>>(a);
  return cudaGetLastError();
}
#endif

template <bool ANTI, bool SPLIT, bool UKEY, int ILP, int THREADS, int MINB, bool R64 = false, int ABL = 0>
static cudaError_t launch_fast2_one(const EuroArgs &a, int sm_count, cudaStream_t st, int *nblocks, bool query_only) {
  auto kern = heston_fast2_kernel<ANTI, SPLIT, UKEY, ILP, THREADS, MINB, R64, ABL>;
  static PerDeviceOnce opted;  // per instantiation and device
  if (cudaError_t e0 = smem_opt_in(opted, kern, fast2_smem_bytes(THREADS)); e0 != cudaSuccess) return e0;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, fast2_smem_bytes(THREADS));
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int64_t batch = (int64_t)THREADS * ILP;
  const int64_t batches = (a.n + batch - 1) / batch;
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > batches) grid = batches;
  *nblocks = (int)grid;
  if (query_only) return cudaSuccess;
  kern<<<(unsigned)grid, THREADS, fast2_smem_bytes(THREADS), st>>>(a);
  return cudaGetLastError();
}

// Block shape of the headline kernel, measured on B200 at 1e8 x 252: 1024 threads x ILP 1 99.4 ms, 256 x ILP 2 100.9 ms,
// 512 x ILP 1 102.5 ms (all within 3 %: the kernel is dispatch-bound, occupancy hardly matters); small jobs keep the
// 256-thread blocks so that every SM gets work. The other shapes that were tried (HH_HESTON_VARIANT) are compiled only
// with -DHH_TUNING.
template <bool ANTI, bool SPLIT, bool UKEY, bool R64>
static cudaError_t launch_fast_v(const EuroArgs &a, int variant, int sm_count, cudaStream_t st, int *nb, bool q) {
#ifdef HH_TUNING
  if (!R64) {
    switch (variant) {
      // measured on B200, 1e7 x 252 (tools/time_heston.py): ILP 2 / 2 blocks per SM 12.75 ms, ILP 1 / 4 blocks 13.42 ms,
      // ILP 1 / 6 blocks 13.83 ms — the kernel is issue-bound, occupancy hardly matters
      case 1: return launch_fast_one<ANTI, SPLIT, UKEY, 1, 4>(a, sm_count, st, nb, q);
      case 2: return launch_fast_one<ANTI, SPLIT, UKEY, 1, 6>(a, sm_count, st, nb, q);
      case 3: return launch_fast_one<ANTI, SPLIT, UKEY, 2, 3>(a, sm_count, st, nb, q);
      case 4: return launch_fast_one<ANTI, SPLIT, UKEY, 2, 2>(a, sm_count, st, nb, q);  // v1 (tables not replicated)
      case 5: return launch_fast2_one<ANTI, SPLIT, UKEY, 1, 256, 2>(a, sm_count, st, nb, q);
      case 6: return launch_fast2_one<ANTI, SPLIT, UKEY, 1, 512, 2>(a, sm_count, st, nb, q);   // 32 warps per SM, <= 64 registers
      case 8: return launch_fast2_one<ANTI, SPLIT, UKEY, 2, 512, 1>(a, sm_count, st, nb, q);   // 16 warps, one block
      case 9: return launch_fast2_one<ANTI, SPLIT, UKEY, 1, 768, 1>(a, sm_count, st, nb, q);   // 24 warps, <= 80 registers
      default: break;
    }
  }
#endif
  (void)variant;
  if (a.n >= (int64_t)sm_count * 1024 * 4) return launch_fast2_one<ANTI, SPLIT, UKEY, 1, 1024, 1, R64>(a, sm_count, st, nb, q);
  return launch_fast2_one<ANTI, SPLIT, UKEY, 2, 256, 2, R64>(a, sm_count, st, nb, q);
}

static cudaError_t launch_fast(const EuroArgs &a, bool anti, int sm_count, cudaStream_t st, int *nb, bool q) {
  static const int variant = getenv("HH_HESTON_VARIANT") ? atoi(getenv("HH_HESTON_VARIANT")) : 0;
  const bool ukey = a.seeds == nullptr;
#define HH_FAST(A, S, U)                                                                    \
  do {                                                                                      \
    if (a.rng64) return launch_fast_v<A, S, U, true>(a, variant, sm_count, st, nb, q);      \
    return launch_fast_v<A, S, U, false>(a, variant, sm_count, st, nb, q);                  \
  } while (0)
  if (anti) {
    if (a.split) { if (ukey) HH_FAST(true, true, true); else HH_FAST(true, true, false); }
    else { if (ukey) HH_FAST(true, false, true); else HH_FAST(true, false, false); }
  } else {
    if (a.split) { if (ukey) HH_FAST(false, true, true); else HH_FAST(false, true, false); }
    else { if (ukey) HH_FAST(false, false, true); else HH_FAST(false, false, false); }
  }
#undef HH_FAST
}

template <bool ANTI, bool SPLIT, bool UKEY, int ILP, int MINB>
static cudaError_t launch_f32_one(const EuroArgs &a, const HestonF32 &h, int sm_count, cudaStream_t st, int *nblocks,
                                  bool query_only) {
  auto kern = heston_f32_kernel<ANTI, SPLIT, UKEY, ILP, MINB>;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int64_t batch = (int64_t)kThreads * ILP;
  const int64_t batches = (a.n + batch - 1) / batch;
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > batches) grid = batches;
  *nblocks = (int)grid;
  if (query_only) return cudaSuccess;
  kern<<<(unsigned)grid, kThreads, 0, st>>>(a, h);
  return cudaGetLastError();
}

static cudaError_t launch_f32(const EuroArgs &a, bool anti, int sm_count, cudaStream_t st, int *nb, bool q) {
  HestonF32 h;
  h.neg_half_dt = (float)a.f.neg_half_dt;
  h.neg_kdt = (float)a.f.neg_kdt;
  h.ktdt = (float)a.f.ktdt;
  h.v0 = (float)a.p.v0;
  h.a11 = (float)a.p.a11;
  h.a12 = (float)a.p.a12;
  h.b21 = (float)a.f.b21;
  h.b22 = (float)a.f.b22;
  const bool ukey = a.seeds == nullptr;
  static const int variant = getenv("HH_F32_VARIANT") ? atoi(getenv("HH_F32_VARIANT")) : 0;
#define HH_F32(A, S, U)                                                                   \
  switch (variant) {                                                                      \
    /* measured on B200, 5e7 x 252: ILP 4 / 3 blocks 23.3 ms, ILP 2 / 4 blocks 25.3 ms, ILP 1 / 8 blocks 25.1 ms */ \
    case 1: return launch_f32_one<A, S, U, 2, 4>(a, h, sm_count, st, nb, q);             \
    case 2: return launch_f32_one<A, S, U, 1, 8>(a, h, sm_count, st, nb, q);             \
    default: return launch_f32_one<A, S, U, 4, 3>(a, h, sm_count, st, nb, q);            \
  }
  if (anti) {
    if (a.split) { if (ukey) { HH_F32(true, true, true) } else { HH_F32(true, true, false) } }
    else { if (ukey) { HH_F32(true, false, true) } else { HH_F32(true, false, false) } }
  } else {
    if (a.split) { if (ukey) { HH_F32(false, true, true) } else { HH_F32(false, true, false) } }
    else { if (ukey) { HH_F32(false, false, true) } else { HH_F32(false, false, false) } }
  }
#undef HH_F32
}

template <bool ANTI, bool SPLIT, int NF, int P>
static cudaError_t launch_tangent_one(const EuroArgs &a, const HestonTanConsts &c, int sm_count, cudaStream_t st, int *nblocks,
                                      bool query_only) {
  auto kern = heston_tangent_kernel<ANTI, SPLIT, NF, P>;
  constexpr int NACC = 3 + 2 * P + kNGamma, STAGE = (1 + P) * (ANTI ? 2 : 1) * kThreads, RED = NACC * kThreads;
  constexpr int smem = (STAGE > RED ? STAGE : RED) * 8 + kLogRepBytes + kTrigRepBytes + kExp2Bytes;
  static PerDeviceOnce opted;  // per instantiation and device
  if (cudaError_t e0 = smem_opt_in(opted, kern, smem); e0 != cudaSuccess) return e0;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int64_t batches = (a.n + kThreads - 1) / kThreads;
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > batches) grid = batches;
  *nblocks = (int)grid;
  if (query_only) return cudaSuccess;
  kern<<<(unsigned)grid, kThreads, smem, st>>>(a, c);
  return cudaGetLastError();
}

// nf: full directions (rounded up to an instantiated count), P: padded direction count (1, 2, 4, 8)
template <bool ANTI, bool SPLIT>
static cudaError_t launch_tangent_as(const EuroArgs &a, const HestonTanConsts &c, int nf, int P, int sm_count, cudaStream_t st,
                                     int *nb, bool q) {
#define HH_TAN(NF_, P_) return launch_tangent_one<ANTI, SPLIT, NF_, P_>(a, c, sm_count, st, nb, q)
  switch (P) {
    case 1: if (nf == 0) HH_TAN(0, 1); HH_TAN(1, 1);
    case 2: if (nf == 0) HH_TAN(0, 2); if (nf == 1) HH_TAN(1, 2); HH_TAN(2, 2);
    case 4: if (nf == 0) HH_TAN(0, 4); if (nf <= 2) HH_TAN(2, 4); if (nf == 3) HH_TAN(3, 4); HH_TAN(4, 4);
    default:
      if (nf <= 2) HH_TAN(2, 8);
      if (nf <= 4) HH_TAN(4, 8);
      if (nf == 5) HH_TAN(5, 8);  // C5: V0, kappa, theta, xi, rho (+ the trivial S0, r)
      if (nf == 6) HH_TAN(6, 8);
      if (nf == 7) HH_TAN(7, 8);
      HH_TAN(8, 8);
  }
#undef HH_TAN
}

static cudaError_t launch_tangent(const EuroArgs &a, const HestonTanConsts &c, int nf, int P, bool anti, int sm_count,
                                  cudaStream_t st, int *nb, bool q) {
  if (anti) return a.split ? launch_tangent_as<true, true>(a, c, nf, P, sm_count, st, nb, q)
                           : launch_tangent_as<true, false>(a, c, nf, P, sm_count, st, nb, q);
  return a.split ? launch_tangent_as<false, true>(a, c, nf, P, sm_count, st, nb, q)
                 : launch_tangent_as<false, false>(a, c, nf, P, sm_count, st, nb, q);
}

template <int KIND, bool ANTI, bool UKEY, int ILP, int THREADS>
static cudaError_t launch_gbm_fast_cfg(const EuroArgs &a, int sm_count, cudaStream_t st, int *nblocks, bool query_only) {
  constexpr int kThreads = THREADS;
  auto kern = gbm_fast_kernel<KIND, ANTI, UKEY, ILP, THREADS>;
  constexpr int smem = gbm_fast_smem(ANTI, THREADS, KIND == K_GBM_STEPS);
  static PerDeviceOnce opted;  // per instantiation and device
  if (cudaError_t e0 = smem_opt_in(opted, kern, smem); e0 != cudaSuccess) return e0;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int64_t batch = (int64_t)kThreads * ILP;
  const int64_t batches = (a.n + batch - 1) / batch;
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > batches) grid = batches;
  *nblocks = (int)grid;
  if (query_only) return cudaSuccess;
  kern<<<(unsigned)grid, kThreads, smem, st>>>(a);
  return cudaGetLastError();
}

// measured at 2e7 x 252 (EM / exact steps, 1e11 path-steps/s): 256 x ILP 2: 4.36 / 2.90, 512 x ILP 2: 4.80 / 3.28,
// 1024 x ILP 1: 4.67 / 3.29 — large jobs take 512-thread blocks (more warps share one copy of the tables)
template <int KIND, bool ANTI, bool UKEY>
static cudaError_t launch_gbm_fast_one(const EuroArgs &a, int sm_count, cudaStream_t st, int *nblocks, bool query_only) {
  static const int variant = getenv("HH_GBM_VARIANT") ? atoi(getenv("HH_GBM_VARIANT")) : 0;
  if (variant == 1) return launch_gbm_fast_cfg<KIND, ANTI, UKEY, 2, 256>(a, sm_count, st, nblocks, query_only);
  if (variant == 2) return launch_gbm_fast_cfg<KIND, ANTI, UKEY, 2, 512>(a, sm_count, st, nblocks, query_only);
  if (variant == 3) return launch_gbm_fast_cfg<KIND, ANTI, UKEY, 1, 1024>(a, sm_count, st, nblocks, query_only);
  if (a.n >= (int64_t)sm_count * 1024 * 4) return launch_gbm_fast_cfg<KIND, ANTI, UKEY, 2, 512>(a, sm_count, st, nblocks, query_only);
  return launch_gbm_fast_cfg<KIND, ANTI, UKEY, 2, 256>(a, sm_count, st, nblocks, query_only);
}

template <int KIND>
static cudaError_t launch_gbm_fast(const EuroArgs &a, bool anti, int sm_count, cudaStream_t st, int *nb, bool q) {
  const bool ukey = a.seeds == nullptr;
  if (anti) return ukey ? launch_gbm_fast_one<KIND, true, true>(a, sm_count, st, nb, q)
                        : launch_gbm_fast_one<KIND, true, false>(a, sm_count, st, nb, q);
  return ukey ? launch_gbm_fast_one<KIND, false, true>(a, sm_count, st, nb, q)
              : launch_gbm_fast_one<KIND, false, false>(a, sm_count, st, nb, q);
}

static cudaError_t launch_any(int kind, const EuroArgs &a, const TangentPack *tp, int P, bool anti, int sm_count,
                              cudaStream_t st, int *nb, bool q) {
  if (a.f32) return launch_f32(a, anti, sm_count, st, nb, q);
  if (kind == K_HESTON_EM && P == 0 && !a.parity) return launch_fast(a, anti, sm_count, st, nb, q);
  static const bool gbm_generic = getenv("HH_GBM_GENERIC") != nullptr;
  if (P == 0 && !a.parity && !gbm_generic) {
    if (kind == K_GBM_EM) return launch_gbm_fast<K_GBM_EM>(a, anti, sm_count, st, nb, q);
    if (kind == K_GBM_STEPS) return launch_gbm_fast<K_GBM_STEPS>(a, anti, sm_count, st, nb, q);
  }
  switch (kind) {
    case K_GBM_EM: return launch_kind<K_GBM_EM>(a, tp, P, anti, sm_count, st, nb, q);
    case K_GBM_TERMINAL: return launch_kind<K_GBM_TERMINAL>(a, tp, P, anti, sm_count, st, nb, q);
    case K_GBM_STEPS: return launch_kind<K_GBM_STEPS>(a, tp, P, anti, sm_count, st, nb, q);
    default: return launch_kind<K_HESTON_EM>(a, tp, P, anti, sm_count, st, nb, q);
  }
}

// Scalars derived on the host from the model, and their tangents (chain rule through the same formulas).
static int build_args(hh_ctx *ctx, const hh_model *m, const hh_sim *s, const hh_tangent *tg, int ntan, EuroArgs &a,
                      TangentPack &tp, int &kind) {
  memset(&a, 0, sizeof a);
  memset(&tp, 0, sizeof tp);
  const int nsteps = s->scheme == HH_SCHEME_EXACT_TERMINAL ? 1 : s->n_steps;
  a.n = s->n_paths;
  a.path_offset = s->path_offset;
  a.base_seed = s->base_seed;
  a.n_steps = nsteps;
  a.split = (m->flags & HH_FLAG_SPLIT_STEP) ? 1 : 0;
  a.parity = s->rng_mode == HH_RNG_NORMALS;
  a.f32 = s->precision == HH_PREC_F32;
  a.rk = philox_round_keys(s->base_seed);
  a.one_hi = 0x3FF00000u;
  a.magic_hi = 0x43300000u;
  a.f32_one = 0x3F800000u;
  a.lo_fill = 0x00080000u;
  a.rng64 = s->rng_mode == HH_RNG_PHILOX_64;
  a.qmc = s->vr == HH_VR_QUASI_RANDOM;
  a.qmc_shift = qmc_shift_of(s->base_seed);
  PathParams<double> &p = a.p;
  const double dt = m->T / nsteps;  // montecarlo.jl:349
  const double sqdt = sqrt(dt);
  p.dt = dt;
  p.sqdt = sqdt;
  p.x0 = log(m->S0);
  p.S0 = m->S0;
  p.r = m->r;
  for (int q = 0; q < ntan; ++q) {
    tp.x0[q] = tg[q].dS0 / m->S0;
    tp.S0[q] = tg[q].dS0;
    tp.r[q] = tg[q].dr;
  }
  if (m->kind == HH_MODEL_GBM) {
    p.sigma = m->sigma;
    const double drift = m->r - 0.5 * (m->sigma * m->sigma);
    for (int q = 0; q < ntan; ++q) tp.sigma[q] = tg[q].dsigma;
    if (s->scheme == HH_SCHEME_EXACT_TERMINAL) {
      kind = K_GBM_TERMINAL;
      const double alpha = m->T;
      const double c = (m->flags & HH_FLAG_Q1_SQRT_MEAN) ? sqrt(alpha) : alpha;  // Q1, montecarlo.jl:302
      p.mu = log(m->S0) + (m->r - m->sigma * m->sigma / 2) * c;
      p.sd = m->sigma * sqrt(alpha);
      for (int q = 0; q < ntan; ++q) {
        tp.mu[q] = tg[q].dS0 / m->S0 + (tg[q].dr - m->sigma * tg[q].dsigma) * c;
        tp.sd[q] = tg[q].dsigma * sqrt(alpha);
      }
    } else {
      kind = s->scheme == HH_SCHEME_EM ? K_GBM_EM : K_GBM_STEPS;
      p.dt_drift = s->scheme == HH_SCHEME_EM ? dt * drift : drift * dt;
      p.sig_sqdt = m->sigma * sqdt;
      for (int q = 0; q < ntan; ++q) {
        tp.dt_drift[q] = (tg[q].dr - m->sigma * tg[q].dsigma) * dt;
        tp.sig_sqdt[q] = tg[q].dsigma * sqdt;
      }
    }
  } else {
    kind = K_HESTON_EM;
    p.v0 = m->V0;
    p.kappa = m->kappa;
    p.theta = m->theta;
    p.xi = m->xi;
    p.a11 = sqdt * m->m11;
    p.a12 = sqdt * m->m12;
    p.a21 = sqdt * m->m21;
    p.a22 = sqdt * m->m22;
    a.f.rdt = m->r * dt;
    a.f.neg_half_dt = -0.5 * dt;
    a.f.neg_kdt = -(m->kappa * dt);
    a.f.ktdt = m->kappa * m->theta * dt;
    a.f.b21 = m->xi * p.a21;
    a.f.b22 = m->xi * p.a22;
    for (int q = 0; q < ntan; ++q) {
      tp.v0[q] = tg[q].dV0;
      tp.kappa[q] = tg[q].dkappa;
      tp.theta[q] = tg[q].dtheta;
      tp.xi[q] = tg[q].dxi;
      tp.a11[q] = sqdt * tg[q].dm11;
      tp.a12[q] = sqdt * tg[q].dm12;
      tp.a21[q] = sqdt * tg[q].dm21;
      tp.a22[q] = sqdt * tg[q].dm22;
    }
  }
  (void)ctx;
  return HH_OK;
}

static int upload_inputs(hh_ctx *ctx, const hh_model *m, const hh_sim *s, const hh_payoff *payoffs, int npay,
                         EuroArgs &a) {
  cudaStream_t st = ctx->stream;
  HH_CUDA(ctx, upload_fast_tables(ctx->device, st));
  HH_CUDA(ctx, upload_fast_tables2(ctx->device, st));
  const int64_t N = s->n_paths;
  const int ncomp = m->kind == HH_MODEL_HESTON ? 2 : 1;
  a.npay = npay;
  a.kp_log2 = 0;
  while ((1 << a.kp_log2) < npay) a.kp_log2++;
  HH_CUDA(ctx, ctx->d_payoffs.ensure(sizeof(hh_payoff) * (size_t)npay));
  HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_payoffs.ptr, payoffs, sizeof(hh_payoff) * (size_t)npay, cudaMemcpyHostToDevice, st));
  a.payoffs = ctx->d_payoffs.as<hh_payoff>();
  if (a.parity) {
    const size_t bytes = sizeof(double) * (size_t)N * (size_t)a.n_steps * (size_t)ncomp;
    HH_CUDA(ctx, ctx->d_normals.ensure(bytes));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_normals.ptr, s->normals, bytes, cudaMemcpyHostToDevice, st));
    a.normals = ctx->d_normals.as<double>();
  } else if (s->seeds) {
    const size_t bytes = sizeof(uint64_t) * (size_t)N;
    HH_CUDA(ctx, ctx->d_seeds.ensure(bytes));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_seeds.ptr, s->seeds, bytes, cudaMemcpyHostToDevice, st));
    a.seeds = ctx->d_seeds.as<uint64_t>();
  }
  return HH_OK;
}

int european_launch(hh_ctx *ctx, const hh_model *m, const hh_sim *s, const hh_payoff *payoffs, int npay,
                    int want_terminal) {
  int rc = validate_model_sim(ctx, m, s);
  if (rc) return rc;
  if (npay < 1 || npay > kThreads || !payoffs)
    return ctx->fail(HH_ERR_ARG, "npayoffs must be in [1, %d] (got %d)", kThreads, npay);
  if (s->scheme == HH_SCHEME_HESTON_BK) return ctx->fail(HH_ERR_UNSUPPORTED, "use the Broadie-Kaya driver");
  if (s->rng_mode == HH_RNG_PHILOX_64 && m->kind != HH_MODEL_HESTON)
    return ctx->fail(HH_ERR_UNSUPPORTED, "HH_RNG_PHILOX_64 prices HestonDynamics + EulerMaruyama; under LognormalDynamics it is the "
                                         "LSM path generator's stream only");
  if (s->precision == HH_PREC_F32 && (m->kind != HH_MODEL_HESTON || s->scheme != HH_SCHEME_EM || s->rng_mode != HH_RNG_PHILOX))
    return ctx->fail(HH_ERR_UNSUPPORTED, "the f32 fast mode covers Heston Euler-Maruyama with the in-kernel RNG (config C2)");
  if (s->precision != HH_PREC_F64 && s->precision != HH_PREC_F32) return ctx->fail(HH_ERR_ARG, "unknown precision %d", s->precision);

  const int64_t N = s->n_paths;
  const bool anti = s->vr == HH_VR_ANTITHETIC;
  EuroArgs a;
  TangentPack tp;
  int kind = 0;
  build_args(ctx, m, s, nullptr, 0, a, tp, kind);

  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  rc = upload_inputs(ctx, m, s, payoffs, npay, a);
  if (rc) return rc;
  if (want_terminal) {
    HH_CUDA(ctx, ctx->d_terminal.ensure(sizeof(double) * (size_t)N * (anti ? 2 : 1)));
    a.terminal = ctx->d_terminal.as<double>();
  }
  constexpr int NACC = nacc_of<double>();
  int nblocks = 0;
  HH_CUDA(ctx, launch_any(kind, a, nullptr, 0, anti, ctx->sm_count, st, &nblocks, true));
  HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)nblocks * npay * NACC));
  HH_CUDA(ctx, ctx->d_final.ensure(sizeof(double) * (size_t)npay * NACC));
  a.partials = ctx->d_partials.as<double>();

  // MonteCarloSolution.ensemble of a large job: launch the trajectories in segments, so that collect() can copy the
  // terminal values of segment i to the host while segment i+1 is being simulated (the kernels are grid-stride over
  // a.n with the global trajectory index in the Philox counter, so a segment is just a shard)
  static const int64_t seg_min = getenv("HH_SEGMENT_MIN") ? atoll(getenv("HH_SEGMENT_MIN")) : ((int64_t)4 << 20);
  int nseg = 1;
  if (want_terminal && !anti && !a.parity && !s->seeds && N >= 2 * seg_min) {
    nseg = (int)(N / seg_min);
    if (nseg > 8) nseg = 8;
  }
  ctx->pend.nseg = 0;
  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  if (nseg == 1) {
    HH_CUDA(ctx, launch_any(kind, a, nullptr, 0, anti, ctx->sm_count, st, &nblocks, false));
  } else {
    // every segment runs a full grid: its own slice of the partial sums
    int64_t per = ((N / nseg) + 1023) & ~(int64_t)1023;
    int total_blocks = 0, nb = 0;
    std::vector<EuroArgs> segs;
    for (int64_t lo = 0; lo < N; lo += per) {
      EuroArgs b = a;
      b.n = N - lo < per ? N - lo : per;
      b.path_offset = a.path_offset + lo;
      b.terminal = a.terminal + lo;
      HH_CUDA(ctx, launch_any(kind, b, nullptr, 0, anti, ctx->sm_count, st, &nb, true));
      total_blocks += nb;
      segs.push_back(b);
    }
    HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)total_blocks * npay * NACC));
    a.partials = ctx->d_partials.as<double>();
    int done = 0, i = 0;
    int64_t end = 0;
    for (EuroArgs &b : segs) {
      b.partials = a.partials + (size_t)done * npay * NACC;
      HH_CUDA(ctx, launch_any(kind, b, nullptr, 0, anti, ctx->sm_count, st, &nb, false));
      done += nb;
      end += b.n;
      if (!ctx->ev_seg[i]) HH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_seg[i], cudaEventDisableTiming));
      HH_CUDA(ctx, cudaEventRecord(ctx->ev_seg[i], st));
      ctx->pend.seg_end[i] = end;
      ++i;
    }
    ctx->pend.nseg = i;
    nblocks = total_blocks;
  }
  finalize_kernel<<<npay, kThreads, 0, st>>>(a.partials, nblocks, npay, NACC, ctx->d_final.as<double>());
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));

  ctx->pend.active = true;
  ctx->pend.npay = npay;
  ctx->pend.nblocks = nblocks;
  ctx->pend.n = N;
  ctx->pend.anti = anti;
  ctx->pend.want_terminal = want_terminal != 0;
  ctx->pend.bk = false;
  return HH_OK;
}

// Enqueue the payoff reduction over terminal spots already on the device; ev0 was recorded by the caller before its
// own kernels. Leaves the context in the same pending state as european_launch.
int terminal_payoffs_launch(hh_ctx *ctx, const double *d_terminal, int64_t n, const hh_payoff *payoffs, int npay,
                            int64_t) {
  cudaStream_t st = ctx->stream;
  int kp_log2 = 0;
  while ((1 << kp_log2) < npay) kp_log2++;
  HH_CUDA(ctx, ctx->d_payoffs.ensure(sizeof(hh_payoff) * (size_t)npay));
  HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_payoffs.ptr, payoffs, sizeof(hh_payoff) * (size_t)npay, cudaMemcpyHostToDevice, st));
  constexpr int NACC = 3;
  const int64_t batches = (n + kThreads - 1) / kThreads;
  int64_t grid = (int64_t)ctx->sm_count * 4;
  if (grid > batches) grid = batches;
  HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)grid * npay * NACC));
  HH_CUDA(ctx, ctx->d_final.ensure(sizeof(double) * (size_t)npay * NACC));
  terminal_payoff_kernel<<<(unsigned)grid, kThreads, 0, st>>>(d_terminal, n, ctx->d_payoffs.as<hh_payoff>(), npay, kp_log2,
                                                              ctx->d_partials.as<double>());
  HH_CUDA(ctx, cudaGetLastError());
  finalize_kernel<<<npay, kThreads, 0, st>>>(ctx->d_partials.as<double>(), (int)grid, npay, NACC, ctx->d_final.as<double>());
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
  ctx->pend.active = true;
  ctx->pend.npay = npay;
  ctx->pend.nblocks = (int)grid;
  ctx->pend.n = n;
  ctx->pend.anti = false;
  ctx->pend.want_terminal = false;
  ctx->pend.bk = false;
  return HH_OK;
}

// Large device -> host copies into the CALLER's pageable buffer (MonteCarloSolution.ensemble: 800 MB at config C2;
// LSMSolution.stopping_info: 40 + 80 MB at config C3). A plain cudaMemcpy to pageable memory runs at ~4 GB/s here (the
// driver stages it and the fresh pages fault in one by one). Instead: chunks DMA into a pinned double buffer on the
// stream while a few host threads copy the previous chunk into the caller's buffer (first-touch page faults in
// parallel). Chunks are a quarter of the job, between 4 and 32 MB, so that medium-sized outputs are pipelined too.
int copy_to_pageable_host(hh_ctx *ctx, void *dst, const void *src_dev, size_t bytes, cudaStream_t st) {
  constexpr size_t kMaxChunk = (size_t)32 << 20, kMinChunk = (size_t)4 << 20, kMB = (size_t)1 << 20;
  if (bytes < 2 * kMinChunk) {
    HH_CUDA(ctx, cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    return HH_OK;
  }
  for (int b = 0; b < 2; ++b) {
    if (!ctx->h_stage[b]) HH_CUDA(ctx, cudaHostAlloc(&ctx->h_stage[b], kMaxChunk, cudaHostAllocDefault));
    if (!ctx->ev_stage[b]) HH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_stage[b], cudaEventDisableTiming));
  }
  size_t quarter = ((bytes / 4) + kMB - 1) & ~(kMB - 1);
  const size_t kChunk = quarter < kMinChunk ? kMinChunk : quarter > kMaxChunk ? kMaxChunk : quarter;
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  auto chunk_bytes = [&](size_t i) { return i + 1 < nchunks ? kChunk : bytes - i * kChunk; };
  auto issue = [&](size_t i) -> cudaError_t {
    cudaError_t e = cudaMemcpyAsync(ctx->h_stage[i & 1], static_cast<const char *>(src_dev) + i * kChunk, chunk_bytes(i),
                                    cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_stage[i & 1], st);
    return e;
  };
  unsigned hw = std::thread::hardware_concurrency();
  const int nthreads = (int)(hw == 0 ? 4 : hw > 16 ? 16 : hw);
  HH_CUDA(ctx, issue(0));
  if (nchunks > 1) HH_CUDA(ctx, issue(1));
  for (size_t i = 0; i < nchunks; ++i) {
    HH_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage[i & 1]));
    const size_t nb = chunk_bytes(i);
    char *d = static_cast<char *>(dst) + i * kChunk;
    const char *sp = static_cast<const char *>(ctx->h_stage[i & 1]);
    std::vector<std::thread> pool;
    const size_t per = ((nb / nthreads) + 4095) & ~(size_t)4095;
    for (int t = 0; t < nthreads; ++t) {
      const size_t o = (size_t)t * per;
      if (o >= nb) break;
      const size_t len = o + per < nb ? per : nb - o;
      pool.emplace_back([=] { memcpy(d + o, sp + o, len); });
    }
    for (auto &th : pool) th.join();
    if (i + 2 < nchunks) HH_CUDA(ctx, issue(i + 2));  // the buffer just drained
  }
  return HH_OK;
}

int european_collect(hh_ctx *ctx, double discount, hh_result *results, double *terminal, size_t terminal_len) {
  if (!ctx->pend.active) return ctx->fail(HH_ERR_ARG, "collect without a pending launch");
  const int npay = ctx->pend.npay;
  const int64_t N = ctx->pend.n;
  constexpr int NACC = nacc_of<double>();
  if (!results) return ctx->fail(HH_ERR_ARG, "results is NULL");
  const size_t tlen = (size_t)N * (ctx->pend.anti ? 2 : 1);
  if (terminal) {
    if (!ctx->pend.want_terminal) return ctx->fail(HH_ERR_ARG, "terminal requested at collect but not at launch");
    if (terminal_len < tlen) return ctx->fail(HH_ERR_ARG, "terminal buffer too short: %zu < %zu", terminal_len, tlen);
  }
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  std::vector<double> fin((size_t)npay * NACC);
  // (the copy of the sums into pageable memory blocks the host until the stream has drained: it comes AFTER the
  // segment copies, which must run while the later segments are still being simulated)
  if (terminal && ctx->pend.nseg > 1) {
    if (!ctx->copy_stream) HH_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    int64_t lo = 0;
    for (int i = 0; i < ctx->pend.nseg; ++i) {  // segment i travels while segments i+1.. are still being simulated
      const int64_t hi = ctx->pend.seg_end[i];
      HH_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_seg[i], 0));
      const int rc = copy_to_pageable_host(ctx, terminal + lo, ctx->d_terminal.as<double>() + lo, sizeof(double) * (size_t)(hi - lo),
                                           ctx->copy_stream);
      if (rc) return rc;
      HH_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
      lo = hi;
    }
  } else if (terminal) {
    const int rc = copy_to_pageable_host(ctx, terminal, ctx->d_terminal.ptr, sizeof(double) * tlen, st);
    if (rc) return rc;
  }
  HH_CUDA(ctx, cudaMemcpyAsync(fin.data(), ctx->d_final.ptr, sizeof(double) * fin.size(), cudaMemcpyDeviceToHost, st));
  unsigned long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ctx->pend.bk)
    HH_CUDA(ctx, cudaMemcpyAsync(counters, ctx->d_counters.ptr, sizeof counters, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->pend.active = false;
  ctx->pend.nseg = 0;
  if (ctx->pend.bk)
    for (int i = 0; i < 5; ++i) ctx->bk_stats[i] = (double)counters[i];
  float ms = 0.f;
  HH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  for (int k = 0; k < npay; ++k) {
    hh_result *r = &results[k];
    memset(r, 0, sizeof *r);
    r->sum = fin[(size_t)k * NACC + 0];
    r->sumsq = fin[(size_t)k * NACC + 1];
    r->n = N;
    const double mean = r->sum / (double)N;
    r->price = discount * mean;  // montecarlo.jl:489-490
    double var = N > 1 ? (r->sumsq - (double)N * mean * mean) / (double)(N - 1) : 0.0;
    if (var < 0) var = 0;
    r->std_error = discount * sqrt(var / (double)N);
    r->n_nonfinite = (int64_t)fin[2];
    r->n_fallback = ctx->pend.bk ? (int64_t)counters[0] : 0;
    r->kernel_ms = ms;
  }
  return HH_OK;
}

int tangent_sums(hh_ctx *ctx, const hh_model *m, const hh_tangent *tg, int ntan, const hh_sim *s,
                 const hh_payoff *payoffs, int npay, double *sums, double spot_bump, double *second_sums, double *kernel_ms) {
  int rc = validate_model_sim(ctx, m, s);
  if (rc) return rc;
  if (!tg || ntan < 1 || ntan > kMaxTan) return ctx->fail(HH_ERR_ARG, "ntangents must be in [1, %d] (got %d)", kMaxTan, ntan);
  if (npay < 1 || npay > kThreads || !payoffs || !sums)
    return ctx->fail(HH_ERR_ARG, "npayoffs must be in [1, %d] and payoffs/sums non-NULL", kThreads);
  if (s->scheme == HH_SCHEME_HESTON_BK)
    return ctx->fail(HH_ERR_UNSUPPORTED, "pathwise tangents are not defined through the Broadie-Kaya sampler");
  if (s->precision != HH_PREC_F64) return ctx->fail(HH_ERR_UNSUPPORTED, "tangents run in f64 only");
  if (s->rng_mode == HH_RNG_PHILOX_64) return ctx->fail(HH_ERR_UNSUPPORTED, "HH_RNG_PHILOX_64 prices only; tangents run on HH_RNG_PHILOX");
  // d_partials / d_final are shared with a pending hh_mc_european_launch
  if (ctx->pend.active) return ctx->fail(HH_ERR_ARG, "a European launch is pending on this context: collect it first");

  const bool anti = s->vr == HH_VR_ANTITHETIC;
  const int P = ntan <= 1 ? 1 : ntan <= 2 ? 2 : ntan <= 4 ? 4 : 8;
  const int NACC = 3 + 2 * P + kNGamma;
  EuroArgs a;
  TangentPack tp;
  int kind = 0;
  build_args(ctx, m, s, tg, ntan, a, tp, kind);
  if (second_sums) {
    if (!(spot_bump > 0.0) || !(spot_bump < m->S0))
      return ctx->fail(HH_ERR_ARG, "second order: the absolute spot bump must be in (0, S0) (got %g)", spot_bump);
    a.g_on = 1;
    a.g_up = (m->S0 + spot_bump) / m->S0;
    a.g_dn = (m->S0 - spot_bump) / m->S0;
    a.g_iu = 1.0 / (m->S0 + spot_bump);
    a.g_id = 1.0 / (m->S0 - spot_bump);
  }

  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  rc = upload_inputs(ctx, m, s, payoffs, npay, a);
  if (rc) return rc;
  HH_CUDA(ctx, ctx->d_tangents.ensure(sizeof(TangentPack)));
  HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_tangents.ptr, &tp, sizeof tp, cudaMemcpyHostToDevice, st));
  const TangentPack *dtp = ctx->d_tangents.as<TangentPack>();

  // Heston EM with the in-kernel RNG takes the specialised tangent kernel: directions in kernel order = full ones
  // first (they need the per-step recursion), then the trivial ones (only dS0 / dr non-zero)
  static const bool generic_tangent = getenv("HH_TANGENT_GENERIC") != nullptr;
  const bool special = kind == K_HESTON_EM && !a.parity && !generic_tangent;
  HestonTanConsts hc;
  memset(&hc, 0, sizeof hc);
  int order[kMaxTan], nfull = 0;  // order[kernel position] = caller's direction index
  if (special) {
    const PathParams<double> &pp = a.p;
    bool full[kMaxTan];
    for (int q = 0; q < ntan; ++q)
      full[q] = tp.v0[q] != 0.0 || tp.kappa[q] != 0.0 || tp.theta[q] != 0.0 || tp.xi[q] != 0.0 || tp.a11[q] != 0.0 ||
                tp.a12[q] != 0.0 || tp.a21[q] != 0.0 || tp.a22[q] != 0.0;
    int pos = 0;
    for (int q = 0; q < ntan; ++q) if (full[q]) order[pos++] = q;
    nfull = pos;
    for (int q = 0; q < ntan; ++q) if (!full[q]) order[pos++] = q;
    for (int kq = 0; kq < ntan; ++kq) {
      const int q = order[kq];
      hc.dx0[kq] = tp.x0[q];
      hc.dv0[kq] = tp.v0[q];
      hc.drift[kq] = (double)a.n_steps * (pp.dt * tp.r[q]);
      hc.A[kq] = pp.dt * (tp.kappa[q] * pp.theta + pp.kappa * tp.theta[q]);
      hc.B[kq] = pp.dt * tp.kappa[q];
      hc.da11[kq] = tp.a11[q];
      hc.da12[kq] = tp.a12[q];
      hc.db21[kq] = tp.xi[q] * pp.a21 + pp.xi * tp.a21[q];
      hc.db22[kq] = tp.xi[q] * pp.a22 + pp.xi * tp.a22[q];
    }
  } else {
    for (int q = 0; q < ntan; ++q) order[q] = q;
  }

  int nblocks = 0;
  if (special) HH_CUDA(ctx, launch_tangent(a, hc, nfull, P, anti, ctx->sm_count, st, &nblocks, true));
  else HH_CUDA(ctx, launch_any(kind, a, dtp, P, anti, ctx->sm_count, st, &nblocks, true));
  HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)nblocks * npay * NACC));
  HH_CUDA(ctx, ctx->d_final.ensure(sizeof(double) * (size_t)npay * NACC));
  a.partials = ctx->d_partials.as<double>();

  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  if (special) HH_CUDA(ctx, launch_tangent(a, hc, nfull, P, anti, ctx->sm_count, st, &nblocks, false));
  else HH_CUDA(ctx, launch_any(kind, a, dtp, P, anti, ctx->sm_count, st, &nblocks, false));
  finalize_kernel<<<npay, kThreads, 0, st>>>(a.partials, nblocks, npay, NACC, ctx->d_final.as<double>());
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));

  std::vector<double> fin((size_t)npay * NACC);
  HH_CUDA(ctx, cudaMemcpyAsync(fin.data(), ctx->d_final.ptr, sizeof(double) * fin.size(), cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  HH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  if (kernel_ms) *kernel_ms = ms;
  const int stride = 2 + 2 * ntan;
  for (int k = 0; k < npay; ++k) {
    const double *f = fin.data() + (size_t)k * NACC;
    double *o = sums + (size_t)k * stride;
    o[0] = f[0];
    o[1] = f[1];
    for (int kq = 0; kq < ntan; ++kq) {  // kernel order -> the caller's order
      o[2 + order[kq]] = f[3 + kq];
      o[2 + ntan + order[kq]] = f[3 + P + kq];
    }
    if (second_sums)
      for (int c = 0; c < kNGamma; ++c) second_sums[(size_t)k * kNGamma + c] = f[3 + 2 * P + c];
  }
  return HH_OK;
}

// hh_bench_heston_ablation: the headline instantiation (no antithetic, split step, base-seed keys, 1024 threads) with one
// part of its step removed, on the C2 model. One launch, timed with events on the context stream.
int heston_ablation(hh_ctx *ctx, int64_t n_paths, int n_steps, int rng_mode, int part, double *ms_out) {
  if (!ms_out || n_paths < 1 || n_steps < 1) return ctx->fail(HH_ERR_ARG, "ablation: n_paths, n_steps >= 1 and ms non-NULL");
  if (rng_mode != HH_RNG_PHILOX && rng_mode != HH_RNG_PHILOX_64) return ctx->fail(HH_ERR_ARG, "ablation: rng_mode must be a Philox stream");
  if (part < 0 || part > 2) return ctx->fail(HH_ERR_ARG, "ablation: part must be 0, 1 or 2");
  if (ctx->pend.active) return ctx->fail(HH_ERR_ARG, "a European launch is pending on this context: collect it first");
  hh_model m;
  memset(&m, 0, sizeof m);
  m.kind = HH_MODEL_HESTON;
  m.flags = HH_FLAG_SPLIT_STEP;
  m.S0 = 100.0, m.r = 0.03, m.T = 1.0, m.V0 = 0.04, m.kappa = 2.0, m.theta = 0.04, m.xi = 0.3, m.rho = -0.7;
  m.m11 = 1.0, m.m12 = 0.0, m.m21 = m.rho, m.m22 = sqrt(1.0 - m.rho * m.rho);
  hh_sim s;
  memset(&s, 0, sizeof s);
  s.n_paths = n_paths, s.n_steps = n_steps, s.scheme = HH_SCHEME_EM, s.rng_mode = rng_mode, s.base_seed = 42;
  const hh_payoff pay = {100.0, 1.0};
  EuroArgs a;
  TangentPack tp;
  int kind = 0;
  build_args(ctx, &m, &s, nullptr, 0, a, tp, kind);
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int rc = upload_inputs(ctx, &m, &s, &pay, 1, a);
  if (rc) return rc;
  int nb = 0;
  HH_CUDA(ctx, (launch_fast2_one<false, true, true, 1, 1024, 1>(a, ctx->sm_count, st, &nb, true)));
  HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)nb * 3));
  a.partials = ctx->d_partials.as<double>();
  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  cudaError_t e = cudaErrorInvalidValue;
  const bool r64 = rng_mode == HH_RNG_PHILOX_64;
  if (part == 0) e = r64 ? launch_fast2_one<false, true, true, 1, 1024, 1, true, 0>(a, ctx->sm_count, st, &nb, false)
                         : launch_fast2_one<false, true, true, 1, 1024, 1, false, 0>(a, ctx->sm_count, st, &nb, false);
  if (part == 1) e = r64 ? launch_fast2_one<false, true, true, 1, 1024, 1, true, 1>(a, ctx->sm_count, st, &nb, false)
                         : launch_fast2_one<false, true, true, 1, 1024, 1, false, 1>(a, ctx->sm_count, st, &nb, false);
  if (part == 2) e = r64 ? launch_fast2_one<false, true, true, 1, 1024, 1, true, 2>(a, ctx->sm_count, st, &nb, false)
                         : launch_fast2_one<false, true, true, 1, 1024, 1, false, 2>(a, ctx->sm_count, st, &nb, false);
  HH_CUDA(ctx, e);
  HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  HH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_out = ms;
  return HH_OK;
}

}  // namespace hh
