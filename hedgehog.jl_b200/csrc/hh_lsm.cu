// hh_lsm.cu — Longstaff-Schwartz American pricing on stored paths: solve(::PricingProblem{American}, ::LSM),
// reference src/pricing_methods/least_squares_montecarlo.jl:99-136.
//
//   simulate_paths + extract_spot_grid (:105-107, :47-85)   -> lsm_paths_kernel: one trajectory (or antithetic pair)
//       per thread, GBM exact steps in registers, every date stored ONCE to a date-major grid G[t][col]
//       (coalesced 8 B stores; col = i for the normal path, N + i for its antithetic partner, as the
//       reference's [normal | antithetic] column order).
//   backward induction (:112-130)                            -> one lsm_pass_kernel per exercise date, HBM-streaming:
//       reads the two date slices G[t+1], G[t] and the per-column cash flow z (f64, discounted to the current
//       date), applies the exercise decision of date t+1 with the coefficients fitted in the previous pass,
//       discounts one step, and accumulates the regression moments of date t over the in-the-money columns
//       (warp shuffle -> shared -> per-block partials, fixed order). 32 B of HBM traffic per column-date.
//   Polynomials.fit (:124-126)                                -> lsm_fit_kernel: the (degree+1)^2 normal equations.
//       The reference regresses on raw monomials of S (QR); here the same polynomial space is spanned by Chebyshev
//       polynomials of an affinely mapped spot u = a_t S + b_t — the interval mapped to [-1, 1] follows the reach of the
//       spot at each date (see hh_lsm_american: uab) — so that the Gram matrix is well conditioned (raw
//       monomials of S ~ 100 to degree 5 give entries ~1e20), and T_i T_j = (T_{i+j} + T_{|i-j|}) / 2 means only
//       2 deg + 1 moment sums are needed for it. The fitted VALUES are those of the reference's least-squares
//       polynomial up to rounding; this is a reduction, not a dense contraction, so no tensor cores.
//   update_stopping_info! (:156-165)                          -> folded into the next pass (strict >), tau written only
//       when stopping info is requested.
//   price (:132-133)                                          -> last pass: sum / sum of squares of D z.
// Multi-GPU: columns are sharded; the per-date moment vector (3 deg + 3 doubles) is sum-allreduced through the
// caller's hh_comm callback between the pass and the fit, so every rank fits the same global polynomial.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "hh_ctx.h"
#include "hh_fastnormal.cuh"
#include "hh_paths.cuh"
#include "hh_tma.cuh"

namespace hh {

constexpr int kLsmThreads = 256;
constexpr int kLsmMaxDeg = 8;

struct LsmPathArgs {
  int64_t n, path_offset, stride;  // stride: columns per date slice (padded)
  uint64_t base_seed;
  const uint64_t *seeds;
  const double *normals;  // parity mode: Z[path][step]
  double *grid;
  int n_steps, parity;
  double S0, dt_drift, sig_sqdt;
  PhiloxRoundKeys rk;          // round keys of base_seed (uniform across threads when seeds == NULL)
  uint32_t one_hi, magic_hi;   // 0x3FF00000, 0x43300000 as arguments (single-LOP3 bit assembly, see hh_european.cu)
};

// Dynamic shared memory of the path generator: [log table x8 | trig table x8 | exponent table | expm1 table]
constexpr int kLsmPathSmem = kLogRepBytes + kTrigRepBytes + kExp2Bytes + kExpm1TabBytes;

// 64.9 KB of tables per block allow 3 blocks per SM: 512-thread blocks give 48 warps per SM to hide the dependent
// chains (Philox rounds -> log -> sqrt -> exp), 32 registers per thread (ncu at 24 warps: issue slots 59 % busy, "wait"
// the top stall)
constexpr int kLsmPathThreads = 512;

// SMALL: the host has proven |(r - s^2/2) dt +- s sqrt(dt) z| <= 1/2 for every normal the in-kernel stream can produce
// (|z| <= 8.6), so fast_expm1_small runs without its range test (never with caller-supplied normals).
// R64: the opt-in HH_RNG_PHILOX_64 stream — 64 random bits per Box-Muller pair (hh_fastnormal.cuh), i.e. ONE Philox block
// per FOUR steps: steps 4b .. 4b+3 take the pairs of words (0, 1) and (2, 3) of block b (counter stream word 2), which is
// hho_normal_pair64(key, idx, n >> 1) component n & 1 in the oracle.
template <bool ANTI, bool PARITY, bool UKEY, bool SMALL, bool R64>
__global__ void __launch_bounds__(kLsmPathThreads, 2) lsm_paths_kernel(const LsmPathArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  char *s_log = reinterpret_cast<char *>(dsm);
  char *s_trig = s_log + kLogRepBytes;
  double *s_e2 = reinterpret_cast<double *>(s_trig + kTrigRepBytes);
  double2 *s_exp = reinterpret_cast<double2 *>(reinterpret_cast<char *>(s_e2) + kExp2Bytes);
  const int tid = threadIdx.x;
  if (!PARITY) {
    for (int e = tid; e < tables::kLog2Buckets * kRep; e += kLsmPathThreads)
      reinterpret_cast<double2 *>(s_log)[e] = g_fast_tables2.log_tab[e / kRep];
    for (int e = tid; e < tables::kTrigN * kRep; e += kLsmPathThreads)
      reinterpret_cast<double2 *>(s_trig)[e] = g_fast_tables2.trig_tab[e / kRep];
    for (int e = tid; e < tables::kExp2N; e += kLsmPathThreads) s_e2[e] = g_fast_tables2.exp_tab[e];
  }
  fill_expm1_table(s_exp);
  __syncthreads();
  const char *log_lane = s_log + (tid & (kRep - 1)) * 16;
  const char *trig_lane = s_trig + (tid & (kRep - 1)) * 16;
  const char *exp_biased = reinterpret_cast<const char *>(s_e2) - tables::kExp2Bias * 8;
  const int M = a.n_steps;
  for (int64_t i = (int64_t)blockIdx.x * kLsmPathThreads + tid; i < a.n; i += (int64_t)gridDim.x * kLsmPathThreads) {
    uint64_t idx = (uint64_t)(a.path_offset + i);
    PhiloxRoundKeys rk_own;
    if (!UKEY && !PARITY) {
      rk_own = philox_round_keys(a.seeds[i]);
      idx = 0;
    }
    const double *z = PARITY ? a.normals + (size_t)i * (size_t)M : nullptr;
    double Sp = a.S0, Sm = a.S0;
    double *gp = a.grid + i;
    double *gm = a.grid + a.n + i;
    gp[0] = Sp;
    if (ANTI) gm[0] = Sm;
    constexpr int NB = R64 ? 4 : 2;  // steps per Philox block
#pragma unroll 1
    for (int n = 0; n < M; n += NB) {
      double zq[NB];
      if (PARITY) {
#pragma unroll
        for (int h = 0; h < NB; ++h) zq[h] = n + h < M ? z[n + h] : 0.0;
      } else if (R64) {
        const u32x4 w = philox4x32_10_rk((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(n >> 2), 2u, UKEY ? a.rk : rk_own);
        fast_normal_pair64_v2(log_lane, exp_biased, trig_lane, w.x, w.y, a.one_hi, a.magic_hi, zq[0], zq[1]);
        fast_normal_pair64_v2(log_lane, exp_biased, trig_lane, w.z, w.w, a.one_hi, a.magic_hi, zq[NB - 2], zq[NB - 1]);
      } else {
        const u32x4 w = philox4x32_10_rk((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(n >> 1), 0u, UKEY ? a.rk : rk_own);
        fast_normal_pair_v2(log_lane, exp_biased, trig_lane, w.x, w.y, w.z, w.w, a.one_hi, a.magic_hi, zq[0], zq[1]);
      }
#pragma unroll
      for (int h = 0; h < NB; ++h) {
        if (n + h < M) {
          const double zz = zq[h];
          // GeometricBrownianMotionProcess increment [upstream]: S += S (exp((r - s^2/2) dt + s sqrt(dt) Z) - 1)
          Sp = fma(Sp, fast_expm1_small<!SMALL>(s_exp, fma(a.sig_sqdt, zz, a.dt_drift)), Sp);
          gp += a.stride;
          *gp = Sp;
          if (ANTI) {  // same normals, sigma -> -sigma (montecarlo.jl:270-284)
            Sm = fma(Sm, fast_expm1_small<!SMALL>(s_exp, fma(-a.sig_sqdt, zz, a.dt_drift)), Sm);
            gm += a.stride;
            *gm = Sm;
          }
        }
      }
    }
  }
}

// ---- log-space generators --------------------------------------------------------------------------
// The reference's extract_spot_grid takes component 1 of the saved state raw (least_squares_montecarlo.jl:53), which
// under EulerMaruyama / HestonNoise is log S, not S: it only tests BlackScholesExact. This generator is the
// corrected extraction for those schemes — the state is advanced in log space exactly as the pricing kernels do
// (LogGBMProblem / LogHestonProblem, heston.jl:7-52) and every date stores S = exp(x), so the backward induction
// below sees spots whatever the scheme.
struct LsmLogArgs {
  LsmPathArgs b;
  PathParams<double> p;
  HestonFolded f;  // host-folded step constants (native-RNG mode)
  int split;
};

constexpr int kLsmLogSmem = kLogRepBytes + kTrigRepBytes + kExpFullBytes + kExp2Bytes;

template <bool HESTON, bool ANTI, bool PARITY, bool UKEY>
__global__ void __launch_bounds__(kLsmPathThreads, 2) lsm_logspace_paths_kernel(const LsmLogArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  char *s_log = reinterpret_cast<char *>(dsm);
  char *s_trig = s_log + kLogRepBytes;
  double *s_expf = reinterpret_cast<double *>(s_trig + kTrigRepBytes);
  double *s_e2 = s_expf + kExpFullN;
  const int tid = threadIdx.x;
  if (!PARITY) {
    for (int e = tid; e < tables::kLog2Buckets * kRep; e += kLsmPathThreads)
      reinterpret_cast<double2 *>(s_log)[e] = g_fast_tables2.log_tab[e / kRep];
    for (int e = tid; e < tables::kTrigN * kRep; e += kLsmPathThreads)
      reinterpret_cast<double2 *>(s_trig)[e] = g_fast_tables2.trig_tab[e / kRep];
    for (int e = tid; e < tables::kExp2N; e += kLsmPathThreads) s_e2[e] = g_fast_tables2.exp_tab[e];
    fill_exp_full_table(s_expf);
  }
  __syncthreads();
  const char *log_lane = s_log + (tid & (kRep - 1)) * 16;
  const char *trig_lane = s_trig + (tid & (kRep - 1)) * 16;
  const char *exp_biased = reinterpret_cast<const char *>(s_e2) - tables::kExp2Bias * 8;
  // spots from log-spots: the table-driven exp with the in-kernel RNG, libm's in parity mode
  auto spot_of = [&](double x) { return PARITY ? exp(x) : fast_exp_full(s_expf, x); };
  const LsmPathArgs &b = a.b;
  const PathParams<double> &p = a.p;
  const bool split = a.split != 0;
  const int M = b.n_steps;
  constexpr int NC = HESTON ? 2 : 1;
  for (int64_t i = (int64_t)blockIdx.x * kLsmPathThreads + tid; i < b.n; i += (int64_t)gridDim.x * kLsmPathThreads) {
    uint64_t idx = (uint64_t)(b.path_offset + i);
    PhiloxRoundKeys rk_own;
    if (!UKEY && !PARITY) {
      rk_own = philox_round_keys(b.seeds[i]);
      idx = 0;
    }
    const double *z = PARITY ? b.normals + (size_t)i * (size_t)M * NC : nullptr;
    double xp = p.x0, xm = p.x0, vp = p.v0, vm = p.v0;
    double *gp = b.grid + i;
    double *gm = b.grid + b.n + i;
    gp[0] = b.S0;
    if (ANTI) gm[0] = b.S0;
    if (HESTON) {
#pragma unroll 1
      for (int n = 0; n < M; ++n) {
        double z1, z2;
        if (PARITY) {
          z1 = z[2 * n];
          z2 = z[2 * n + 1];
        } else {
          const u32x4 w = philox4x32_10_rk((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)n, 0u, UKEY ? b.rk : rk_own);
          fast_normal_pair_v2(log_lane, exp_biased, trig_lane, w.x, w.y, w.z, w.w, b.one_hi, b.magic_hi, z1, z2);
        }
        const double dW1 = fma(p.a12, z2, p.a11 * z1);
        if (PARITY) {
          const double dW2 = fma(p.a22, z2, p.a21 * z1);
          heston_em_step<double>(p, split, xp, vp, dW1, dW2);
          if (ANTI) heston_em_step<double>(p, split, xm, vm, -dW1, -dW2);  // NoiseGrid(t, -W), montecarlo.jl:258
        } else {  // host-folded constants, xi folded into dW2, branch-free clamps and sqrt (hh_paths.cuh)
          const double xdW2 = fma(a.f.b22, z2, a.f.b21 * z1);
          heston_em_step_fast(a.f, split, xp, vp, dW1, xdW2);
          if (ANTI) heston_em_step_fast(a.f, split, xm, vm, -dW1, -xdW2);
        }
        gp += b.stride;
        *gp = spot_of(xp);
        if (ANTI) {
          gm += b.stride;
          *gm = spot_of(xm);
        }
      }
    } else {
#pragma unroll 1
      for (int n = 0; n < M; n += 2) {
        double za, zb;
        if (PARITY) {
          za = z[n];
          zb = n + 1 < M ? z[n + 1] : 0.0;
        } else {
          const u32x4 w = philox4x32_10_rk((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(n >> 1), 0u, UKEY ? b.rk : rk_own);
          fast_normal_pair_v2(log_lane, exp_biased, trig_lane, w.x, w.y, w.z, w.w, b.one_hi, b.magic_hi, za, zb);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (n + h < M) {
            const double dW = p.sqdt * (h ? zb : za);
            gbm_em_step<double>(p, xp, dW);
            gp += b.stride;
            *gp = spot_of(xp);
            if (ANTI) {
              gbm_em_step<double>(p, xm, -dW);
              gm += b.stride;
              *gm = spot_of(xm);
            }
          }
        }
      }
    }
  }
}

// Fitted polynomial of one date, in the Chebyshev basis of u = a S + b. active = 0: the date was skipped
// (no in-the-money column, least_squares_montecarlo.jl:122) or is the terminal date.
struct LsmFit {
  double c[kLsmMaxDeg + 1];
  double q[kLsmMaxDeg + 1];  // the same polynomial in powers of u (Horner form for the pass kernel); q[0] = +inf when inactive
  double active;
  double used_degree;
  double count;
};

// Exchange of the regression moments with the peer GPUs (mailboxes mapped through CUDA IPC, see lsm_peer_exchange)
struct PeerX {
  int world, rank;
  unsigned long long epoch;  // unique per exchanged date, identical on every rank; its parity selects the slot
  double *mail[HH_MAX_PEERS];
  int *error;                // set to 1 if a peer's contribution did not arrive in time
  long long timeout_cycles;  // in-kernel limit on one wait (hh_peer_set_timeout; SM clock cycles)
};

struct LsmPassArgs {
  int64_t ncols;
  const double *S_next;  // G[t+1]
  const double *S_cur;   // G[t]   (unused in the last pass)
  double *z;             // cash flow per column, discounted to the current date
  int32_t *tau;          // nullable
  const LsmFit *fit_next;
  double *partials;      // [grid][nacc]
  double Dp, rscale, strike, cp, ua, ub;  // Dp = D^(t+1); rscale = D^-t turns the time-0 sums into date-t money for the fit
  double ua_n, ub_n;       // the Chebyshev variable of date t+1 (the fitted polynomial of that date is in it); ua, ub: date t
  int t_next;
  int first;  // t+1 is the terminal date: z = payoff(S_M)
  int last;   // t = 0: no regression, accumulate sum / sumsq of D z
  int reverse;             // walk the columns from the end (alternates per pass for L2 reuse)
  unsigned int *done;      // arrival counter of the blocks of this pass
  double *moments;         // [nacc] sums over all blocks, written by the last block
  LsmFit *fit_out;         // nullable: where the last block writes the fit of date t (single-GPU form)
  PeerX px;                // peer exchange in the kernel tail (hh_peer_*): world > 1 switches it on
};

constexpr int kMailSlot = 32;  // doubles per (parity, rank) slot; flags follow the payload (see hh_api.cu)
__device__ __forceinline__ double *mail_payload(double *mail, int parity, int rank) {
  return mail + ((size_t)parity * HH_MAX_PEERS + rank) * kMailSlot;
}
__device__ __forceinline__ unsigned long long *mail_flag(double *mail, int parity, int rank) {
  return reinterpret_cast<unsigned long long *>(mail + 2 * HH_MAX_PEERS * kMailSlot) + parity * HH_MAX_PEERS + rank;
}
// "abort" word behind the flags: a rank that gives up waiting sets it in EVERY rank's mailbox, so that all ranks leave
// the exchange with the same error instead of one raising while the others finish (and later hang in a host collective)
__device__ __forceinline__ unsigned long long *mail_abort(double *mail) {
  return reinterpret_cast<unsigned long long *>(mail + 2 * HH_MAX_PEERS * kMailSlot) + 2 * HH_MAX_PEERS;
}

template <int DEG>
__device__ __forceinline__ double clenshaw(const double *c, double u) {
  // sum_k c_k T_k(u)
  double b1 = 0.0, b2 = 0.0;
  const double u2 = 2.0 * u;
#pragma unroll
  for (int k = DEG; k >= 1; --k) {
    const double b0 = fma(u2, b1, c[k] - b2);
    b2 = b1;
    b1 = b0;
  }
  return fma(u, b1, c[0] - b2);
}

// accumulators: m[0..2 DEG] = sum T_k(u),  r[0..DEG] = sum T_k(u) y,  count   (last pass: sum, sumsq, count)
template <int DEG>
__host__ __device__ constexpr int lsm_nacc() { return 3 * DEG + 3; }

// One column of one pass, branch free. The pass kernel is bound by the SM's dispatch port, not by HBM, unless the
// per-column instruction count is kept near the minimum (profiles/r1_c_ncu_lsm_pass_2e6.csv: DRAM 25 %, issue 57 %):
//   decision at t+1:  e = cp S - cp K;  cont = Horner(q, ua_n S + ub_n);  exercise iff e > 0 and e > cont (strict, :163-164)
//   moments at t:     g = 1{cp S_t - cp K > 0};  T_0 = g, T_1 = g u, T_k = 2 u T_{k-1} - T_{k-2}: the indicator rides
//                     through the linear recurrence, so every sum is an unconditional add / fma.
// Sign tests read the high word of the double on the integer pipe (x > 0 <=> hi(x) > 0 for the values that occur:
// cp (S - K) is either 0 or at least one ulp of S).
// Cash flows are kept in TIME-0 money, z = D^tau v (what the reference averages at the end, :132-133): a column's z only
// changes when it is exercised, so most columns are never rewritten (the pass is L2-throughput bound and z stores were a
// sixth of its L2 traffic), and the regression target of date t, D^(tau-t) v (:117-118), is z D^-t — a factor common to
// all columns that is applied to the right-hand-side sums in the fit instead of per column.
template <int DEG, bool FIRST, bool LAST, bool TAU, class A>
__device__ __forceinline__ void lsm_column(const A &a, const double *q, double cpK, double sn, double sc,
                                           double zin, int64_t p, double &zout, bool &changed, double *acc, int &cnt) {
  constexpr int NM = 2 * DEG + 1;
  const double e = fma(a.cp, sn, -cpK);
  const double e_now = e * a.Dp;  // exercise value at date t+1 in time-0 money, Dp = D^(t+1)
  double zz;
  if (FIRST) {
    zz = __double2hiint(e) > 0 ? e_now : 0.0;  // stopping_info = (nsteps, payoff(S_T))  :112
    changed = true;
  } else {
    const double un = fma(a.ua_n, sn, a.ub_n);
    double cont = q[DEG];  // poly.(x)  :127  (in date-(t+1) money, like e)
#pragma unroll
    for (int k = DEG - 1; k >= 0; --k) cont = fma(cont, un, q[k]);
    const bool ex = (__double2hiint(e) > 0) && (e > cont);
    zz = ex ? e_now : zin;
    changed = ex;
    if (TAU && ex) a.tau[p] = a.t_next;
  }
  zout = zz;
  if (LAST) {
    acc[0] += zz;
    acc[1] = fma(zz, zz, acc[1]);
    cnt += 1;
  } else {
    const double e0 = fma(a.cp, sc, -cpK);
    const bool itm = __double2hiint(e0) > 0;  // in the money  :120-121
    const double u = itm ? fma(a.ua, sc, a.ub) : 0.0;
    const double g = itm ? 1.0 : 0.0;
    cnt += itm ? 1 : 0;
    acc[NM] = fma(g, zz, acc[NM]);
    if (DEG >= 1 || NM > 1) {
      const double u2 = u + u;
      double t0 = g, t1 = u;
      acc[1] += u;
      if (DEG >= 1) acc[NM + 1] = fma(u, zz, acc[NM + 1]);
#pragma unroll
      for (int k = 2; k < NM; ++k) {
        const double tk = fma(u2, t1, -t0);
        t0 = t1;
        t1 = tk;
        acc[k] += tk;
        if (k <= DEG) acc[NM + k] = fma(tk, zz, acc[NM + k]);
      }
    }
  }
}

// Sum the per-block partials in a fixed order: moments[c] = sum_b partials[b][c]. The whole block takes part:
// thread (c, g) sums blocks g, g + G, g + 2G, ... with eight independent accumulators (all loads in flight at once:
// one L2 round trip instead of a dependent chain), then the G group sums of an accumulator are added in order.
// scratch: >= 256 doubles of shared memory. Deterministic for a given grid size.
__device__ __forceinline__ void lsm_reduce_partials(const double *partials, int nblocks, int nacc, double *moments,
                                                    double *scratch) {
  const int G = (int)blockDim.x / nacc;
  const int c = (int)threadIdx.x % nacc, g = (int)threadIdx.x / nacc;
  if (g < G) {
    double t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = 0.0;
    int b = g;
    for (; b + 7 * G < nblocks; b += 8 * G) {
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] += __ldcg(partials + (size_t)(b + k * G) * nacc + c);
    }
    for (int k = 0; b < nblocks; b += G, ++k) t[k & 7] += __ldcg(partials + (size_t)b * nacc + c);
    scratch[g * nacc + c] = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
  }
  __syncthreads();
  if ((int)threadIdx.x < nacc) {
    double v = 0.0;
    for (int gg = 0; gg < G; ++gg) v += scratch[gg * nacc + threadIdx.x];
    moments[threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) lsm_reduce_kernel(const double *partials, int nblocks, int nacc, double *moments) {
  __shared__ double scratch[256];
  lsm_reduce_partials(partials, nblocks, nacc, moments, scratch);
}

// 1 / sqrt(d) to full double precision without the IEEE division/sqrt sequences (the fit is a serial chain on one
// thread in the tail of every pass): MUFU.RSQ64H seed, two Newton steps.
__device__ __forceinline__ double rsqrt_fast(double d) {
  double y = rsqrt_seed(d);
  double e = fma(-d * y, y, 1.0);
  y = fma(y * e, fma(e, 0.375, 0.5), y);
  e = fma(-d * y, y, 1.0);
  return fma(y * e, 0.5, y);
}

// Normal equations in the Chebyshev basis: G[i][j] = (m[i+j] + m[|i-j|]) / 2, rhs[i] = r[i]; Cholesky.
// If the matrix is numerically singular at the requested degree (fewer distinct in-the-money spots than
// coefficients), the leading block that factorises is used (a lower-degree fit in the same nested basis).
// DEG is a compile-time constant so that every array lives in registers.
template <int DEG>
__device__ __forceinline__ void lsm_fit(const double *moments, double rscale, LsmFit *out) {
  constexpr int nm = 2 * DEG + 1;
  double m[nm], r[DEG + 1];
#pragma unroll
  for (int k = 0; k < nm; ++k) m[k] = moments[k];
#pragma unroll
  for (int k = 0; k <= DEG; ++k) r[k] = moments[nm + k] * rscale;  // time-0 money -> date-t money
  const double count = moments[nm + DEG + 1];
  double c[DEG + 1];
#pragma unroll
  for (int k = 0; k <= DEG; ++k) c[k] = 0.0;
  double active = 0.0, used = -1.0;
  if (count > 0.0) {
    double L[DEG + 1][DEG + 1], inv[DEG + 1];
    int n = 0;  // size of the leading block that factorises
    bool ok = true;
#pragma unroll
    for (int j = 0; j <= DEG; ++j) {
      double d = 0.5 * (m[2 * j] + m[0]);
#pragma unroll
      for (int k = 0; k < j; ++k) d = fma(-L[j][k], L[j][k], d);
      ok = ok && (d > 1e-13 * m[0]);
      const double rs = rsqrt_fast(ok ? d : 1.0);
      inv[j] = rs;
      L[j][j] = d * rs;
#pragma unroll
      for (int i = j + 1; i <= DEG; ++i) {
        double s = 0.5 * (m[i + j] + m[i - j]);
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[i][k], L[j][k], s);
        L[i][j] = s * rs;
      }
      if (ok) n = j + 1;
    }
    if (n > 0) {
      double y[DEG + 1];
#pragma unroll
      for (int i = 0; i <= DEG; ++i) {
        double s = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-L[i][k], y[k], s);
        y[i] = i < n ? s * inv[i] : 0.0;
      }
#pragma unroll
      for (int i = DEG; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k <= DEG; ++k) s = fma(-L[k][i], c[k], s);  // c[k] = 0 for k >= n
        c[i] = i < n ? s * inv[i] : 0.0;
      }
      active = 1.0;
      used = (double)(n - 1);
    }
  }
  // powers of u: T_0 = 1, T_1 = u, T_{k+1} = 2 u T_k - T_{k-1}
  double q[DEG + 1], ta[DEG + 1], tb[DEG + 1];
#pragma unroll
  for (int k = 0; k <= DEG; ++k) q[k] = ta[k] = tb[k] = 0.0;
  ta[0] = 1.0;
  q[0] = c[0];
  if (DEG >= 1) {
    tb[1] = 1.0;
    q[1] = c[1];
  }
#pragma unroll
  for (int k = 2; k <= DEG; ++k) {
    double tc[DEG + 1];
#pragma unroll
    for (int j = 0; j <= DEG; ++j) tc[j] = (j > 0 ? 2.0 * tb[j - 1] : 0.0) - ta[j];
#pragma unroll
    for (int j = 0; j <= DEG; ++j) {
      q[j] = fma(c[k], tc[j], q[j]);
      ta[j] = tb[j];
      tb[j] = tc[j];
    }
  }
  if (active == 0.0) q[0] = __longlong_as_double(0x7ff0000000000000LL);  // never exercise against +inf
#pragma unroll
  for (int k = 0; k <= kLsmMaxDeg; ++k) {
    out->c[k] = k <= DEG ? c[k < DEG ? k : DEG] : 0.0;
    out->q[k] = k <= DEG ? q[k < DEG ? k : DEG] : 0.0;
  }
  out->active = active;
  out->used_degree = used;
  out->count = count;
}

template <int DEG>
__global__ void lsm_fit_kernel(const double *moments, double rscale, LsmFit *out) {
  if (threadIdx.x == 0) lsm_fit<DEG>(moments, rscale, out);
}

// Peer exchange, called by ONE block per rank with all its threads (one process per GPU, mailboxes mapped through CUDA
// IPC): the local sums mom[0..nacc) go into slot [parity][my rank] of EVERY rank's mailbox (plain stores to the mapped peer
// pointers, i.e. over NVLink), a system-scope fence, then the release flag; the block waits for the flags of all ranks
// in its own mailbox and adds the slots in rank order, so all ranks hold bit-identical global sums in mom[]. Two
// parities alternate per date: a rank can only be one date ahead of a peer (it needs the peer's sums to finish a
// date), so the slot written at date d+2 has been consumed.
__device__ __forceinline__ void lsm_peer_exchange(const PeerX &px, unsigned long long epoch, double *mom, int nacc) {
  const int tid = threadIdx.x;
  const int parity = (int)(epoch & 1ull);
  if (tid < nacc) {
    const double v = mom[tid];
    for (int qq = 0; qq < px.world; ++qq) mail_payload(px.mail[qq], parity, px.rank)[tid] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < px.world) {
    unsigned long long *f = mail_flag(px.mail[tid], parity, px.rank);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    // wait for rank `tid`'s flag in MY mailbox, bounded by px.timeout_cycles (hh_peer_set_timeout). Giving up poisons
    // every rank's mailbox, and a poisoned mailbox ends every wait: all ranks return HH_ERR_PEER_TIMEOUT together.
    const unsigned long long *mine = mail_flag(px.mail[px.rank], parity, tid);
    const unsigned long long *poison = mail_abort(px.mail[px.rank]);
    const long long t0 = clock64();
    unsigned long long seen = 0, ab = 0;
    bool dead = *reinterpret_cast<volatile int *>(px.error) != 0;
    unsigned spins = 0;
    while (!dead) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= epoch) break;
      if ((++spins & 63u) == 0u) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(ab) : "l"(poison) : "memory");
        if (ab != 0ull) {
          *px.error = 1;
          dead = true;
        } else if (clock64() - t0 > px.timeout_cycles) {
          for (int qq = 0; qq < px.world; ++qq)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(mail_abort(px.mail[qq])), "l"(1ull) : "memory");
          *px.error = 1;
          dead = true;
        }
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (tid < nacc) {
    double v = 0.0;
    for (int qq = 0; qq < px.world; ++qq)
      v += *reinterpret_cast<volatile double *>(mail_payload(px.mail[px.rank], parity, qq) + tid);
    mom[tid] = v;
  }
  __syncthreads();
}

// Tail of a pass: block reduction -> per-block partials -> the block that arrives last sums all partials in a fixed
// order, exchanges them with the peer GPUs if there are any (lsm_peer_exchange), and fits the date's polynomial.
template <int DEG, bool LAST>
__device__ __forceinline__ void lsm_pass_tail(const LsmPassArgs &a, double *acc, int cnt, double (*s_red)[kLsmThreads / 32],
                                              double *scratch, bool *s_last) {
  constexpr int NACC = lsm_nacc<DEG>();
  constexpr int NM = 2 * DEG + 1;
  const int tid = threadIdx.x;
  // the counts were kept on the integer pipe: m[0] = sum of the indicator, and the trailing count slot
  if (LAST) {
    acc[2] = (double)cnt;
  } else {
    acc[0] = (double)cnt;
    acc[NM + DEG + 1] = (double)cnt;
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int c = 0; c < NACC; ++c) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[c][warp] = v;
  }
  __syncthreads();
  if (tid < NACC) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kLsmThreads / 32; ++w) t += s_red[tid][w];
    a.partials[(size_t)blockIdx.x * NACC + tid] = t;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) *s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  lsm_reduce_partials(a.partials, (int)gridDim.x, NACC, a.moments, scratch);
  __syncthreads();
  if (a.px.world > 1) lsm_peer_exchange(a.px, a.px.epoch, a.moments, NACC);
  if (tid == 0) {
    if (a.fit_out) lsm_fit<DEG>(a.moments, a.rscale, a.fit_out);
    *a.done = 0u;
  }
}

template <int DEG, bool FIRST, bool LAST, bool TAU>
__global__ void __launch_bounds__(kLsmThreads) lsm_pass_kernel(const LsmPassArgs a) {
  constexpr int NACC = lsm_nacc<DEG>();
  constexpr int NM = 2 * DEG + 1;
  __shared__ double s_red[NACC][kLsmThreads / 32];
  __shared__ double s_scratch[kLsmThreads];
  __shared__ bool s_last;
  double q[DEG + 1];
#pragma unroll
  for (int k = 0; k <= DEG; ++k) q[k] = FIRST ? 0.0 : a.fit_next->q[k];
  const double cpK = a.cp * a.strike;
  double acc[NACC];
#pragma unroll
  for (int c = 0; c < NACC; ++c) acc[c] = 0.0;
  int cnt = 0;

  // Two columns per thread per iteration (16 B loads and stores; the slices are 256 B aligned), and the loads of the
  // next iteration are issued before the arithmetic of the current one so that every thread keeps 6 x 16 B in flight.
  const int64_t npairs = a.ncols >> 1;
  const int64_t step = (int64_t)gridDim.x * kLsmThreads;
  const int64_t first = (int64_t)blockIdx.x * kLsmThreads + threadIdx.x;
  const int iters = first < npairs ? (int)((npairs - 1 - first) / step) + 1 : 0;
  // passes alternate their direction over the columns: what the previous pass touched last (z and the shared date
  // slice) is what this pass touches first, while it is still in L2
  const bool rev = a.reverse != 0;
  int64_t idx = rev ? npairs - 1 - first : first;  // pair index of this thread's current iteration
  const int64_t dstep = rev ? -step : step;
  const double2 *pn = reinterpret_cast<const double2 *>(a.S_next) + idx;
  const double2 *pc = reinterpret_cast<const double2 *>(a.S_cur) + idx;
  double2 *pz = reinterpret_cast<double2 *>(a.z) + idx;
  // software pipeline, unrolled by two with ping-pong register sets (no register moves): the loads of iteration
  // i+1 are in flight while iteration i computes
  double2 snA = make_double2(0.0, 0.0), scA = snA, ziA = snA, snB = snA, scB = snA, ziB = snA;
  auto load = [&](double2 &sn, double2 &sc, double2 &zi) {
    sn = __ldcs(pn);
    if (!LAST) sc = __ldg(pc);
    if (!FIRST) zi = *pz;
  };
  auto work = [&](const double2 &sn, const double2 &sc, const double2 &zi, double2 *zdst, int64_t col) {
    double2 zo;
    bool cx, cy;
    lsm_column<DEG, FIRST, LAST, TAU>(a, q, cpK, sn.x, sc.x, zi.x, col, zo.x, cx, acc, cnt);
    lsm_column<DEG, FIRST, LAST, TAU>(a, q, cpK, sn.y, sc.y, zi.y, col + 1, zo.y, cy, acc, cnt);
    if (cx || cy) *zdst = zo;
  };
  if (iters > 0) load(snA, scA, ziA);
  int it = 0;
  for (; it + 2 <= iters; it += 2) {
    double2 *zA = pz;
    const int64_t colA = 2 * idx;
    pn += dstep; pc += dstep; pz += dstep; idx += dstep;
    load(snB, scB, ziB);
    work(snA, scA, ziA, zA, colA);
    double2 *zB = pz;
    const int64_t colB = 2 * idx;
    pn += dstep; pc += dstep; pz += dstep; idx += dstep;
    if (it + 2 < iters) load(snA, scA, ziA);
    work(snB, scB, ziB, zB, colB);
  }
  if (it < iters) work(snA, scA, ziA, pz, 2 * idx);
  if ((a.ncols & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t p = a.ncols - 1;
    double zo;
    bool ch;
    lsm_column<DEG, FIRST, LAST, TAU>(a, q, cpK, a.S_next[p], LAST ? 0.0 : a.S_cur[p], FIRST ? 0.0 : a.z[p], p, zo, ch, acc, cnt);
    if (ch) a.z[p] = zo;
  }
  lsm_pass_tail<DEG, LAST>(a, acc, cnt, s_red, s_scratch, &s_last);
}

// ---- the same pass with TMA-staged streaming ------------------------------------------------------------------
// ncu on the register-prefetch version (profiles/r1_d_ncu_lsm_pass_1e7.csv): long-scoreboard stalls 11.8 per issue,
// issue slots 35 % busy, DRAM at 70 % of peak — latency bound: one 16 B load per array per thread in flight is not
// enough bytes in flight per SM. Here thread 0 of each block issues 1-D bulk copies (cp.async.bulk -> UBLKCP) of
// whole 512-column chunks of G[t+1], G[t] and z into a kStages-deep ring in shared memory, tracked by one mbarrier per
// stage; the block consumes chunk i while chunks i+1 .. i+kStages-1 are in flight (12 KB per chunk, 36 KB per block).
constexpr int kLsmChunk = 2 * kLsmThreads;  // columns per chunk: one double2 per thread per array
constexpr int kLsmStages = 3;

template <int DEG, bool FIRST, bool LAST, bool TAU>
__global__ void __launch_bounds__(kLsmThreads) lsm_pass_tma_kernel(const LsmPassArgs a) {
  constexpr int NACC = lsm_nacc<DEG>();
  constexpr int NM = 2 * DEG + 1;
  __shared__ __align__(128) double s_next[kLsmStages][kLsmChunk];
  __shared__ __align__(128) double s_cur[LAST ? 1 : kLsmStages][LAST ? 2 : kLsmChunk];
  __shared__ __align__(128) double s_z[FIRST ? 1 : kLsmStages][FIRST ? 2 : kLsmChunk];
  __shared__ __align__(8) uint64_t s_full[kLsmStages];
  __shared__ double s_red[NACC][kLsmThreads / 32];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kLsmStages; ++s) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  double q[DEG + 1];
#pragma unroll
  for (int k = 0; k <= DEG; ++k) q[k] = FIRST ? 0.0 : a.fit_next->q[k];
  const double cpK = a.cp * a.strike;
  double acc[NACC];
#pragma unroll
  for (int c = 0; c < NACC; ++c) acc[c] = 0.0;
  int cnt = 0;

  const int64_t neven = a.ncols & ~(int64_t)1;  // the odd last column is handled by one thread below
  const int64_t nchunks = (neven + kLsmChunk - 1) / kLsmChunk;
  const int my_chunks = (int64_t)blockIdx.x < nchunks ? (int)((nchunks - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  const bool rev = a.reverse != 0;  // alternate direction: start where the previous pass ended (still in L2)
  auto chunk_of = [&](int i) {
    const int64_t c = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
    return rev ? nchunks - 1 - c : c;
  };
  const uint64_t pol_first = l2_policy_evict_first();
  auto issue = [&](int i) {  // thread 0 only
    const int s = i % kLsmStages;
    const int64_t c0 = chunk_of(i) * kLsmChunk;
    const int64_t ncol = neven - c0 < kLsmChunk ? neven - c0 : kLsmChunk;
    const uint32_t bytes = (uint32_t)ncol * 8u;
    mbar_arrive_expect_tx(&s_full[s], bytes * (1u + (LAST ? 0u : 1u) + (FIRST ? 0u : 1u)));
    bulk_load_hint(&s_next[s][0], a.S_next + c0, bytes, &s_full[s], pol_first);  // dead after this pass
    if (!LAST) bulk_load(&s_cur[s][0], a.S_cur + c0, bytes, &s_full[s]);
    if (!FIRST) bulk_load(&s_z[s][0], a.z + c0, bytes, &s_full[s]);
  };
  if (tid == 0) {
    for (int i = 0; i < kLsmStages && i < my_chunks; ++i) issue(i);
  }
  for (int i = 0; i < my_chunks; ++i) {
    const int s = i % kLsmStages;
    const uint32_t parity = (uint32_t)(i / kLsmStages) & 1u;
    const int64_t c0 = chunk_of(i) * kLsmChunk;
    const int64_t ncol = neven - c0 < kLsmChunk ? neven - c0 : kLsmChunk;
    const bool mine = 2 * tid < ncol;
    mbar_wait(&s_full[s], parity);
    double2 sn = make_double2(0.0, 0.0), sc = sn, zi = sn;
    if (mine) {
      sn = reinterpret_cast<const double2 *>(&s_next[s][0])[tid];
      if (!LAST) sc = reinterpret_cast<const double2 *>(&s_cur[s][0])[tid];
      if (!FIRST) zi = reinterpret_cast<const double2 *>(&s_z[s][0])[tid];
    }
    __syncthreads();  // every thread has taken its columns out of stage s
    if (tid == 0 && i + kLsmStages < my_chunks) issue(i + kLsmStages);
    if (mine) {
      const int64_t col = c0 + 2 * tid;
      double2 zo;
      bool cx, cy;
      lsm_column<DEG, FIRST, LAST, TAU>(a, q, cpK, sn.x, sc.x, zi.x, col, zo.x, cx, acc, cnt);
      lsm_column<DEG, FIRST, LAST, TAU>(a, q, cpK, sn.y, sc.y, zi.y, col + 1, zo.y, cy, acc, cnt);
      if (cx || cy) *reinterpret_cast<double2 *>(a.z + col) = zo;
    }
  }
  if ((a.ncols & 1) && blockIdx.x == 0 && tid == 0) {
    const int64_t p = a.ncols - 1;
    double zo;
    bool ch;
    lsm_column<DEG, FIRST, LAST, TAU>(a, q, cpK, a.S_next[p], LAST ? 0.0 : a.S_cur[p], FIRST ? 0.0 : a.z[p], p, zo, ch, acc, cnt);
    if (ch) a.z[p] = zo;
  }
  lsm_pass_tail<DEG, LAST>(a, acc, cnt, s_red, &s_next[0][0] /* the ring is idle by now */, &s_last);
}

// ---- the whole backward induction in ONE persistent cooperative kernel ---------------------------------------------
// One launch per date costs ~9.5 us of fixed time at C3 (drain, last-block reduction, launch, ramp): 0.47 ms of 2.4.
// Here every block stays resident for all dates (cooperative launch guarantees co-residency), owns a fixed set of
// 512-column chunks, and the dates are separated by one grid barrier: after it EVERY block sums the per-block partials of
// the date in the same fixed order and fits the polynomial itself (0.5 us of serial work, done redundantly instead of
// broadcast). Partials are double-buffered by date parity, so a block that races ahead writes the other buffer.
// With peers, block 0 exchanges the local sums (lsm_peer_exchange) and publishes the global sums through a flag.
struct LsmBackArgs {
  int64_t ncols, stride;
  const double *G;         // [M+1][stride]
  double *z;
  int32_t *tau;            // nullable
  double *partials;        // [2][grid][nacc]
  unsigned int *barrier;   // grid barrier counter, zeroed before the launch
  double *gmom;            // [2][32] global sums published by block 0 (peer mode)
  unsigned int *gflag;     // generation of gmom
  double *moments_out;     // final [sum, sumsq, count]
  LsmFit *fits;            // [M+1], written by block 0 (statistics for the host)
  double logD, strike, cp;  // logD = log of the one-step discount factor
  const double *uab;       // [M+1][2]: Chebyshev variable u = ua S + ub of every date (hh_lsm_american: uab)
  int M;
  PeerX px;                // px.epoch = epoch of the first exchanged date minus one
  int z_policy, cur_policy;  // L2 hints of the bulk loads: 0 none, 1 evict_last, 2 evict_first
};
struct LsmColArgs {
  double cp, ua, ub, ua_n, ub_n, Dp;
  int32_t *tau;
  int t_next;
};

__device__ __forceinline__ void lsm_grid_barrier(unsigned int *counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

constexpr int kBackCpt = 2;                            // columns per consumer thread and chunk (4: 2.24 ms, 2: 2.07 ms at C3)
constexpr int kBackChunk = kBackCpt * kLsmThreads;     // 512 columns: 4 KB per array, 12 KB per stage
constexpr int kBackStages = 4;
constexpr int kBackSmem = kBackStages * 3 * kBackChunk * 8;
constexpr int kBackThreads = kLsmThreads + 32;         // 8 consumer warps + 1 producer warp

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// 16-byte load from the shared window (ordered after the mbarrier wait that precedes it: volatile, memory clobber)
__device__ __forceinline__ double2 lds_v2(unsigned int saddr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr) : "memory");
  return v;
}

// Warp-specialised: warp 8 is the producer (one lane issues the bulk copies as ring slots drain), warps 0-7 consume.
// ncu on the version whose consumers met at a __syncthreads per chunk (profiles/r1_j_ncu_lsm_backward.csv): "barrier" was the
// top stall (2.4 cycles per issue), issue slots 49 % busy, DRAM 49 %. Here a consumer warp announces that it has taken
// its columns out of a stage by arriving on the stage's `empty` mbarrier and goes on computing; only the producer waits.
template <int DEG, bool TAU>
__global__ void __launch_bounds__(kBackThreads) lsm_backward_kernel(const LsmBackArgs a) {
  constexpr int NACC = lsm_nacc<DEG>();
  constexpr int NM = 2 * DEG + 1;
  constexpr int NCW = kLsmThreads / 32;  // consumer warps
  // ring in dynamic shared memory: [stage][array (next, cur, z)][kBackChunk doubles]
  extern __shared__ __align__(128) unsigned char dsm_back[];
  double *ring = reinterpret_cast<double *>(dsm_back);
  auto s_next = [&](int s) { return ring + ((size_t)s * 3 + 0) * kBackChunk; };
  auto s_cur = [&](int s) { return ring + ((size_t)s * 3 + 1) * kBackChunk; };
  auto s_z = [&](int s) { return ring + ((size_t)s * 3 + 2) * kBackChunk; };
  __shared__ __align__(8) uint64_t s_full[kBackStages], s_empty[kBackStages];
  __shared__ double s_red[NACC][NCW];
  __shared__ double s_mom[32];
  __shared__ double s_scratch[kBackThreads + 32];
  __shared__ LsmFit s_fit;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool producer = warp == NCW;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kBackStages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NCW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const double cpK = a.cp * a.strike;
  const unsigned int ring_saddr = smem_addr(ring) + (unsigned)tid * 16u;  // this thread's first column pair of stage 0
  const int64_t neven = a.ncols & ~(int64_t)1;
  const int64_t nchunks = (neven + kBackChunk - 1) / kBackChunk;
  const int64_t cstep = (int64_t)gridDim.x * kBackChunk;
  const int my_chunks = (int64_t)blockIdx.x < nchunks ? (int)((nchunks - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  const uint64_t pol_first = l2_policy_evict_first();
  const uint64_t pol_last = l2_policy_evict_last();
  unsigned int seq = 0;  // chunks of this block so far, over all dates (same count on both sides): ring slot and phase
  LsmColArgs ca;
  ca.cp = a.cp;
  ca.tau = a.tau;

  // one lane of the producer warp: chunk i of date `td` (its S_next = G[td+1]) into ring position `at`, after the
  // consumers have drained what was there
  auto issue_date = [&](int td, int i, unsigned int at) {
    const bool f = (td + 1 == a.M), l = (td == 0);
    const int s = (int)(at % kBackStages);
    mbar_wait(&s_empty[s], ((at / kBackStages) & 1u) ^ 1u);  // passes at once the first time round the ring
    const int64_t c0 = (int64_t)blockIdx.x * kBackChunk + (int64_t)i * cstep;
    const int64_t ncol = neven - c0 < kBackChunk ? neven - c0 : kBackChunk;
    const uint32_t bytes = (uint32_t)ncol * 8u;
    mbar_arrive_expect_tx(&s_full[s], bytes * (1u + (l ? 0u : 1u) + (f ? 0u : 1u)));
    bulk_load_hint(s_next(s), a.G + (size_t)(td + 1) * a.stride + c0, bytes, &s_full[s], pol_first);  // dead after date td
    if (!l) {
      if (a.cur_policy == 0) bulk_load(s_cur(s), a.G + (size_t)td * a.stride + c0, bytes, &s_full[s]);
      else bulk_load_hint(s_cur(s), a.G + (size_t)td * a.stride + c0, bytes, &s_full[s], a.cur_policy == 1 ? pol_last : pol_first);
    }
    if (!f) {
      if (a.z_policy == 0) bulk_load(s_z(s), a.z + c0, bytes, &s_full[s]);
      else bulk_load_hint(s_z(s), a.z + c0, bytes, &s_full[s], a.z_policy == 1 ? pol_last : pol_first);
    }
  };
  const int nprime = my_chunks < kBackStages ? my_chunks : kBackStages;  // chunks of a date issued ahead of its start
  if (producer && lane == 0) {
    for (int i = 0; i < nprime; ++i) issue_date(a.M - 1, i, (unsigned)i);
  }

  for (int t = a.M - 1; t >= 0; --t) {
    const int date = a.M - 1 - t;  // 0, 1, ...
    const bool first = (t + 1 == a.M), last = (t == 0);
    const double *S_next = a.G + (size_t)(t + 1) * a.stride;
    const double *S_cur = a.G + (size_t)t * a.stride;
    ca.t_next = t + 1;
    ca.Dp = exp((double)(t + 1) * a.logD);
    ca.ua = __ldg(a.uab + 2 * t);
    ca.ub = __ldg(a.uab + 2 * t + 1);
    ca.ua_n = __ldg(a.uab + 2 * t + 2);
    ca.ub_n = __ldg(a.uab + 2 * t + 3);
    double q[DEG + 1];
#pragma unroll
    for (int k = 0; k <= DEG; ++k) q[k] = first ? 0.0 : s_fit.q[k];
    double acc[NACC];
#pragma unroll
    for (int c = 0; c < NACC; ++c) acc[c] = 0.0;
    int cnt = 0;

    if (producer) {
      // the first nprime chunks of this date are already in flight (issued before the previous grid barrier)
      if (lane == 0)
        for (int i = nprime; i < my_chunks; ++i) issue_date(t, i, seq + (unsigned)i);
      seq += (unsigned)my_chunks;
    } else {
      auto chunks = [&](auto first_c, auto last_c) {
        constexpr bool F = decltype(first_c)::value, L = decltype(last_c)::value;
        int64_t c0 = (int64_t)blockIdx.x * kBackChunk;
        // stage and phase advance incrementally; the ring is addressed through the shared window (a 32-bit address kept in a
        // register: the generic form recomputes the window base every chunk). Only the last chunk of the grid can be ragged:
        // every other chunk takes the path without per-column predicates (SASS: 138 -> ~115 instructions per chunk and thread).
        unsigned int s = seq % kBackStages, parity = (seq / kBackStages) & 1u;
        for (int i = 0; i < my_chunks; ++i, c0 += cstep) {
          mbar_wait(&s_full[s], parity);
          const unsigned int st_addr = ring_saddr + s * (unsigned)(3 * kBackChunk * sizeof(double));
          if (c0 + kBackChunk <= neven) {
            double2 sn[kBackCpt / 2], sc[kBackCpt / 2], zi[kBackCpt / 2];
#pragma unroll
            for (int h = 0; h < kBackCpt / 2; ++h) {
              const unsigned int ad = st_addr + (unsigned)(h * kLsmThreads * 16);
              sn[h] = lds_v2(ad);
              sc[h] = L ? make_double2(0.0, 0.0) : lds_v2(ad + (unsigned)(kBackChunk * sizeof(double)));
              zi[h] = F ? make_double2(0.0, 0.0) : lds_v2(ad + (unsigned)(2 * kBackChunk * sizeof(double)));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);  // this warp has taken its columns out of stage s
#pragma unroll
            for (int h = 0; h < kBackCpt / 2; ++h) {
              const int64_t col = c0 + 2 * (tid + h * kLsmThreads);
              double2 zo;
              bool cx, cy;
              lsm_column<DEG, F, L, TAU>(ca, q, cpK, sn[h].x, sc[h].x, zi[h].x, col, zo.x, cx, acc, cnt);
              lsm_column<DEG, F, L, TAU>(ca, q, cpK, sn[h].y, sc[h].y, zi[h].y, col + 1, zo.y, cy, acc, cnt);
              if (cx || cy) *reinterpret_cast<double2 *>(a.z + col) = zo;
            }
          } else {
            const int ncol = (int)(neven - c0);
            // kBackCpt / 2 column pairs per thread, a block-wide stride apart (conflict-free 16 B shared loads)
            double2 sn[kBackCpt / 2], sc[kBackCpt / 2], zi[kBackCpt / 2];
#pragma unroll
            for (int h = 0; h < kBackCpt / 2; ++h) {
              const int pr = tid + h * kLsmThreads;
              sn[h] = sc[h] = zi[h] = make_double2(0.0, 0.0);
              if (2 * pr < ncol) {
                sn[h] = reinterpret_cast<const double2 *>(s_next((int)s))[pr];
                if (!L) sc[h] = reinterpret_cast<const double2 *>(s_cur((int)s))[pr];
                if (!F) zi[h] = reinterpret_cast<const double2 *>(s_z((int)s))[pr];
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
#pragma unroll
            for (int h = 0; h < kBackCpt / 2; ++h) {
              const int pr = tid + h * kLsmThreads;
              if (2 * pr < ncol) {
                const int64_t col = c0 + 2 * pr;
                double2 zo;
                bool cx, cy;
                lsm_column<DEG, F, L, TAU>(ca, q, cpK, sn[h].x, sc[h].x, zi[h].x, col, zo.x, cx, acc, cnt);
                lsm_column<DEG, F, L, TAU>(ca, q, cpK, sn[h].y, sc[h].y, zi[h].y, col + 1, zo.y, cy, acc, cnt);
                if (cx || cy) *reinterpret_cast<double2 *>(a.z + col) = zo;
              }
            }
          }
          s = (s + 1) & (kBackStages - 1);
          parity ^= (s == 0);
        }
        seq += (unsigned)my_chunks;
        if ((a.ncols & 1) && blockIdx.x == 0 && tid == 0) {
          const int64_t p = a.ncols - 1;
          double zo;
          bool ch;
          lsm_column<DEG, F, L, TAU>(ca, q, cpK, S_next[p], L ? 0.0 : S_cur[p], F ? 0.0 : a.z[p], p, zo, ch, acc, cnt);
          if (ch) a.z[p] = zo;
        }
        if (L) {
          acc[2] = (double)cnt;
        } else {
          acc[0] = (double)cnt;
          acc[NM + DEG + 1] = (double)cnt;
        }
      };
      using T_ = std::true_type;
      using F_ = std::false_type;
      if (first && last) chunks(T_{}, T_{});
      else if (first) chunks(T_{}, F_{});
      else if (last) chunks(F_{}, T_{});
      else chunks(F_{}, F_{});
    }

    // block reduction -> partials[date parity][block] (the producer warp holds zeros and stays out)
    double *part = a.partials + (size_t)(date & 1) * gridDim.x * NACC;
    if (!producer) {
#pragma unroll
      for (int c = 0; c < NACC; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[c][warp] = v;
      }
    }
    // this block's z stores of the date must be ordered before the bulk copies (async proxy) that re-read them
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (tid < NACC) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NCW; ++w) v += s_red[tid][w];
      part[(size_t)blockIdx.x * NACC + tid] = v;
    }
    // prime the ring with the first chunks of the NEXT date while the block waits at the grid barrier and fits
    if (producer && lane == 0 && !last) {
      for (int i = 0; i < nprime; ++i) issue_date(t - 1, i, seq + (unsigned)i);
    }
    lsm_grid_barrier(a.barrier, (unsigned)(date + 1) * gridDim.x);

    if (last) {
      if (blockIdx.x == 0) {
        lsm_reduce_partials(part, (int)gridDim.x, NACC, s_mom, s_scratch);
        __syncthreads();
        if (tid < 3) a.moments_out[tid] = s_mom[tid];
      }
      break;
    }
    // every block: the same fixed-order sum of all partials, then the same fit
    lsm_reduce_partials(part, (int)gridDim.x, NACC, s_mom, s_scratch);
    __syncthreads();
    if (a.px.world > 1) {
      double *gm = a.gmom + (size_t)(date & 1) * 32;
      if (blockIdx.x == 0) {
        lsm_peer_exchange(a.px, a.px.epoch + (unsigned long long)(date + 1), s_mom, NACC);
        if (tid < NACC) gm[tid] = s_mom[tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.gflag), "r"((unsigned)(date + 1)) : "memory");
      } else {
        if (tid == 0) {
          unsigned int seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.gflag) : "memory");
          } while (seen < (unsigned)(date + 1));
        }
        __syncthreads();
        if (tid < NACC) s_mom[tid] = __ldcg(gm + tid);
        __syncthreads();
      }
    }
    if (tid == 0) {
      lsm_fit<DEG>(s_mom, exp(-(double)t * a.logD), &s_fit);
      if (blockIdx.x == 0) a.fits[t] = s_fit;
    }
    __syncthreads();
  }
}

// stopping_info values: v_p = payoff(G[tau_p][p])  (:112, :163-164)
__global__ void lsm_stop_values_kernel(const double *grid, int64_t stride, int64_t ncols, const int32_t *tau, double strike,
                                       double cp, double *val) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ncols; p += (int64_t)gridDim.x * blockDim.x)
    val[p] = fmax(cp * (grid[(size_t)tau[p] * stride + p] - strike), 0.0);
}

__global__ void lsm_fill_tau_kernel(int32_t *tau, int64_t ncols, int32_t v) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ncols; p += (int64_t)gridDim.x * blockDim.x) tau[p] = v;
}

// spot_paths for the host: Matrix (nsteps+1) x ncols, column-major (one column = one trajectory, :50):
// out[(p - p0) * (M+1) + t] = G[t][p], transposed through shared memory in 32 x 32 tiles.
__global__ void lsm_transpose_kernel(const double *grid, int64_t stride, int64_t p0, int64_t np, int nrows, double *out) {
  __shared__ double tile[32][33];
  const int64_t pb = (int64_t)blockIdx.x * 32;
  const int tb = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = tb + r;
    const int64_t p = pb + threadIdx.x;
    if (t < nrows && p < np) tile[r][threadIdx.x] = grid[(size_t)t * stride + p0 + p];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t p = pb + r;
    const int t = tb + threadIdx.x;
    if (t < nrows && p < np) out[(size_t)p * nrows + t] = tile[threadIdx.x][r];
  }
}

static const int g_lsm_tma = getenv("HH_LSM_TMA") ? atoi(getenv("HH_LSM_TMA")) : 1;

template <int DEG>
static cudaError_t launch_pass(const LsmPassArgs &a, int grid, cudaStream_t st) {
  if (g_lsm_tma) {
    if (a.first && a.last) lsm_pass_tma_kernel<DEG, true, true, false><<<grid, kLsmThreads, 0, st>>>(a);
    else if (a.first) lsm_pass_tma_kernel<DEG, true, false, false><<<grid, kLsmThreads, 0, st>>>(a);
    else if (a.last && a.tau) lsm_pass_tma_kernel<DEG, false, true, true><<<grid, kLsmThreads, 0, st>>>(a);
    else if (a.last) lsm_pass_tma_kernel<DEG, false, true, false><<<grid, kLsmThreads, 0, st>>>(a);
    else if (a.tau) lsm_pass_tma_kernel<DEG, false, false, true><<<grid, kLsmThreads, 0, st>>>(a);
    else lsm_pass_tma_kernel<DEG, false, false, false><<<grid, kLsmThreads, 0, st>>>(a);
    return cudaGetLastError();
  }
  // the first pass never writes tau (lsm_fill_tau_kernel already set tau = M)
  if (a.first && a.last) lsm_pass_kernel<DEG, true, true, false><<<grid, kLsmThreads, 0, st>>>(a);
  else if (a.first) lsm_pass_kernel<DEG, true, false, false><<<grid, kLsmThreads, 0, st>>>(a);
  else if (a.last && a.tau) lsm_pass_kernel<DEG, false, true, true><<<grid, kLsmThreads, 0, st>>>(a);
  else if (a.last) lsm_pass_kernel<DEG, false, true, false><<<grid, kLsmThreads, 0, st>>>(a);
  else if (a.tau) lsm_pass_kernel<DEG, false, false, true><<<grid, kLsmThreads, 0, st>>>(a);
  else lsm_pass_kernel<DEG, false, false, false><<<grid, kLsmThreads, 0, st>>>(a);
  return cudaGetLastError();
}

template <int DEG>
static int pass_occupancy() {
  int occ = 0;
  cudaError_t e = g_lsm_tma ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lsm_pass_tma_kernel<DEG, false, false, true>, kLsmThreads, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lsm_pass_kernel<DEG, false, false, true>, kLsmThreads, 0);
  if (e != cudaSuccess || occ < 1) occ = 1;
  return occ;
}

static int pass_occupancy_deg(int deg) {
  switch (deg) {
    case 0: return pass_occupancy<0>();
    case 1: return pass_occupancy<1>();
    case 2: return pass_occupancy<2>();
    case 3: return pass_occupancy<3>();
    case 4: return pass_occupancy<4>();
    case 5: return pass_occupancy<5>();
    case 6: return pass_occupancy<6>();
    case 7: return pass_occupancy<7>();
    default: return pass_occupancy<8>();
  }
}

static void launch_fit_deg(int deg, const double *moments, double rscale, LsmFit *out, cudaStream_t st) {
  switch (deg) {
    case 0: lsm_fit_kernel<0><<<1, 32, 0, st>>>(moments, rscale, out); break;
    case 1: lsm_fit_kernel<1><<<1, 32, 0, st>>>(moments, rscale, out); break;
    case 2: lsm_fit_kernel<2><<<1, 32, 0, st>>>(moments, rscale, out); break;
    case 3: lsm_fit_kernel<3><<<1, 32, 0, st>>>(moments, rscale, out); break;
    case 4: lsm_fit_kernel<4><<<1, 32, 0, st>>>(moments, rscale, out); break;
    case 5: lsm_fit_kernel<5><<<1, 32, 0, st>>>(moments, rscale, out); break;
    case 6: lsm_fit_kernel<6><<<1, 32, 0, st>>>(moments, rscale, out); break;
    case 7: lsm_fit_kernel<7><<<1, 32, 0, st>>>(moments, rscale, out); break;
    default: lsm_fit_kernel<8><<<1, 32, 0, st>>>(moments, rscale, out); break;
  }
}

static cudaError_t launch_pass_deg(int deg, const LsmPassArgs &a, int grid, cudaStream_t st) {
  switch (deg) {
    case 0: return launch_pass<0>(a, grid, st);
    case 1: return launch_pass<1>(a, grid, st);
    case 2: return launch_pass<2>(a, grid, st);
    case 3: return launch_pass<3>(a, grid, st);
    case 4: return launch_pass<4>(a, grid, st);
    case 5: return launch_pass<5>(a, grid, st);
    case 6: return launch_pass<6>(a, grid, st);
    case 7: return launch_pass<7>(a, grid, st);
    default: return launch_pass<8>(a, grid, st);
  }
}

template <int DEG, bool TAU>
static cudaError_t launch_backward(const LsmBackArgs &b, int sm_count, int64_t nchunks, cudaStream_t st, bool query_only, int *grid_out) {
  auto kern = lsm_backward_kernel<DEG, TAU>;
  static PerDeviceOnce opted;  // per instantiation and device
  if (cudaError_t e0 = smem_opt_in(opted, kern, kBackSmem); e0 != cudaSuccess) return e0;
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBackThreads, kBackSmem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  static const int cap = getenv("HH_LSM_BLOCKS_PER_SM") ? atoi(getenv("HH_LSM_BLOCKS_PER_SM")) : 2;
  // measured at C3 (1e7 x 50, degree 3): 1 block per SM 2.79 ms, 2: 2.07 ms, 3: 2.29 ms, 4: 2.42 ms — barrier and partial-sum
  // costs grow with the grid, and two 8-warp blocks per SM already keep enough bulk copies in flight
  if (occ > cap) occ = cap;
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > nchunks) grid = nchunks;
  if (grid < 1) grid = 1;
  *grid_out = (int)grid;
  if (query_only) return cudaSuccess;
  LsmBackArgs copy = b;
  void *args[] = {&copy};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kern), dim3((unsigned)grid), dim3(kBackThreads), args, kBackSmem, st);
}

static cudaError_t launch_backward_deg(int deg, bool tau, const LsmBackArgs &b, int sm_count, int64_t nchunks, cudaStream_t st,
                                       bool query_only, int *grid_out) {
#define HH_BACK(D)                                                                                   \
  case D:                                                                                            \
    return tau ? launch_backward<D, true>(b, sm_count, nchunks, st, query_only, grid_out)          \
               : launch_backward<D, false>(b, sm_count, nchunks, st, query_only, grid_out)
  switch (deg) {
    HH_BACK(0);
    HH_BACK(1);
    HH_BACK(2);
    HH_BACK(3);
    HH_BACK(4);
    HH_BACK(5);
    HH_BACK(6);
    HH_BACK(7);
    default:
      return tau ? launch_backward<8, true>(b, sm_count, nchunks, st, query_only, grid_out)
                 : launch_backward<8, false>(b, sm_count, nchunks, st, query_only, grid_out);
  }
#undef HH_BACK
}

int lsm_american(hh_ctx *ctx, const hh_model *m, const hh_sim *s, const hh_payoff *payoff, int degree,
                 double step_discount, const hh_comm *comm, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                 double *spot_paths) {
  int rc = validate_model_sim(ctx, m, s);
  if (rc) return rc;
  if (!payoff || !out) return ctx->fail(HH_ERR_ARG, "payoff/out is NULL");
  if (ctx->pend.active) return ctx->fail(HH_ERR_ARG, "a European launch is pending on this context: collect it first");
  if (s->rng_mode == HH_RNG_PHILOX_64 && !(m->kind == HH_MODEL_GBM && s->scheme == HH_SCHEME_EXACT_STEPS))
    return ctx->fail(HH_ERR_UNSUPPORTED, "HH_RNG_PHILOX_64 under LSM is defined for the exact GBM generator "
                                         "(LognormalDynamics + BlackScholesExact): four steps per Philox block");
  if (degree < 0 || degree > kLsmMaxDeg) return ctx->fail(HH_ERR_ARG, "degree must be in [0, %d] (got %d)", kLsmMaxDeg, degree);
  // Q7: the reference reads component 1 of the saved state as the spot (least_squares_montecarlo.jl:53), which is
  // only true for the S-space BlackScholesExact generator — the only LSM configuration it tests.
  // BlackScholesExact is the drop-in configuration; the log-space EulerMaruyama / HestonNoise generators are accepted
  // too, with the corrected extraction S = exp(x) (lsm_logspace_paths_kernel).
  // HestonBroadieKaya: the exact sampler writes the spots of its n_steps dates (bk_path_grid_launch, hh_bk.cu).
  const bool logspace = s->scheme == HH_SCHEME_EM;
  const bool bk = m->kind == HH_MODEL_HESTON && s->scheme == HH_SCHEME_HESTON_BK;
  if (!(m->kind == HH_MODEL_GBM && s->scheme == HH_SCHEME_EXACT_STEPS) && !logspace && !bk)
    return ctx->fail(HH_ERR_UNSUPPORTED, "LSM runs on BlackScholesExact, EulerMaruyama, HestonNoise or HestonBroadieKaya paths "
                     "(SURVEY Q7)");
  if (s->precision != HH_PREC_F64) return ctx->fail(HH_ERR_UNSUPPORTED, "LSM paths are generated in binary64");
  if ((stop_idx == nullptr) != (stop_val == nullptr))
    return ctx->fail(HH_ERR_ARG, "stop_idx and stop_val must be both NULL or both non-NULL");
  const bool peer_mode = comm && comm->world > 1 && !comm->allreduce_sum_f64;  // in-kernel exchange over peer memory
  if (comm && comm->world <= 1 && !comm->allreduce_sum_f64) comm = nullptr;
  if (peer_mode && (ctx->peer_world != comm->world || ctx->peer_rank != comm->rank))
    return ctx->fail(HH_ERR_ARG, "hh_comm asks for the peer exchange (rank %d of %d) but the context is connected as rank %d of %d "
                     "(hh_peer_export / hh_peer_connect)", comm->rank, comm->world, ctx->peer_rank, ctx->peer_world);

  const int64_t N = s->n_paths;
  const bool anti = s->vr == HH_VR_ANTITHETIC;
  const int64_t ncols = anti ? 2 * N : N;
  const int M = s->n_steps;
  const int64_t stride = (ncols + 31) & ~(int64_t)31;  // 256 B aligned date slices
  const bool parity = s->rng_mode == HH_RNG_NORMALS;
  const bool want_stop = stop_idx != nullptr;

  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  HH_CUDA(ctx, upload_fast_tables(ctx->device, st));
  HH_CUDA(ctx, upload_fast_tables2(ctx->device, st));
  HH_CUDA(ctx, ctx->d_grid.ensure(sizeof(double) * (size_t)stride * (size_t)(M + 1)));
  HH_CUDA(ctx, ctx->d_cash.ensure(sizeof(double) * (size_t)stride));
  if (want_stop) HH_CUDA(ctx, ctx->d_tau.ensure(sizeof(int32_t) * (size_t)stride));

  LsmPathArgs pa;
  memset(&pa, 0, sizeof pa);
  pa.n = N;
  pa.path_offset = s->path_offset;
  pa.stride = stride;
  pa.base_seed = s->base_seed;
  pa.grid = ctx->d_grid.as<double>();
  pa.n_steps = M;
  pa.parity = parity;
  const double dt = m->T / M;
  pa.rk = philox_round_keys(s->base_seed);
  pa.one_hi = 0x3FF00000u;
  pa.magic_hi = 0x43300000u;
  pa.S0 = m->S0;
  pa.dt_drift = (m->r - 0.5 * (m->sigma * m->sigma)) * dt;
  pa.sig_sqdt = m->sigma * sqrt(dt);
  if (parity) {
    const size_t bytes = sizeof(double) * (size_t)N * (size_t)M * (m->kind == HH_MODEL_HESTON ? 2 : 1);
    HH_CUDA(ctx, ctx->d_normals.ensure(bytes));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_normals.ptr, s->normals, bytes, cudaMemcpyHostToDevice, st));
    pa.normals = ctx->d_normals.as<double>();
  } else if (s->seeds) {
    const size_t bytes = sizeof(uint64_t) * (size_t)N;
    HH_CUDA(ctx, ctx->d_seeds.ensure(bytes));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_seeds.ptr, s->seeds, bytes, cudaMemcpyHostToDevice, st));
    pa.seeds = ctx->d_seeds.as<uint64_t>();
  }

  // grids: a multiple of the SM count, capped by the work
  const int64_t path_blocks = (N + kLsmPathThreads - 1) / kLsmPathThreads;
  // 64.9 KB of tables per block: 3 blocks per SM; one resident wave, grid-stride over the trajectories
  const int grid_paths = (int)(path_blocks < (int64_t)ctx->sm_count * 3 ? path_blocks : (int64_t)ctx->sm_count * 3);
  const int64_t pass_blocks = ((ncols >> 1) + kLsmThreads - 1) / kLsmThreads;
  const int64_t resident = (int64_t)ctx->sm_count * pass_occupancy_deg(degree);  // one wave: every block is resident
  int grid_pass = (int)(pass_blocks < resident ? pass_blocks : resident);
  if (grid_pass < 1) grid_pass = 1;
  const int nacc = 3 * degree + 3;
  static const int persistent_env = getenv("HH_LSM_PERSISTENT") ? atoi(getenv("HH_LSM_PERSISTENT")) : 1;
  const int64_t neven = ncols & ~(int64_t)1;
  const int64_t nchunks = (neven + kBackChunk - 1) / kBackChunk;  // chunks of the persistent kernel
  const bool host_exchange = comm && comm->world > 1 && !peer_mode;  // NCCL through the callback needs a launch per date
  int grid_back = 0;
  bool persistent = persistent_env && !host_exchange && nchunks >= 1;
  if (persistent) {
    LsmBackArgs probe;
    memset(&probe, 0, sizeof probe);
    if (launch_backward_deg(degree, want_stop, probe, ctx->sm_count, nchunks, st, true, &grid_back) != cudaSuccess) {
      (void)cudaGetLastError();
      persistent = false;
    }
  }
  {
    const size_t need = (size_t)(persistent && 2 * grid_back > grid_pass ? 2 * grid_back : grid_pass) * nacc;
    HH_CUDA(ctx, ctx->d_lsm_partials.ensure(sizeof(double) * need));
  }
  // state: [moments (nacc)] [done counter] [fits (M+1)]
  const size_t done_off = ((size_t)nacc * sizeof(double) + 255) & ~(size_t)255;
  const size_t fit_off = done_off + 1024;  // +0 done, +64 grid barrier, +128 peer error, +192 gmom flag, +256 gmom[2][32]
  HH_CUDA(ctx, ctx->d_lsm_state.ensure(fit_off + sizeof(LsmFit) * (size_t)(M + 1)));
  double *d_moments = ctx->d_lsm_state.as<double>();
  LsmFit *d_fits = reinterpret_cast<LsmFit *>(static_cast<char *>(ctx->d_lsm_state.ptr) + fit_off);
  HH_CUDA(ctx, cudaMemsetAsync(ctx->d_lsm_state.ptr, 0, fit_off + sizeof(LsmFit) * (size_t)(M + 1), st));

  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  if (bk) {
    rc = bk_path_grid_launch(ctx, m, s, ctx->d_grid.as<double>(), stride);
    if (rc) return rc;
  } else if (logspace) {
    LsmLogArgs la;
    memset(&la, 0, sizeof la);
    la.b = pa;
    la.split = (m->flags & HH_FLAG_SPLIT_STEP) != 0;
    PathParams<double> &p = la.p;
    const double sqdt = sqrt(dt);
    p.dt = dt;
    p.sqdt = sqdt;
    p.x0 = log(m->S0);
    p.S0 = m->S0;
    p.r = m->r;
    const bool heston = m->kind == HH_MODEL_HESTON;
    if (heston) {
      p.v0 = m->V0;
      p.kappa = m->kappa;
      p.theta = m->theta;
      p.xi = m->xi;
      p.a11 = sqdt * m->m11;
      p.a12 = sqdt * m->m12;
      p.a21 = sqdt * m->m21;
      p.a22 = sqdt * m->m22;
      la.f.rdt = m->r * dt;
      la.f.neg_half_dt = -0.5 * dt;
      la.f.neg_kdt = -(m->kappa * dt);
      la.f.ktdt = m->kappa * m->theta * dt;
      la.f.b21 = m->xi * p.a21;
      la.f.b22 = m->xi * p.a22;
    } else {
      p.sigma = m->sigma;
      p.dt_drift = dt * (m->r - 0.5 * (m->sigma * m->sigma));
    }
    const bool ukey = pa.seeds == nullptr;
    cudaError_t le = cudaSuccess;
#define HH_LSM_LOG(H, A, P, U)                                                                                            \
  do {                                                                                                                  \
    static PerDeviceOnce opted;                                                                                         \
    le = smem_opt_in(opted, lsm_logspace_paths_kernel<H, A, P, U>, kLsmLogSmem);                                        \
    if (le == cudaSuccess) lsm_logspace_paths_kernel<H, A, P, U><<<grid_paths, kLsmPathThreads, kLsmLogSmem, st>>>(la); \
  } while (0)
#define HH_LSM_LOG_H(H)                                                                                    \
  do {                                                                                                     \
    if (parity) { if (anti) HH_LSM_LOG(H, true, true, true); else HH_LSM_LOG(H, false, true, true); }     \
    else if (ukey) { if (anti) HH_LSM_LOG(H, true, false, true); else HH_LSM_LOG(H, false, false, true); } \
    else { if (anti) HH_LSM_LOG(H, true, false, false); else HH_LSM_LOG(H, false, false, false); }        \
  } while (0)
    if (heston) HH_LSM_LOG_H(true); else HH_LSM_LOG_H(false);
#undef HH_LSM_LOG_H
#undef HH_LSM_LOG
    HH_CUDA(ctx, le);
  } else {
    const bool ukey = pa.seeds == nullptr;
    cudaError_t le = cudaSuccess;
    // the largest exponent argument the in-kernel stream can produce: |z| <= sqrt(-2 ln 2^-52) = 8.49
    const bool small = !parity && fabs(pa.dt_drift) + 8.6 * fabs(pa.sig_sqdt) <= 0.5;
    const bool r64 = s->rng_mode == HH_RNG_PHILOX_64;
#define HH_LSM_PATHS(A, P, U, S, R)                                                                                     \
  do {                                                                                                                  \
    static PerDeviceOnce opted;                                                                                         \
    le = smem_opt_in(opted, lsm_paths_kernel<A, P, U, S, R>, kLsmPathSmem);                                             \
    if (le == cudaSuccess) lsm_paths_kernel<A, P, U, S, R><<<grid_paths, kLsmPathThreads, kLsmPathSmem, st>>>(pa);      \
  } while (0)
#define HH_LSM_PATHS_S(A, P, U)                                                                  \
  do {                                                                                           \
    if (r64) { if (small) HH_LSM_PATHS(A, P, U, true, true); else HH_LSM_PATHS(A, P, U, false, true); }   \
    else { if (small) HH_LSM_PATHS(A, P, U, true, false); else HH_LSM_PATHS(A, P, U, false, false); }     \
  } while (0)
    if (parity) { if (anti) HH_LSM_PATHS(true, true, true, false, false); else HH_LSM_PATHS(false, true, true, false, false); }
    else if (ukey) { if (anti) HH_LSM_PATHS_S(true, false, true); else HH_LSM_PATHS_S(false, false, true); }
    else { if (anti) HH_LSM_PATHS_S(true, false, false); else HH_LSM_PATHS_S(false, false, false); }
#undef HH_LSM_PATHS_S
#undef HH_LSM_PATHS
    HH_CUDA(ctx, le);
  }
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));

  if (want_stop) {
    lsm_fill_tau_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->d_tau.as<int32_t>(), ncols, M);
    HH_CUDA(ctx, cudaGetLastError());
  }

  // Chebyshev variable u = ua_t S + ub_t of date t: the in-the-money side of the strike, as far as the spot of THAT date
  // reaches (the expected extreme of ncols draws of its lognormal law), mapped to [-1, 1]. The map only conditions the basis — the fitted
  // polynomial is the same — but the fit solves NORMAL equations: with one interval for all dates (the put's (0, K), or the
  // call's (K, 6-sigma excursion of the TERMINAL spot)) the early dates of a volatile call, or of a short-dated low-volatility
  // put, had all their data in a few per cent of [-1, 1], the Gram matrix lost its rank in binary64 and up to 4 % of the
  // stopping decisions differed from the QR fit of the reference (found by HH_FUZZ_SCALE=8 tests/test_gpu_fuzz.py).
  std::vector<double> uab((size_t)(2 * (M + 1)));
  {
    // How far the interval reaches: to where the largest of ncols samples is expected, z_n = Phi^-1(1 - 1 / ncols) standard
    // deviations of log S_t (3.7 at 1e4 columns, 5.2 at 1e7), not further. The Gram matrix of degree 6 is sensitive to this:
    // on a call with sigma sqrt(T) = 1 and 5e4 columns its condition number is 6e5 for the interval [min, max] of the data,
    // 1e7 for z_n = 4.1, 2e13 for 5 sigma (where 0.7 % of the decisions left the reference's) and 3e11 for 3 sigma
    // (points beyond the interval: |u| > 1, T_k(u)^2 explodes). Heston: log S_t has fatter tails than the normal law
    // with the expected integrated variance, + 0.3.
    // The count is that of the JOB, rounded up to a power of two: ranks that hold shards of one job must all use the SAME
    // variable (their moment sums are added), and shard sizes differ by one. The host layers pass the bit length of the
    // job's trajectory count in hh_sim.reserved (0: this call is the whole job), so that 1 GPU and 8 GPUs fit in the same
    // basis and their prices stay equal to the last bit.
    double zn = 4.0;
    {
      int bits = s->reserved;
      if (bits <= 0)
        for (bits = 1; bits < 62 && ((int64_t)1 << bits) <= N; ++bits) {}
      const double job_cols = ldexp(anti ? 2.0 : 1.0, bits < 62 ? bits : 62);
      const double target = 1.0 / job_cols;
      double lo_z = 0.0, hi_z = 8.5;
      for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo_z + hi_z);
        if (0.5 * erfc(mid * 0.7071067811865476) > target) lo_z = mid; else hi_z = mid;
      }
      zn = fmin(fmax(0.5 * (lo_z + hi_z), 3.0), 6.0) + (m->kind == HH_MODEL_HESTON ? 0.3 : 0.0);
    }
    const double K = payoff->strike;
    for (int t = 0; t <= M; ++t) {
      const double ty = m->T * (double)(t > 0 ? t : 1) / (double)M;  // (date 0 has no regression)
      // variance of log S_t: sigma^2 t, or for Heston the expected integrated variance
      // int_0^t E[V_s] ds = theta t + (V0 - theta) (1 - e^{-kappa t}) / kappa  (the larger of V0 and theta as a stand-in
      // put the early dates of a model with V0 << theta back into a tenth of the interval: degree 5 and 6 lost up to 18 %
      // of their decisions)
      double var_t = m->sigma * m->sigma * ty;
      if (m->kind == HH_MODEL_HESTON) {
        const double kt = m->kappa * ty;
        const double w = fabs(kt) > 1e-8 ? -expm1(-kt) / m->kappa : ty;
        var_t = fmax(m->theta * ty + (m->V0 - m->theta) * w, 0.0);
      }
      const double med = m->S0 * exp(m->r * ty - 0.5 * var_t), reach = exp(zn * sqrt(var_t));
      double lo, hi;
      if (payoff->cp < 0) {  // put: S in (lo, K)
        hi = K;
        lo = fmin(med / reach, 0.9 * K);
      } else {               // call: S in (K, hi)
        lo = K;
        hi = fmax(med * reach, 1.1 * K);
      }
      if (!(hi > lo) || !(hi < 1e300)) { lo = payoff->cp < 0 ? 0.0 : K; hi = payoff->cp < 0 ? K : 2.0 * K + 1.0; }
      uab[2 * t] = 2.0 / (hi - lo);
      uab[2 * t + 1] = -1.0 - uab[2 * t] * lo;
    }
  }
  HH_CUDA(ctx, ctx->d_lsm_uab.ensure(sizeof(double) * uab.size()));
  HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_lsm_uab.ptr, uab.data(), sizeof(double) * uab.size(), cudaMemcpyHostToDevice, st));

  LsmPassArgs a;
  memset(&a, 0, sizeof a);
  a.ncols = ncols;
  a.z = ctx->d_cash.as<double>();
  a.tau = want_stop ? ctx->d_tau.as<int32_t>() : nullptr;
  a.partials = ctx->d_lsm_partials.as<double>();
  a.strike = payoff->strike;
  a.cp = payoff->cp;
  a.done = reinterpret_cast<unsigned int *>(static_cast<char *>(ctx->d_lsm_state.ptr) + done_off);
  a.moments = d_moments;
  a.px.world = 1;
  if (peer_mode) {
    a.px.rank = ctx->peer_rank;
    for (int qq = 0; qq < ctx->peer_world; ++qq) a.px.mail[qq] = static_cast<double *>(ctx->peer_mail[qq]);
    a.px.error = reinterpret_cast<int *>(static_cast<char *>(ctx->d_lsm_state.ptr) + done_off + 128);  // zeroed above
    a.px.timeout_cycles = (long long)(ctx->peer_timeout_s * ctx->sm_clock_hz);
  }
  const double *G = ctx->d_grid.as<double>();
  // The cash-flow vector z is read and written by EVERY pass while each date slice is read twice and then dead:
  // pin z in L2 (persisting access-policy window on the stream; misses and everything else stream through), so a
  // pass moves 16 B per column through HBM (the two date slices) instead of 32 B.
  static const int l2_mode = getenv("HH_LSM_L2") ? atoi(getenv("HH_LSM_L2")) : 1;
  bool l2_window = false;
  // Only when z fits in the carve-out (C3: 80 MB of 82.9). Beyond it a partly persisting window is slower than no window
  // at all (1.5e7 columns 3.64 / 3.43 ms, 4.5e7 10.2 / 9.5 ms, with equal DRAM traffic: ncu, profiles/r2_n_*): the
  // evict-last hint on the bulk loads of z (z_policy below) is what helps there.
  const size_t zbytes = sizeof(double) * (size_t)stride;
  if (l2_mode && ctx->l2_persist_max > 0 && ctx->l2_window_max > 0 && (l2_mode == 2 || zbytes <= ctx->l2_persist_max)) {
    size_t carve = ctx->l2_persist_max;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof attr);
      const size_t win = zbytes < ctx->l2_window_max ? zbytes : ctx->l2_window_max;
      attr.accessPolicyWindow.base_ptr = ctx->d_cash.ptr;
      attr.accessPolicyWindow.num_bytes = win;
      const double ratio = (double)carve / (double)win;
      attr.accessPolicyWindow.hitRatio = (float)(ratio < 1.0 ? ratio : 1.0);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      l2_window = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
    }
    (void)cudaGetLastError();
  }
  if (persistent) {
    LsmBackArgs bk;
    memset(&bk, 0, sizeof bk);
    char *state = static_cast<char *>(ctx->d_lsm_state.ptr);
    bk.ncols = ncols;
    bk.stride = stride;
    bk.G = G;
    bk.z = ctx->d_cash.as<double>();
    bk.tau = want_stop ? ctx->d_tau.as<int32_t>() : nullptr;
    bk.partials = ctx->d_lsm_partials.as<double>();
    bk.barrier = reinterpret_cast<unsigned int *>(state + done_off + 64);
    bk.gflag = reinterpret_cast<unsigned int *>(state + done_off + 192);
    bk.gmom = reinterpret_cast<double *>(state + done_off + 256);
    bk.moments_out = d_moments;
    bk.fits = d_fits;
    bk.logD = log(step_discount);
    bk.strike = payoff->strike;
    bk.cp = payoff->cp;
    bk.uab = ctx->d_lsm_uab.as<double>();
    bk.M = M;
    // z (the discounted cash flows, 8 B per column) is re-read on every date: evict_last keeps as much of it in the
    // 126 MB L2 as fits. Neutral at C3 (80 MB, resident anyway: 2.043 / 2.047 ms), -40 % beyond it (2e7 columns 7.82 ->
    // 4.72 ms, 4.5e7 17.2 -> 10.1 ms). The same hint on the grid slice (HH_LSM_CPOL=1) costs 30 % at C3 and gains less.
    static const int zpol = getenv("HH_LSM_ZPOL") ? atoi(getenv("HH_LSM_ZPOL")) : 1;
    static const int cpol = getenv("HH_LSM_CPOL") ? atoi(getenv("HH_LSM_CPOL")) : 0;
    bk.z_policy = zpol;
    bk.cur_policy = cpol;
    bk.px.world = 1;
    if (peer_mode) {
      bk.px = a.px;
      bk.px.world = ctx->peer_world;
      bk.px.epoch = ctx->peer_epoch;         // date d (0-based) uses epoch base + d + 1
      ctx->peer_epoch += (unsigned long long)(M > 1 ? M - 1 : 0);
    }
    const cudaError_t le = launch_backward_deg(degree, want_stop, bk, ctx->sm_count, nchunks, st, false, &grid_back);
    if (le != cudaSuccess && !peer_mode) {
      (void)cudaGetLastError();  // no cooperative launch here (e.g. a partitioned device): one launch per date instead
      persistent = false;
    } else {
      HH_CUDA(ctx, le);
    }
  }
  if (!persistent) {
  // pass(t), t = M-1 .. 0: decision at t+1 (with fit[t+1]), one-step discount, moments of date t (t >= 1)
  for (int t = M - 1; t >= 0; --t) {
    a.S_next = G + (size_t)(t + 1) * stride;
    a.S_cur = G + (size_t)t * stride;
    a.fit_next = d_fits + (t + 1);
    a.t_next = t + 1;
    a.ua = uab[2 * t];
    a.ub = uab[2 * t + 1];
    a.ua_n = uab[2 * t + 2];
    a.ub_n = uab[2 * t + 3];
    a.Dp = pow(step_discount, (double)(t + 1));
    a.rscale = pow(step_discount, -(double)t);
    a.first = (t + 1 == M);
    a.last = (t == 0);
    const bool exchange = t >= 1 && comm && comm->world > 1 && !peer_mode;  // host-callback (NCCL) form
    a.fit_out = (t >= 1 && !exchange) ? d_fits + t : nullptr;
    if (peer_mode) {
      a.px.world = t >= 1 ? ctx->peer_world : 1;  // the last pass only sums prices; the host layer reduces those
      a.px.epoch = t >= 1 ? ++ctx->peer_epoch : 0;
    }
    a.reverse = (M - 1 - t) & 1;
    HH_CUDA(ctx, launch_pass_deg(degree, a, grid_pass, st));
    if (exchange) {
      if (comm->allreduce_sum_f64(comm->user, d_moments, (size_t)nacc, (void *)st) != 0)
        return ctx->fail(HH_ERR_COMM, "allreduce callback failed at date %d", t);
      launch_fit_deg(degree, d_moments, a.rscale, d_fits + t, st);
      HH_CUDA(ctx, cudaGetLastError());
    }
  }
  }
  HH_CUDA(ctx, cudaEventRecord(ctx->ev2, st));
  if (l2_window) {  // later launches on this stream must not inherit the window (the kernels above already have it)
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.num_bytes = 0;
    (void)cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    (void)cudaGetLastError();
  }

  // results
  std::vector<double> mom((size_t)nacc);
  std::vector<LsmFit> fits((size_t)(M + 1));
  HH_CUDA(ctx, cudaMemcpyAsync(mom.data(), d_moments, sizeof(double) * nacc, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaMemcpyAsync(fits.data(), d_fits, sizeof(LsmFit) * (size_t)(M + 1), cudaMemcpyDeviceToHost, st));
  if (want_stop) {
    HH_CUDA(ctx, ctx->d_misc.ensure(sizeof(double) * (size_t)stride));
    lsm_stop_values_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(G, stride, ncols, ctx->d_tau.as<int32_t>(), payoff->strike,
                                                             payoff->cp, ctx->d_misc.as<double>());
    HH_CUDA(ctx, cudaGetLastError());
    if (int rc2 = copy_to_pageable_host(ctx, stop_idx, ctx->d_tau.ptr, sizeof(int32_t) * (size_t)ncols, st)) return rc2;
    if (int rc2 = copy_to_pageable_host(ctx, stop_val, ctx->d_misc.ptr, sizeof(double) * (size_t)ncols, st)) return rc2;
  }
  int px_error = 0;
  if (peer_mode)
    HH_CUDA(ctx, cudaMemcpyAsync(&px_error, static_cast<char *>(ctx->d_lsm_state.ptr) + done_off + 128, sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  if (l2_window) {  // the induction has finished: give the persisting lines and the carve-out back
    (void)cudaCtxResetPersistingL2Cache();
    (void)cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    (void)cudaGetLastError();
  }
  if (px_error)
    return ctx->fail(HH_ERR_PEER_TIMEOUT, "a peer's regression moments did not arrive within the in-kernel time limit (%.1f s, "
                     "hh_peer_set_timeout) or a peer gave up; every rank fails together: reconnect with hh_peer_connect",
                     ctx->peer_timeout_s);
  if (spot_paths) {
    // chunks of columns transposed on the device, then copied out (the staging buffer reuses d_terminal)
    const int nrows = M + 1;
    const int64_t chunk = 1 << 20;
    HH_CUDA(ctx, ctx->d_terminal.ensure(sizeof(double) * (size_t)(ncols < chunk ? ncols : chunk) * nrows));
    for (int64_t p0 = 0; p0 < ncols; p0 += chunk) {
      const int64_t np = ncols - p0 < chunk ? ncols - p0 : chunk;
      dim3 g((unsigned)((np + 31) / 32), (unsigned)((nrows + 31) / 32));
      lsm_transpose_kernel<<<g, dim3(32, 8), 0, st>>>(G, stride, p0, np, nrows, ctx->d_terminal.as<double>());
      HH_CUDA(ctx, cudaGetLastError());
      if (int rc2 = copy_to_pageable_host(ctx, spot_paths + (size_t)p0 * nrows, ctx->d_terminal.ptr, sizeof(double) * (size_t)np * nrows, st))
        return rc2;
      HH_CUDA(ctx, cudaStreamSynchronize(st));
    }
  }
  float ms_paths = 0.f, ms_reg = 0.f;
  HH_CUDA(ctx, cudaEventElapsedTime(&ms_paths, ctx->ev0, ctx->ev1));
  HH_CUDA(ctx, cudaEventElapsedTime(&ms_reg, ctx->ev1, ctx->ev2));

  memset(out, 0, sizeof *out);
  out->sum = mom[0];
  out->sumsq = mom[1];
  out->n = ncols;
  const double mean = out->sum / (double)ncols;
  out->price = mean;  // mean(discount^tau * value)  :132-133
  double var = ncols > 1 ? (out->sumsq - (double)ncols * mean * mean) / (double)(ncols - 1) : 0.0;
  out->std_error = sqrt((var > 0 ? var : 0) / (double)ncols);
  int64_t skipped = 0;
  for (int t = 1; t <= M - 1; ++t)
    if (fits[(size_t)t].count == 0.0) skipped++;
  out->n_dates_skipped = skipped;
  out->path_ms = ms_paths;
  out->regress_ms = ms_reg;
  out->kernel_ms = ms_paths + ms_reg;
  return HH_OK;
}

}  // namespace hh
