// hh_pathdep.cu — path-dependent payoffs on the simulation grid (include/hedgehog_mc.h
// hh_mc_path_dependent). The reference does not have these yet — its roadmap lists them as Phase 5
// (derivatives_pricing_roadmap.md:73-80) — so the trajectories are the ones the European path simulates
// (LogGBMProblem / LogHestonProblem under Euler-Maruyama, heston.jl:7-52; antithetic = the negated Wiener path,
// montecarlo.jl:258; pair-averaged payoffs, montecarlo.jl:430-432) and what is new is only what the roadmap asks for:
// running statistics of each trajectory, kept in registers while it is advanced, and payoffs that read them.
//
//   statistics per column (trajectory, or each side of an antithetic pair), over the monitoring dates:
//     S_T, A = mean S, G = exp(mean log S), max log S, min log S
//   payoffs: evaluated for all requested contracts on the SAME trajectories by transposing through shared memory
//     (thread = (contract k, path group g), as european_kernel does for a strike grid).
//
// Cost: the log-space step plus 4 FP64 per monitoring date (sum, max, min) and one exp per monitoring date when an
// arithmetic average is requested (template switch ARITH: barriers, digitals and geometric Asians never leave log space).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hh_ctx.h"
#include "hh_device.cuh"
#include "hh_fastnormal.cuh"
#include "hh_paths.cuh"

namespace hh {

constexpr int kPdThreads = 512;
constexpr int kPdStats = HH_PD_NSTATS;
constexpr int kPdStage = HH_PD_NSTATS + 1;  // staged per column: the statistics and the control's terminal spot
constexpr int kPdAcc = 3;  // sum, sumsq, nonfinite

struct PdPayoff {  // device form of hh_path_payoff: the barrier in log space (stepping kernels) and as given
  int kind;
  double strike, cp, log_barrier, amount, barrier;
};

struct PdArgs {
  int64_t n, path_offset;
  const uint64_t *seeds;
  const double *normals;
  const PdPayoff *payoffs;
  double *partials;  // [grid][npay][kPdAcc]
  double *stats;     // nullable: [kPdStats][ncols]
  int npay, kp_log2, n_steps, monitor_every, split;
  double inv_m;      // 1 / number of monitoring dates
  int cv_on;         // advance the Black-Scholes control trajectory (HH_PD_BS_CONTROL / HH_PD_VANILLA_MINUS_BS requested)
  double cv_sigma, cv_drift;  // its volatility and (r - sigma_cv^2 / 2) dt
  PathParams<double> p;
  HestonFolded f;  // folded step constants of the native-RNG Heston kernel
  PhiloxRoundKeys rk;
  uint32_t one_hi, magic_hi;
};

struct PdRunning {
  double sum_s, sum_x, max_x, min_x;
  __device__ __forceinline__ void reset() {
    sum_s = 0.0;
    sum_x = 0.0;
    max_x = -INFINITY;
    min_x = INFINITY;
  }
  // FASTEXP: table-driven exp (hh_fastnormal.cuh fast_exp_full); the parity mode keeps libm's
  template <bool ARITH, bool FASTEXP>
  __device__ __forceinline__ void monitor(double x, const double *exp_tab) {
    sum_x += x;
    max_x = fmax(max_x, x);
    min_x = fmin(min_x, x);
    if (ARITH) sum_s += FASTEXP ? fast_exp_full(exp_tab, x) : exp(x);
  }
};

__device__ __forceinline__ double pd_payoff(const PdPayoff &c, double ST, double A, double G, double mx, double mn, double Scv) {
  const double vanilla = fmax(c.cp * (ST - c.strike), 0.0);  // payoffs.jl:154-156
  switch (c.kind) {
    case HH_PD_ASIAN_ARITH: return fmax(c.cp * (A - c.strike), 0.0);
    case HH_PD_ASIAN_GEOM: return fmax(c.cp * (G - c.strike), 0.0);
    case HH_PD_ASIAN_ARITH_MINUS_GEOM: return fmax(c.cp * (A - c.strike), 0.0) - fmax(c.cp * (G - c.strike), 0.0);
    case HH_PD_UP_OUT: return mx >= c.log_barrier ? c.amount : vanilla;
    case HH_PD_UP_IN: return mx >= c.log_barrier ? vanilla : c.amount;
    case HH_PD_DOWN_OUT: return mn <= c.log_barrier ? c.amount : vanilla;
    case HH_PD_DOWN_IN: return mn <= c.log_barrier ? vanilla : c.amount;
    case HH_PD_DIGITAL_CASH: return c.cp * (ST - c.strike) > 0.0 ? c.amount : 0.0;
    case HH_PD_DIGITAL_ASSET: return c.cp * (ST - c.strike) > 0.0 ? ST : 0.0;
    case HH_PD_BS_CONTROL: return fmax(c.cp * (Scv - c.strike), 0.0);
    case HH_PD_VANILLA_MINUS_BS: return vanilla - c.amount * fmax(c.cp * (Scv - c.strike), 0.0);
    default: return vanilla;
  }
}

// Stages the five statistics of each thread's column(s) as [stat][side][thread], writes them out when requested, and
// evaluates every contract on the staged columns: thread (k, g) = (contract, path group), as european_kernel does for a
// strike grid. Called by all threads of the block (it synchronises).
template <bool ANTI, bool ARITH, int THREADS>
__device__ __forceinline__ void pd_stage_and_pay(const PdArgs &a, double *stage, int tid, int64_t i, int64_t base, double xp,
                                                 double xm, double xbp, double xbm, const PdRunning &rp, const PdRunning &rm,
                                                 const PdPayoff &mine, int k, int g, int G, double *acc) {
  constexpr int NSIDE = ANTI ? 2 : 1;
  const double ST = exp(xp);  // final_sample, montecarlo.jl:398
  stage[(0 * NSIDE + 0) * THREADS + tid] = ST;
  stage[(1 * NSIDE + 0) * THREADS + tid] = ARITH ? rp.sum_s * a.inv_m : 0.0;
  stage[(2 * NSIDE + 0) * THREADS + tid] = exp(rp.sum_x * a.inv_m);
  stage[(3 * NSIDE + 0) * THREADS + tid] = rp.max_x;
  stage[(4 * NSIDE + 0) * THREADS + tid] = rp.min_x;
  stage[(5 * NSIDE + 0) * THREADS + tid] = a.cv_on ? exp(xbp) : 0.0;
  if (ANTI) {
    const double STm = exp(xm);
    stage[(0 * NSIDE + 1) * THREADS + tid] = STm;
    stage[(1 * NSIDE + 1) * THREADS + tid] = ARITH ? rm.sum_s * a.inv_m : 0.0;
    stage[(2 * NSIDE + 1) * THREADS + tid] = exp(rm.sum_x * a.inv_m);
    stage[(3 * NSIDE + 1) * THREADS + tid] = rm.max_x;
    stage[(4 * NSIDE + 1) * THREADS + tid] = rm.min_x;
    stage[(5 * NSIDE + 1) * THREADS + tid] = a.cv_on ? exp(xbm) : 0.0;
  }
  if (a.stats && i < a.n) {  // S_T, A, G, max S, min S
    const int64_t ncols = a.n * NSIDE;
#pragma unroll
    for (int side = 0; side < NSIDE; ++side) {
      const int64_t col = i + side * a.n;
#pragma unroll
      for (int s = 0; s < kPdStats; ++s) {
        const double v = stage[(s * NSIDE + side) * THREADS + tid];
        a.stats[(int64_t)s * ncols + col] = s >= 3 ? exp(v) : v;
      }
    }
  }
  __syncthreads();
  const int64_t rem = a.n - base;
  const int nvalid = rem < THREADS ? (int)rem : THREADS;
  if (k < a.npay) {
    for (int j = g; j < nvalid; j += G) {
      const double sT = stage[(0 * NSIDE + 0) * THREADS + j];
      double pay = pd_payoff(mine, sT, stage[(1 * NSIDE + 0) * THREADS + j], stage[(2 * NSIDE + 0) * THREADS + j],
                             stage[(3 * NSIDE + 0) * THREADS + j], stage[(4 * NSIDE + 0) * THREADS + j],
                             stage[(5 * NSIDE + 0) * THREADS + j]);
      bool bad = !isfinite(sT);
      if (ANTI) {
        const double sTm = stage[(0 * NSIDE + 1) * THREADS + j];
        const double paym = pd_payoff(mine, sTm, stage[(1 * NSIDE + 1) * THREADS + j], stage[(2 * NSIDE + 1) * THREADS + j],
                                      stage[(3 * NSIDE + 1) * THREADS + j], stage[(4 * NSIDE + 1) * THREADS + j],
                                      stage[(5 * NSIDE + 1) * THREADS + j]);
        pay = 0.5 * (pay + paym);  // reduce_payoffs, montecarlo.jl:430-432
        bad = bad || !isfinite(sTm);
      }
      acc[0] += pay;
      acc[1] = fma(pay, pay, acc[1]);
      if (k == 0 && bad) acc[2] += 1.0;
    }
  }
  __syncthreads();
}

// fixed-order reduction over the path groups that share a contract (stage holds at least kPdAcc * THREADS doubles)
template <int THREADS>
__device__ __forceinline__ void pd_block_reduce(const PdArgs &a, double *stage, int tid, int G, const double *acc) {
#pragma unroll
  for (int c = 0; c < kPdAcc; ++c) stage[c * THREADS + tid] = acc[c];
  __syncthreads();
  if (tid < a.npay) {
    double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * kPdAcc;
    for (int c = 0; c < kPdAcc; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += stage[c * THREADS + (gg << a.kp_log2) + tid];
      out[c] = t;
    }
  }
}

constexpr int kPdTableBytes = kLogRepBytes + kTrigRepBytes + kExpFullBytes + kExp2Bytes;
constexpr int kPdStageDoubles = kPdStage * 2 * kPdThreads;  // also holds the kPdAcc * kPdThreads of the final reduction
constexpr int kPdSmem = ((kPdTableBytes + 15) & ~15) + kPdStageDoubles * 8;

template <bool HESTON, bool ANTI, bool PARITY, bool UKEY, bool ARITH>
__global__ void __launch_bounds__(kPdThreads, 2) pathdep_kernel(const PdArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  char *s_log = reinterpret_cast<char *>(dsm);
  char *s_trig = s_log + kLogRepBytes;
  double *s_expf = reinterpret_cast<double *>(s_trig + kTrigRepBytes);
  double *s_e2 = s_expf + kExpFullN;
  double *stage = reinterpret_cast<double *>(dsm + ((kPdTableBytes + 15) & ~15));
  const int tid = threadIdx.x;
  fill_exp_full_table(s_expf);
  if (!PARITY) {
    for (int e = tid; e < tables::kLog2Buckets * kRep; e += kPdThreads)
      reinterpret_cast<double2 *>(s_log)[e] = g_fast_tables2.log_tab[e / kRep];
    for (int e = tid; e < tables::kTrigN * kRep; e += kPdThreads)
      reinterpret_cast<double2 *>(s_trig)[e] = g_fast_tables2.trig_tab[e / kRep];
    for (int e = tid; e < tables::kExp2N; e += kPdThreads) s_e2[e] = g_fast_tables2.exp_tab[e];
  }
  __syncthreads();
  const char *log_lane = s_log + (tid & (kRep - 1)) * 16;
  const char *trig_lane = s_trig + (tid & (kRep - 1)) * 16;
  const char *exp_biased = reinterpret_cast<const char *>(s_e2) - tables::kExp2Bias * 8;
  const PathParams<double> &p = a.p;
  const bool split = a.split != 0;
  const int M = a.n_steps;
  const int every = a.monitor_every;
  constexpr int NC = HESTON ? 2 : 1;
  constexpr int NSIDE = ANTI ? 2 : 1;

  const int KP = 1 << a.kp_log2;          // contracts padded to a power of two <= 256
  const int k = tid & (KP - 1);           // my contract
  const int g = tid >> a.kp_log2;         // my path group
  const int G = kPdThreads >> a.kp_log2;  // number of path groups
  PdPayoff mine;
  mine.kind = HH_PD_VANILLA;
  mine.strike = mine.cp = mine.log_barrier = mine.amount = mine.barrier = 0.0;
  if (k < a.npay) mine = a.payoffs[k];
  double acc[kPdAcc] = {0.0, 0.0, 0.0};

  for (int64_t base = (int64_t)blockIdx.x * kPdThreads; base < a.n; base += (int64_t)gridDim.x * kPdThreads) {
    const int64_t i = base + tid;
    double xp = p.x0, xm = p.x0, vp = p.v0, vm = p.v0;
    double xbp = p.x0, xbm = p.x0;  // Black-Scholes control on the same dW1 (a.cv_on)
    PdRunning rp, rm;
    rp.reset();
    rm.reset();
    if (i < a.n) {
      uint64_t idx = (uint64_t)(a.path_offset + i);
      PhiloxRoundKeys rk_own;
      if (!UKEY && !PARITY) {
        rk_own = philox_round_keys(a.seeds[i]);
        idx = 0;
      }
      const double *z = PARITY ? a.normals + (size_t)i * (size_t)M * NC : nullptr;
      int due = every;  // steps until the next monitoring date
      if (HESTON) {
#pragma unroll 1
        for (int n = 0; n < M; ++n) {
          double z1, z2;
          if (PARITY) {
            z1 = z[2 * n];
            z2 = z[2 * n + 1];
          } else {
            const u32x4 w = philox4x32_10_rk((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)n, 0u, UKEY ? a.rk : rk_own);
            fast_normal_pair_v2(log_lane, exp_biased, trig_lane, w.x, w.y, w.z, w.w, a.one_hi, a.magic_hi, z1, z2);
          }
          const double dW1 = fma(p.a12, z2, p.a11 * z1);
          const double dW2 = fma(p.a22, z2, p.a21 * z1);
          heston_em_step<double>(p, split, xp, vp, dW1, dW2);
          if (ANTI) heston_em_step<double>(p, split, xm, vm, -dW1, -dW2);  // NoiseGrid(t, -W), montecarlo.jl:258
          if (a.cv_on) {
            xbp = fma(a.cv_sigma, dW1, xbp + a.cv_drift);
            if (ANTI) xbm = fma(a.cv_sigma, -dW1, xbm + a.cv_drift);
          }
          if (--due == 0) {
            due = every;
            rp.monitor<ARITH, !PARITY>(xp, s_expf);
            if (ANTI) rm.monitor<ARITH, !PARITY>(xm, s_expf);
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
            const u32x4 w = philox4x32_10_rk((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(n >> 1), 0u, UKEY ? a.rk : rk_own);
            fast_normal_pair_v2(log_lane, exp_biased, trig_lane, w.x, w.y, w.z, w.w, a.one_hi, a.magic_hi, za, zb);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (n + h < M) {
              const double dW = p.sqdt * (h ? zb : za);
              gbm_em_step<double>(p, xp, dW);
              if (ANTI) gbm_em_step<double>(p, xm, -dW);
              if (--due == 0) {
                due = every;
                rp.monitor<ARITH, !PARITY>(xp, s_expf);
                if (ANTI) rm.monitor<ARITH, !PARITY>(xm, s_expf);
              }
            }
          }
        }
      }
    }
    pd_stage_and_pay<ANTI, ARITH, kPdThreads>(a, stage, tid, i, base, xp, xm, xbp, xbm, rp, rm, mine, k, g, G, acc);
  }
  pd_block_reduce<kPdThreads>(a, stage, tid, G, acc);
}

// ---- LogHestonProblem with the in-kernel RNG: the step of the headline kernel (heston_fast2_kernel, hh_european.cu) ----
// Box-Muller radius folded into the diffusion's square root, rotation folded into the Brownian factor through the phase
// table, integer clamps, branch-free sqrt; the drift r dt stays inside the step so that the state is log S on every
// monitoring date. Same trajectories as pathdep_kernel<true, ...> and the oracle up to rounding.
// Shared memory: [staging | log table x8 | phase table x8 | T_j | exponent table]; THREADS = 1024 for plain runs with the
// uniform key (<= 64 registers), 512 otherwise (antithetic pairs / per-trajectory keys need the registers).
template <bool ANTI, int THREADS>
__host__ __device__ constexpr int pd_fast_stage_bytes() { return kPdStage * (ANTI ? 2 : 1) * THREADS * 8; }
template <bool ANTI, int THREADS>
__host__ __device__ constexpr int pd_fast_smem() {
  return pd_fast_stage_bytes<ANTI, THREADS>() + kLogRepBytes + kPhaseRepBytes + kExpFullBytes + kExp2Bytes;
}

template <bool ANTI, bool SPLIT, bool UKEY, bool ARITH, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) pathdep_heston_fast_kernel(const PdArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  double *stage = reinterpret_cast<double *>(dsm);
  char *s_log = reinterpret_cast<char *>(dsm) + pd_fast_stage_bytes<ANTI, THREADS>();
  char *s_phase = s_log + kLogRepBytes;
  double *s_expf = reinterpret_cast<double *>(s_phase + kPhaseRepBytes);
  double *s_e2 = s_expf + kExpFullN;
  const int tid = threadIdx.x;
  for (int e = tid; e < tables::kLog2Buckets * kRep; e += THREADS)
    reinterpret_cast<double2 *>(s_log)[e] = g_fast_tables2.log_tab[e / kRep];
  for (int e = tid; e < tables::kTrigN * kRep; e += THREADS) {
    // phase table: {P1, Q1}, {P2, Q2} per table angle (see heston_fast2_kernel)
    const int j = e / kRep, q = e % kRep;
    const double2 cs = g_fast_tables2.trig_tab[j];
    double2 *dst = reinterpret_cast<double2 *>(s_phase) + (size_t)j * 2 * kRep + q;
    dst[0] = make_double2(fma(a.p.a12, cs.y, a.p.a11 * cs.x), fma(a.p.a12, cs.x, -(a.p.a11 * cs.y)));
    dst[kRep] = make_double2(fma(a.f.b22, cs.y, a.f.b21 * cs.x), fma(a.f.b22, cs.x, -(a.f.b21 * cs.y)));
  }
  for (int e = tid; e < tables::kExp2N; e += THREADS) s_e2[e] = g_fast_tables2.exp_tab[e];
  fill_exp_full_table(s_expf);
  __syncthreads();
  const char *log_lane = s_log + (tid & (kRep - 1)) * 16;
  const char *phase_lane = s_phase + (tid & (kRep - 1)) * 16;
  const char *exp_biased = reinterpret_cast<const char *>(s_e2) - tables::kExp2Bias * 8;

  const int KP = 1 << a.kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> a.kp_log2;
  const int G = THREADS >> a.kp_log2;
  PdPayoff mine;
  mine.kind = HH_PD_VANILLA;
  mine.strike = mine.cp = mine.log_barrier = mine.amount = mine.barrier = 0.0;
  if (k < a.npay) mine = a.payoffs[k];
  double acc[kPdAcc] = {0.0, 0.0, 0.0};
  const int M = a.n_steps;
  const int every = a.monitor_every;

  for (int64_t base = (int64_t)blockIdx.x * THREADS; base < a.n; base += (int64_t)gridDim.x * THREADS) {
    const int64_t i = base + tid;
    const int64_t ic = i < a.n ? i : a.n - 1;  // tail lanes repeat the last trajectory; never accumulated
    double xp = a.p.x0, xm = a.p.x0, vp = a.p.v0, vm = a.p.v0;
    double xbp = a.p.x0, xbm = a.p.x0;  // Black-Scholes control on the same dW1 (a.cv_on)
    PdRunning rp, rm;
    rp.reset();
    rm.reset();
    uint32_t c0 = 0u, c1 = 0u;
    PhiloxRoundKeys rk_own;
    if (UKEY) {
      const uint64_t idx = (uint64_t)(a.path_offset + ic);
      c0 = (uint32_t)idx;
      c1 = (uint32_t)(idx >> 32);
    } else {
      rk_own = philox_round_keys(a.seeds[ic]);
    }
    int due = every;
#pragma unroll 1
    for (int n = 0; n < M; ++n) {
      const u32x4 w = philox4x32_10_rk(c0, c1, (uint32_t)n, 0u, UKEY ? a.rk : rk_own);
      const double R2 = fast_neg2log_v2(log_lane, exp_biased, w.x, w.y, a.one_hi);
      double sn, cs;
      const uint32_t poff = fast_angle_v2(w.z, w.w, a.magic_hi, sn, cs);
      const double2 pq1 = *reinterpret_cast<const double2 *>(phase_lane + poff);
      const double2 pq2 = *reinterpret_cast<const double2 *>(phase_lane + poff + kRep * 16);
      const double cc1 = fma(pq1.x, cs, pq1.y * sn);
      const double cc2 = fma(pq2.x, cs, pq2.y * sn);
      {
        const double vplus = max0_hi(vp);
        const double K1 = fma(a.f.neg_half_dt, vplus, xp + a.f.rdt);
        const double K2 = fma(a.f.neg_kdt, vplus, vp + a.f.ktdt);
        const double sr = fast_sqrt_pos5(max_tiny_hi((SPLIT ? K2 : vplus) * R2));
        xp = fma(sr, cc1, K1);
        vp = fma(sr, cc2, K2);
      }
      if (ANTI) {
        const double vplus = max0_hi(vm);
        const double K1 = fma(a.f.neg_half_dt, vplus, xm + a.f.rdt);
        const double K2 = fma(a.f.neg_kdt, vplus, vm + a.f.ktdt);
        const double sr = fast_sqrt_pos5(max_tiny_hi((SPLIT ? K2 : vplus) * R2));
        xm = fma(-sr, cc1, K1);
        vm = fma(-sr, cc2, K2);
      }
      if (a.cv_on) {  // dW1 = rad cc1 with the Box-Muller radius, which the Heston step itself never forms
        const double dW1 = fast_sqrt_pos5(max_tiny_hi(R2)) * cc1;
        xbp = fma(a.cv_sigma, dW1, xbp + a.cv_drift);
        if (ANTI) xbm = fma(a.cv_sigma, -dW1, xbm + a.cv_drift);
      }
      if (--due == 0) {
        due = every;
        rp.monitor<ARITH, true>(xp, s_expf);
        if (ANTI) rm.monitor<ARITH, true>(xm, s_expf);
      }
    }
    pd_stage_and_pay<ANTI, ARITH, THREADS>(a, stage, tid, i, base, xp, xm, xbp, xbm, rp, rm, mine, k, g, G, acc);
  }
  pd_block_reduce<THREADS>(a, stage, tid, G, acc);
}

// Contracts evaluated from a statistics buffer [HH_PD_NSTATS][n] in S-space (the Broadie-Kaya path kernel writes it:
// exact transitions between the monitoring dates, hh_bk.cu). Same (contract, path group) transpose, no antithetic side.
__global__ void __launch_bounds__(256) pathdep_from_stats_kernel(const double *__restrict__ stats, int64_t n,
                                                                 const PdPayoff *__restrict__ payoffs, int npay, int kp_log2,
                                                                 double *partials) {
  __shared__ double red[kPdAcc * 256];
  const int tid = threadIdx.x;
  const int KP = 1 << kp_log2;
  const int k = tid & (KP - 1);
  const int g = tid >> kp_log2;
  const int G = 256 >> kp_log2;
  PdPayoff mine;
  mine.kind = HH_PD_VANILLA;
  mine.strike = mine.cp = mine.log_barrier = mine.amount = mine.barrier = 0.0;
  if (k < npay) mine = payoffs[k];
  mine.log_barrier = mine.barrier;  // pd_payoff compares its max / min arguments with log_barrier: spots against the barrier here
  double acc[kPdAcc] = {0.0, 0.0, 0.0};
  if (k < npay) {
    for (int64_t j = (int64_t)blockIdx.x * G + g; j < n; j += (int64_t)gridDim.x * G) {
      const double ST = stats[j];
      const double pay = pd_payoff(mine, ST, stats[n + j], stats[2 * n + j], stats[3 * n + j], stats[4 * n + j], 0.0);
      acc[0] += pay;
      acc[1] = fma(pay, pay, acc[1]);
      if (k == 0 && !isfinite(ST)) acc[2] += 1.0;
    }
  }
#pragma unroll
  for (int c = 0; c < kPdAcc; ++c) red[c * 256 + tid] = acc[c];
  __syncthreads();
  if (tid < npay) {
    double *out = partials + ((size_t)blockIdx.x * npay + tid) * kPdAcc;
    for (int c = 0; c < kPdAcc; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += red[c * 256 + (gg << kp_log2) + tid];
      out[c] = t;
    }
  }
}

// Sum the per-block partials in a fixed order: one block per contract.
__global__ void __launch_bounds__(256) pathdep_finalize_kernel(const double *partials, int nblocks, int npay, double *out) {
  __shared__ double scratch[256 / 32];
  const int k = blockIdx.x;
  for (int c = 0; c < kPdAcc; ++c) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) v += partials[((size_t)b * npay + k) * kPdAcc + c];
    const double t = block_sum<256>(v, scratch);
    if (threadIdx.x == 0) out[(size_t)k * kPdAcc + c] = t;
  }
}

template <bool H, bool A, bool P, bool U, bool AR>
static cudaError_t pd_launch_one(const PdArgs &a, int grid, cudaStream_t st) {
  static PerDeviceOnce opted;  // per instantiation and device
  if (cudaError_t e = smem_opt_in(opted, pathdep_kernel<H, A, P, U, AR>, kPdSmem); e != cudaSuccess) return e;
  pathdep_kernel<H, A, P, U, AR><<<grid, kPdThreads, kPdSmem, st>>>(a);
  return cudaGetLastError();
}

template <bool H, bool A, bool AR>
static cudaError_t pd_launch_rng(const PdArgs &a, bool parity, bool ukey, int grid, cudaStream_t st) {
  if (parity) return pd_launch_one<H, A, true, true, AR>(a, grid, st);
  return ukey ? pd_launch_one<H, A, false, true, AR>(a, grid, st) : pd_launch_one<H, A, false, false, AR>(a, grid, st);
}

template <bool H>
static cudaError_t pd_launch_model(const PdArgs &a, bool anti, bool arith, bool parity, bool ukey, int grid, cudaStream_t st) {
  if (anti) return arith ? pd_launch_rng<H, true, true>(a, parity, ukey, grid, st) : pd_launch_rng<H, true, false>(a, parity, ukey, grid, st);
  return arith ? pd_launch_rng<H, false, true>(a, parity, ukey, grid, st) : pd_launch_rng<H, false, false>(a, parity, ukey, grid, st);
}

template <bool A, bool SP, bool U, bool AR, int T>
static cudaError_t pd_fast_one(const PdArgs &a, int sm_count, cudaStream_t st, int *grid_out) {
  constexpr int smem = pd_fast_smem<A, T>();
  static PerDeviceOnce opted;  // per instantiation and device
  if (cudaError_t e = smem_opt_in(opted, pathdep_heston_fast_kernel<A, SP, U, AR, T>, smem); e != cudaSuccess) return e;
  const int64_t batches = (a.n + T - 1) / T;
  const int grid = (int)(batches < (int64_t)sm_count ? batches : (int64_t)sm_count);  // one block per SM
  *grid_out = grid;
  if (!a.partials) return cudaSuccess;  // sizing query
  pathdep_heston_fast_kernel<A, SP, U, AR, T><<<grid, T, smem, st>>>(a);
  return cudaGetLastError();
}

template <bool A, bool SP, bool AR>
static cudaError_t pd_fast_key(const PdArgs &a, bool ukey, int sm_count, cudaStream_t st, int *grid_out) {
  if (!ukey) return pd_fast_one<A, SP, false, AR, 512>(a, sm_count, st, grid_out);
  // uniform key: 1024 threads once the job fills them, as the headline kernel does (HH_PD_THREADS=512 forces the small block)
  static const int force_threads = getenv("HH_PD_THREADS") ? atoi(getenv("HH_PD_THREADS")) : 0;
  return a.n >= (int64_t)sm_count * 1024 && force_threads != 512 ? pd_fast_one<A, SP, true, AR, 1024>(a, sm_count, st, grid_out)
                                                                 : pd_fast_one<A, SP, true, AR, 512>(a, sm_count, st, grid_out);
}

static cudaError_t pd_fast(const PdArgs &a, bool anti, bool arith, bool ukey, int sm_count, cudaStream_t st, int *grid_out) {
  const bool split = a.split != 0;
#define HH_PD_FAST(A, SP)                                                                        \
  (arith ? pd_fast_key<A, SP, true>(a, ukey, sm_count, st, grid_out) : pd_fast_key<A, SP, false>(a, ukey, sm_count, st, grid_out))
  if (anti) return split ? HH_PD_FAST(true, true) : HH_PD_FAST(true, false);
  return split ? HH_PD_FAST(false, true) : HH_PD_FAST(false, false);
#undef HH_PD_FAST
}

// sigma_cv^2 of the Black-Scholes control: the mean of E[V_t] = theta + (V0 - theta) e^(-kappa t) over [0, T]
static double bs_control_variance(const hh_model *m) {
  const double kT = m->kappa * m->T;
  const double w = fabs(kT) > 1e-8 ? -expm1(-kT) / kT : 1.0 - 0.5 * kT;
  const double v = m->theta + (m->V0 - m->theta) * w;
  return v > 1e-12 ? v : 1e-12;
}

static void fill_results(hh_result *results, const std::vector<double> &fin, int npay, int64_t N, double discount, float ms,
                         int64_t n_fallback) {
  for (int k = 0; k < npay; ++k) {
    hh_result *r = &results[k];
    memset(r, 0, sizeof *r);
    r->sum = fin[(size_t)k * kPdAcc + 0];
    r->sumsq = fin[(size_t)k * kPdAcc + 1];
    r->n = N;
    const double mean = r->sum / (double)N;
    r->price = discount * mean;  // montecarlo.jl:489-490
    double var = N > 1 ? (r->sumsq - (double)N * mean * mean) / (double)(N - 1) : 0.0;
    if (var < 0) var = 0;
    r->std_error = discount * sqrt(var / (double)N);
    r->n_nonfinite = (int64_t)fin[2];
    r->n_fallback = n_fallback;
    r->kernel_ms = ms;
  }
}

int path_dependent(hh_ctx *ctx, const hh_model *m, const hh_sim *s, int monitor_every, const hh_path_payoff *payoffs,
                   int npay, double discount, hh_result *results, double *path_stats, size_t path_stats_len) {
  int rc = validate_model_sim(ctx, m, s);
  if (rc) return rc;
  if (ctx->pend.active) return ctx->fail(HH_ERR_ARG, "a European launch is pending on this context: collect it first");
  if (!payoffs || !results) return ctx->fail(HH_ERR_ARG, "payoffs/results is NULL");
  if (s->rng_mode == HH_RNG_PHILOX_64) return ctx->fail(HH_ERR_UNSUPPORTED, "HH_RNG_PHILOX_64 covers European pricing only");
  if (npay < 1 || npay > 256) return ctx->fail(HH_ERR_ARG, "npayoffs must be in [1, 256] (got %d)", npay);
  const bool heston = m->kind == HH_MODEL_HESTON;
  const bool bk = heston && s->scheme == HH_SCHEME_HESTON_BK;  // exact transitions between the dates (n_steps of them)
  if (!(s->scheme == HH_SCHEME_EM || (!heston && s->scheme == HH_SCHEME_EXACT_STEPS) || bk))
    return ctx->fail(HH_ERR_UNSUPPORTED, "path-dependent payoffs run on the stepping schemes (EulerMaruyama; BlackScholesExact "
                     "increments for LognormalDynamics; HestonBroadieKaya transitions); scheme %d saves no intermediate dates",
                     s->scheme);
  if (s->precision != HH_PREC_F64) return ctx->fail(HH_ERR_UNSUPPORTED, "path-dependent payoffs are computed in binary64");
  const int M = s->n_steps;
  if (monitor_every < 1 || M % monitor_every != 0)
    return ctx->fail(HH_ERR_ARG, "n_steps (%d) must be a positive multiple of monitor_every (%d)", M, monitor_every);
  bool arith = path_stats != nullptr, cv = false;
  std::vector<PdPayoff> host((size_t)npay);
  for (int k = 0; k < npay; ++k) {
    const hh_path_payoff &q = payoffs[k];
    if (q.kind < 0 || q.kind >= HH_PD_NKINDS) return ctx->fail(HH_ERR_ARG, "payoff %d: unknown kind %d", k, q.kind);
    if (!(q.cp == 1.0 || q.cp == -1.0)) return ctx->fail(HH_ERR_ARG, "payoff %d: cp must be +1 or -1", k);
    const bool barrier = q.kind >= HH_PD_UP_OUT && q.kind <= HH_PD_DOWN_IN;
    if (barrier && !(q.barrier > 0.0)) return ctx->fail(HH_ERR_ARG, "payoff %d: the barrier must be positive", k);
    host[k].kind = q.kind;
    host[k].strike = q.strike;
    host[k].cp = q.cp;
    host[k].log_barrier = barrier ? log(q.barrier) : 0.0;
    host[k].amount = q.amount;
    host[k].barrier = q.barrier;
    arith = arith || q.kind == HH_PD_ASIAN_ARITH || q.kind == HH_PD_ASIAN_ARITH_MINUS_GEOM;
    if (q.kind == HH_PD_BS_CONTROL || q.kind == HH_PD_VANILLA_MINUS_BS) {
      if (!heston || s->scheme != HH_SCHEME_EM)
        return ctx->fail(HH_ERR_ARG, "payoff %d: the Black-Scholes control variate runs next to HestonDynamics + EulerMaruyama", k);
      cv = true;
    }
  }
  const int64_t N = s->n_paths;
  const bool anti = s->vr == HH_VR_ANTITHETIC;
  const int64_t ncols = anti ? 2 * N : N;
  if (path_stats && path_stats_len < (size_t)ncols * kPdStats)
    return ctx->fail(HH_ERR_ARG, "path_stats buffer too short: %zu < %zu", path_stats_len, (size_t)ncols * kPdStats);
  const bool parity = s->rng_mode == HH_RNG_NORMALS;

  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (bk) {
    // The Broadie-Kaya path kernel writes the statistics of every trajectory (S-space); the contracts are evaluated from
    // that buffer. d_misc holds the statistics (d_terminal and d_grid belong to the path kernel).
    HH_CUDA(ctx, ctx->d_misc.ensure(sizeof(double) * (size_t)N * kPdStats));
    HH_CUDA(ctx, ctx->d_payoffs.ensure(sizeof(PdPayoff) * (size_t)npay));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_payoffs.ptr, host.data(), sizeof(PdPayoff) * (size_t)npay, cudaMemcpyHostToDevice, st));
    rc = bk_path_stats_launch(ctx, m, s, monitor_every, ctx->d_misc.as<double>());  // records ev0
    if (rc) return rc;
    int kp_log2 = 0;
    while ((1 << kp_log2) < npay) kp_log2++;
    const int G = 256 >> kp_log2;
    const int64_t want = (N + G - 1) / G;
    const int grid = (int)(want < (int64_t)ctx->sm_count * 8 ? want : (int64_t)ctx->sm_count * 8);
    HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)grid * npay * kPdAcc));
    HH_CUDA(ctx, ctx->d_final.ensure(sizeof(double) * (size_t)npay * kPdAcc));
    pathdep_from_stats_kernel<<<grid, 256, 0, st>>>(ctx->d_misc.as<double>(), N, ctx->d_payoffs.as<PdPayoff>(), npay, kp_log2,
                                                   ctx->d_partials.as<double>());
    HH_CUDA(ctx, cudaGetLastError());
    pathdep_finalize_kernel<<<npay, 256, 0, st>>>(ctx->d_partials.as<double>(), grid, npay, ctx->d_final.as<double>());
    HH_CUDA(ctx, cudaGetLastError());
    HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    std::vector<double> fin((size_t)npay * kPdAcc);
    if (path_stats) {
      rc = copy_to_pageable_host(ctx, path_stats, ctx->d_misc.ptr, sizeof(double) * (size_t)N * kPdStats, st);
      if (rc) return rc;
    }
    HH_CUDA(ctx, cudaMemcpyAsync(fin.data(), ctx->d_final.ptr, sizeof(double) * fin.size(), cudaMemcpyDeviceToHost, st));
    HH_CUDA(ctx, cudaStreamSynchronize(st));
    int64_t n_fallback = 0;
    rc = bk_read_counters(ctx, &n_fallback);
    if (rc) return rc;
    float ms = 0.f;
    HH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    fill_results(results, fin, npay, N, discount, ms, n_fallback);
    return HH_OK;
  }
  HH_CUDA(ctx, upload_fast_tables2(ctx->device, st));

  PdArgs a;
  memset(&a, 0, sizeof a);
  a.n = N;
  a.path_offset = s->path_offset;
  a.npay = npay;
  while ((1 << a.kp_log2) < npay) a.kp_log2++;
  a.n_steps = M;
  a.monitor_every = monitor_every;
  a.inv_m = 1.0 / (double)(M / monitor_every);
  a.split = (m->flags & HH_FLAG_SPLIT_STEP) != 0;
  a.cv_on = cv;
  if (cv) {
    a.cv_sigma = sqrt(bs_control_variance(m));
    a.cv_drift = (m->r - 0.5 * a.cv_sigma * a.cv_sigma) * (m->T / s->n_steps);
  }
  a.rk = philox_round_keys(s->base_seed);
  a.one_hi = 0x3FF00000u;
  a.magic_hi = 0x43300000u;
  PathParams<double> &p = a.p;
  const double dt = m->T / M;  // montecarlo.jl:349
  const double sqdt = sqrt(dt);
  p.dt = dt;
  p.sqdt = sqdt;
  p.x0 = log(m->S0);
  p.S0 = m->S0;
  p.r = m->r;
  if (heston) {
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
  } else {
    p.sigma = m->sigma;
    p.dt_drift = dt * (m->r - 0.5 * (m->sigma * m->sigma));
  }

  HH_CUDA(ctx, ctx->d_payoffs.ensure(sizeof(PdPayoff) * (size_t)npay));
  HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_payoffs.ptr, host.data(), sizeof(PdPayoff) * (size_t)npay, cudaMemcpyHostToDevice, st));
  a.payoffs = ctx->d_payoffs.as<PdPayoff>();
  if (parity) {
    const size_t bytes = sizeof(double) * (size_t)N * (size_t)M * (heston ? 2 : 1);
    HH_CUDA(ctx, ctx->d_normals.ensure(bytes));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_normals.ptr, s->normals, bytes, cudaMemcpyHostToDevice, st));
    a.normals = ctx->d_normals.as<double>();
  } else if (s->seeds) {
    const size_t bytes = sizeof(uint64_t) * (size_t)N;
    HH_CUDA(ctx, ctx->d_seeds.ensure(bytes));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_seeds.ptr, s->seeds, bytes, cudaMemcpyHostToDevice, st));
    a.seeds = ctx->d_seeds.as<uint64_t>();
  }
  if (path_stats) {
    HH_CUDA(ctx, ctx->d_terminal.ensure(sizeof(double) * (size_t)ncols * kPdStats));
    a.stats = ctx->d_terminal.as<double>();
  }
  // Heston with the in-kernel RNG takes the specialised kernel (HH_PD_GENERIC=1 forces the generic one); everything else
  // the generic kernel: one resident wave (2 blocks of 512 threads per SM), grid-stride over batches
  static const bool force_generic = getenv("HH_PD_GENERIC") && atoi(getenv("HH_PD_GENERIC")) != 0;
  const bool fast = heston && !parity && !force_generic;
  const bool ukey = a.seeds == nullptr;
  int grid = 0;
  if (fast) {
    HH_CUDA(ctx, pd_fast(a, anti, arith, ukey, ctx->sm_count, st, &grid));  // partials == NULL: sizing only
  } else {
    const int64_t batches = (N + kPdThreads - 1) / kPdThreads;
    grid = (int)(batches < (int64_t)ctx->sm_count * 2 ? batches : (int64_t)ctx->sm_count * 2);
  }
  HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)grid * npay * kPdAcc));
  HH_CUDA(ctx, ctx->d_final.ensure(sizeof(double) * (size_t)npay * kPdAcc));
  a.partials = ctx->d_partials.as<double>();

  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  if (fast)
    HH_CUDA(ctx, pd_fast(a, anti, arith, ukey, ctx->sm_count, st, &grid));
  else
    HH_CUDA(ctx, heston ? pd_launch_model<true>(a, anti, arith, parity, ukey, grid, st)
                        : pd_launch_model<false>(a, anti, arith, parity, ukey, grid, st));
  pathdep_finalize_kernel<<<npay, 256, 0, st>>>(a.partials, grid, npay, ctx->d_final.as<double>());
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));

  std::vector<double> fin((size_t)npay * kPdAcc);
  if (path_stats) {
    rc = copy_to_pageable_host(ctx, path_stats, ctx->d_terminal.ptr, sizeof(double) * (size_t)ncols * kPdStats, st);
    if (rc) return rc;
  }
  HH_CUDA(ctx, cudaMemcpyAsync(fin.data(), ctx->d_final.ptr, sizeof(double) * fin.size(), cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  HH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  fill_results(results, fin, npay, N, discount, ms, 0);
  return HH_OK;
}

}  // namespace hh
