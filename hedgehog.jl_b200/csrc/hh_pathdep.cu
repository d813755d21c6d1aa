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
#include <cstring>
#include <vector>

#include "hh_ctx.h"
#include "hh_device.cuh"
#include "hh_fastnormal.cuh"
#include "hh_paths.cuh"

namespace hh {

constexpr int kPdThreads = 512;
constexpr int kPdStats = HH_PD_NSTATS;
constexpr int kPdAcc = 3;  // sum, sumsq, nonfinite

struct PdPayoff {  // device form of hh_path_payoff: the barrier in log space
  int kind;
  double strike, cp, log_barrier, amount;
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
  PathParams<double> p;
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
  template <bool ARITH>
  __device__ __forceinline__ void monitor(double x) {
    sum_x += x;
    max_x = fmax(max_x, x);
    min_x = fmin(min_x, x);
    if (ARITH) sum_s += exp(x);
  }
};

__device__ __forceinline__ double pd_payoff(const PdPayoff &c, double ST, double A, double G, double mx, double mn) {
  const double vanilla = fmax(c.cp * (ST - c.strike), 0.0);  // payoffs.jl:154-156
  switch (c.kind) {
    case HH_PD_ASIAN_ARITH: return fmax(c.cp * (A - c.strike), 0.0);
    case HH_PD_ASIAN_GEOM: return fmax(c.cp * (G - c.strike), 0.0);
    case HH_PD_UP_OUT: return mx >= c.log_barrier ? c.amount : vanilla;
    case HH_PD_UP_IN: return mx >= c.log_barrier ? vanilla : c.amount;
    case HH_PD_DOWN_OUT: return mn <= c.log_barrier ? c.amount : vanilla;
    case HH_PD_DOWN_IN: return mn <= c.log_barrier ? vanilla : c.amount;
    case HH_PD_DIGITAL_CASH: return c.cp * (ST - c.strike) > 0.0 ? c.amount : 0.0;
    case HH_PD_DIGITAL_ASSET: return c.cp * (ST - c.strike) > 0.0 ? ST : 0.0;
    default: return vanilla;
  }
}

constexpr int kPdTableBytes = kLogRepBytes + kTrigRepBytes + kExp2Bytes;
constexpr int kPdStageDoubles = kPdStats * 2 * kPdThreads;  // also holds the kPdAcc * kPdThreads of the final reduction
constexpr int kPdSmem = ((kPdTableBytes + 15) & ~15) + kPdStageDoubles * 8;

template <bool HESTON, bool ANTI, bool PARITY, bool UKEY, bool ARITH>
__global__ void __launch_bounds__(kPdThreads, 2) pathdep_kernel(const PdArgs a) {
  extern __shared__ __align__(16) unsigned char dsm[];
  char *s_log = reinterpret_cast<char *>(dsm);
  char *s_trig = s_log + kLogRepBytes;
  double *s_e2 = reinterpret_cast<double *>(s_trig + kTrigRepBytes);
  double *stage = reinterpret_cast<double *>(dsm + ((kPdTableBytes + 15) & ~15));
  const int tid = threadIdx.x;
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
  mine.strike = mine.cp = mine.log_barrier = mine.amount = 0.0;
  if (k < a.npay) mine = a.payoffs[k];
  double acc[kPdAcc] = {0.0, 0.0, 0.0};

  for (int64_t base = (int64_t)blockIdx.x * kPdThreads; base < a.n; base += (int64_t)gridDim.x * kPdThreads) {
    const int64_t i = base + tid;
    double xp = p.x0, xm = p.x0, vp = p.v0, vm = p.v0;
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
          if (--due == 0) {
            due = every;
            rp.monitor<ARITH>(xp);
            if (ANTI) rm.monitor<ARITH>(xm);
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
                rp.monitor<ARITH>(xp);
                if (ANTI) rm.monitor<ARITH>(xm);
              }
            }
          }
        }
      }
    }
    // stage the statistics of my column(s): [stat][side][thread]
    {
      const double ST = exp(xp);  // final_sample, montecarlo.jl:398
      stage[(0 * 2 + 0) * kPdThreads + tid] = ST;
      stage[(1 * 2 + 0) * kPdThreads + tid] = ARITH ? rp.sum_s * a.inv_m : 0.0;
      stage[(2 * 2 + 0) * kPdThreads + tid] = exp(rp.sum_x * a.inv_m);
      stage[(3 * 2 + 0) * kPdThreads + tid] = rp.max_x;
      stage[(4 * 2 + 0) * kPdThreads + tid] = rp.min_x;
      if (ANTI) {
        const double STm = exp(xm);
        stage[(0 * 2 + 1) * kPdThreads + tid] = STm;
        stage[(1 * 2 + 1) * kPdThreads + tid] = ARITH ? rm.sum_s * a.inv_m : 0.0;
        stage[(2 * 2 + 1) * kPdThreads + tid] = exp(rm.sum_x * a.inv_m);
        stage[(3 * 2 + 1) * kPdThreads + tid] = rm.max_x;
        stage[(4 * 2 + 1) * kPdThreads + tid] = rm.min_x;
      }
      if (a.stats && i < a.n) {  // S_T, A, G, max S, min S
        const int64_t ncols = a.n * NSIDE;
#pragma unroll
        for (int side = 0; side < NSIDE; ++side) {
          const int64_t col = i + side * a.n;
#pragma unroll
          for (int s = 0; s < kPdStats; ++s) {
            const double v = stage[(s * 2 + side) * kPdThreads + tid];
            a.stats[(int64_t)s * ncols + col] = s >= 3 ? exp(v) : v;
          }
        }
      }
    }
    __syncthreads();
    const int64_t rem = a.n - base;
    const int nvalid = rem < kPdThreads ? (int)rem : kPdThreads;
    if (k < a.npay) {
      for (int j = g; j < nvalid; j += G) {
        const double ST = stage[(0 * 2 + 0) * kPdThreads + j];
        double pay = pd_payoff(mine, ST, stage[(1 * 2 + 0) * kPdThreads + j], stage[(2 * 2 + 0) * kPdThreads + j],
                               stage[(3 * 2 + 0) * kPdThreads + j], stage[(4 * 2 + 0) * kPdThreads + j]);
        bool bad = !isfinite(ST);
        if (ANTI) {
          const double STm = stage[(0 * 2 + 1) * kPdThreads + j];
          const double paym = pd_payoff(mine, STm, stage[(1 * 2 + 1) * kPdThreads + j], stage[(2 * 2 + 1) * kPdThreads + j],
                                        stage[(3 * 2 + 1) * kPdThreads + j], stage[(4 * 2 + 1) * kPdThreads + j]);
          pay = 0.5 * (pay + paym);  // reduce_payoffs, montecarlo.jl:430-432
          bad = bad || !isfinite(STm);
        }
        acc[0] += pay;
        acc[1] = fma(pay, pay, acc[1]);
        if (k == 0 && bad) acc[2] += 1.0;
      }
    }
    __syncthreads();
  }

  // fixed-order reduction over the path groups that share a contract
#pragma unroll
  for (int c = 0; c < kPdAcc; ++c) stage[c * kPdThreads + tid] = acc[c];
  __syncthreads();
  if (tid < a.npay) {
    double *out = a.partials + ((size_t)blockIdx.x * a.npay + tid) * kPdAcc;
    for (int c = 0; c < kPdAcc; ++c) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += stage[c * kPdThreads + (gg << a.kp_log2) + tid];
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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pathdep_kernel<H, A, P, U, AR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
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

int path_dependent(hh_ctx *ctx, const hh_model *m, const hh_sim *s, int monitor_every, const hh_path_payoff *payoffs,
                   int npay, double discount, hh_result *results, double *path_stats, size_t path_stats_len) {
  int rc = validate_model_sim(ctx, m, s);
  if (rc) return rc;
  if (ctx->pend.active) return ctx->fail(HH_ERR_ARG, "a European launch is pending on this context: collect it first");
  if (!payoffs || !results) return ctx->fail(HH_ERR_ARG, "payoffs/results is NULL");
  if (npay < 1 || npay > 256) return ctx->fail(HH_ERR_ARG, "npayoffs must be in [1, 256] (got %d)", npay);
  const bool heston = m->kind == HH_MODEL_HESTON;
  if (!(s->scheme == HH_SCHEME_EM || (!heston && s->scheme == HH_SCHEME_EXACT_STEPS)))
    return ctx->fail(HH_ERR_UNSUPPORTED, "path-dependent payoffs run on the stepping schemes (EulerMaruyama; BlackScholesExact "
                     "increments for LognormalDynamics); scheme %d saves no intermediate dates", s->scheme);
  if (s->precision != HH_PREC_F64) return ctx->fail(HH_ERR_UNSUPPORTED, "path-dependent payoffs are computed in binary64");
  const int M = s->n_steps;
  if (monitor_every < 1 || M % monitor_every != 0)
    return ctx->fail(HH_ERR_ARG, "n_steps (%d) must be a positive multiple of monitor_every (%d)", M, monitor_every);
  bool arith = path_stats != nullptr;
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
    arith = arith || q.kind == HH_PD_ASIAN_ARITH;
  }
  const int64_t N = s->n_paths;
  const bool anti = s->vr == HH_VR_ANTITHETIC;
  const int64_t ncols = anti ? 2 * N : N;
  if (path_stats && path_stats_len < (size_t)ncols * kPdStats)
    return ctx->fail(HH_ERR_ARG, "path_stats buffer too short: %zu < %zu", path_stats_len, (size_t)ncols * kPdStats);
  const bool parity = s->rng_mode == HH_RNG_NORMALS;

  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
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
  // one resident wave (2 blocks of 512 threads per SM with 105 KB of tables + staging each), grid-stride over batches
  const int64_t batches = (N + kPdThreads - 1) / kPdThreads;
  const int grid = (int)(batches < (int64_t)ctx->sm_count * 2 ? batches : (int64_t)ctx->sm_count * 2);
  HH_CUDA(ctx, ctx->d_partials.ensure(sizeof(double) * (size_t)grid * npay * kPdAcc));
  HH_CUDA(ctx, ctx->d_final.ensure(sizeof(double) * (size_t)npay * kPdAcc));
  a.partials = ctx->d_partials.as<double>();

  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  const bool ukey = a.seeds == nullptr;
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
    r->kernel_ms = ms;
  }
  return HH_OK;
}

}  // namespace hh
