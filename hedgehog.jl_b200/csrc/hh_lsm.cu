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
//       polynomials of an affinely mapped spot u = a S + b, so that the Gram matrix is well conditioned (raw
//       monomials of S ~ 100 to degree 5 give entries ~1e20), and T_i T_j = (T_{i+j} + T_{|i-j|}) / 2 means only
//       2 deg + 1 moment sums are needed for it. The fitted VALUES are those of the reference's least-squares
//       polynomial up to rounding; this is a reduction, not a dense contraction, so no tensor cores.
//   update_stopping_info! (:156-165)                          -> folded into the next pass (strict >), tau written only
//       when stopping info is requested.
//   price (:132-133)                                          -> last pass: sum / sum of squares of D z.
// Multi-GPU: columns are sharded; the per-date moment vector (3 deg + 3 doubles) is sum-allreduced through the
// caller's hh_comm callback between the pass and the fit, so every rank fits the same global polynomial.
#include <cmath>
#include <cstring>
#include <vector>

#include "hh_ctx.h"
#include "hh_fastnormal.cuh"
#include "hh_paths.cuh"

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
};

template <bool ANTI>
__global__ void __launch_bounds__(kLsmThreads) lsm_paths_kernel(const LsmPathArgs a) {
  __shared__ FastNormalTables s_tables;
  if (!a.parity) {
    load_fast_tables(&s_tables);
    __syncthreads();
  }
  const int M = a.n_steps;
  for (int64_t i = (int64_t)blockIdx.x * kLsmThreads + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * kLsmThreads) {
    uint64_t key = a.base_seed, idx = (uint64_t)(a.path_offset + i);
    if (a.seeds) {
      key = a.seeds[i];
      idx = 0;
    }
    const double *z = a.parity ? a.normals + (size_t)i * (size_t)M : nullptr;
    double Sp = a.S0, Sm = a.S0;
    double *gp = a.grid + i;
    double *gm = a.grid + a.n + i;
    gp[0] = Sp;
    if (ANTI) gm[0] = Sm;
#pragma unroll 1
    for (int n = 0; n < M; n += 2) {
      double za, zb;
      if (a.parity) {
        za = z[n];
        zb = n + 1 < M ? z[n + 1] : 0.0;
      } else {
        const u32x4 w = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)(n >> 1), 0u, (uint32_t)key,
                                      (uint32_t)(key >> 32));
        fast_normal_pair(&s_tables, w.x, w.y, w.z, w.w, za, zb);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (n + h < M) {
          const double zz = h ? zb : za;
          // GeometricBrownianMotionProcess increment [upstream]: S += S (exp((r - s^2/2) dt + s sqrt(dt) Z) - 1)
          const double e = a.sig_sqdt * zz;
          Sp = fma(Sp, exp(a.dt_drift + e) - 1.0, Sp);
          gp[(size_t)(n + h + 1) * a.stride] = Sp;
          if (ANTI) {  // same normals, sigma -> -sigma (montecarlo.jl:270-284)
            Sm = fma(Sm, exp(a.dt_drift - e) - 1.0, Sm);
            gm[(size_t)(n + h + 1) * a.stride] = Sm;
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
  double active;
  double used_degree;
  double count;
};

struct LsmPassArgs {
  int64_t ncols;
  const double *S_next;  // G[t+1]
  const double *S_cur;   // G[t]   (unused in the last pass)
  double *z;             // cash flow per column, discounted to the current date
  int32_t *tau;          // nullable
  const LsmFit *fit_next;
  double *partials;      // [grid][nacc]
  double D, strike, cp, ua, ub;
  int t_next;
  int first;  // t+1 is the terminal date: z = payoff(S_M)
  int last;   // t = 0: no regression, accumulate sum / sumsq of D z
  int reverse;             // walk the columns from the end (alternates per pass for L2 reuse)
  unsigned int *done;      // arrival counter of the blocks of this pass
  double *moments;         // [nacc] sums over all blocks, written by the last block
  LsmFit *fit_out;         // nullable: where the last block writes the fit of date t (single-GPU form)
};

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

template <int DEG>
__device__ __forceinline__ void lsm_column(const LsmPassArgs &a, const double *fit, bool fit_active, double sn, double sc,
                                           double zin, int64_t p, double &zout, double *acc) {
  constexpr int NM = 2 * DEG + 1;
  double zz;
  if (a.first) {
    zz = fmax(a.cp * (sn - a.strike), 0.0);  // stopping_info = (nsteps, payoff(S_T))  :112
  } else {
    zz = zin;
    const double e = fmax(a.cp * (sn - a.strike), 0.0);
    if (fit_active && e > 0.0) {
      const double cont = clenshaw<DEG>(fit, fma(a.ua, sn, a.ub));  // poly.(x)  :127
      if (e > cont) {                                               // strict  :163-164
        zz = e;
        if (a.tau) a.tau[p] = a.t_next;
      }
    }
  }
  zz *= a.D;  // discount^(tau - t) one date at a time  :117-118
  zout = zz;
  if (a.last) {
    acc[0] += zz;
    acc[1] = fma(zz, zz, acc[1]);
    acc[2] += 1.0;
  } else {
    const double e0 = a.cp * (sc - a.strike);
    if (e0 > 0.0) {  // in the money  :120-121
      const double u = fma(a.ua, sc, a.ub);
      const double u2 = 2.0 * u;
      double t0 = 1.0, t1 = u;
      acc[0] += 1.0;
      acc[NM] += zz;
      if (DEG >= 1) {
        acc[1] += u;
        acc[NM + 1] = fma(u, zz, acc[NM + 1]);
      }
#pragma unroll
      for (int k = 2; k < NM; ++k) {
        const double tk = fma(u2, t1, -t0);
        t0 = t1;
        t1 = tk;
        acc[k] += tk;
        if (k <= DEG) acc[NM + k] = fma(tk, zz, acc[NM + k]);
      }
      acc[NM + DEG + 1] += 1.0;
    }
  }
}

// Sum the per-block partials in a fixed order: moments[c] = sum_b partials[b][c]. One warp per accumulator
// (lanes stride over the blocks, then a shuffle tree), so the dependent chain is nblocks / 32 long.
__device__ __forceinline__ void lsm_reduce_partials(const double *partials, int nblocks, int nacc, double *moments) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int c = warp; c < nacc; c += nwarps) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    int b = lane;
    for (; b + 96 < nblocks; b += 128) {
      t0 += partials[(size_t)b * nacc + c];
      t1 += partials[(size_t)(b + 32) * nacc + c];
      t2 += partials[(size_t)(b + 64) * nacc + c];
      t3 += partials[(size_t)(b + 96) * nacc + c];
    }
    for (; b < nblocks; b += 32) t0 += partials[(size_t)b * nacc + c];
    double t = (t0 + t1) + (t2 + t3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) moments[c] = t;
  }
}

__global__ void __launch_bounds__(256) lsm_reduce_kernel(const double *partials, int nblocks, int nacc, double *moments) {
  lsm_reduce_partials(partials, nblocks, nacc, moments);
}

// Normal equations in the Chebyshev basis: G[i][j] = (m[i+j] + m[|i-j|]) / 2, rhs[i] = r[i]; Cholesky.
// If the matrix is numerically singular at the requested degree (fewer distinct in-the-money spots than
// coefficients), the leading block that factorises is used (a lower-degree fit in the same nested basis).
__device__ void lsm_fit(const double *moments, int deg, LsmFit *out) {
  const int nm = 2 * deg + 1;
  const double *m = moments, *r = moments + nm;
  const double count = moments[nm + deg + 1];
  LsmFit f;
  for (int k = 0; k <= kLsmMaxDeg; ++k) f.c[k] = 0.0;
  f.active = 0.0;
  f.used_degree = -1.0;
  f.count = count;
  if (count > 0.0) {
    double L[kLsmMaxDeg + 1][kLsmMaxDeg + 1];
    int n = 0;  // size of the leading block that factorises
    for (int j = 0; j <= deg; ++j) {
      double d = 0.5 * (m[2 * j] + m[0]);
      for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
      if (!(d > 1e-13 * m[0])) break;
      const double ljj = sqrt(d);
      L[j][j] = ljj;
      for (int i = j + 1; i <= deg; ++i) {
        double s = 0.5 * (m[i + j] + m[i - j]);
        for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
        L[i][j] = s / ljj;
      }
      n = j + 1;
    }
    if (n > 0) {
      double y[kLsmMaxDeg + 1];
      for (int i = 0; i < n; ++i) {
        double s = r[i];
        for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
        y[i] = s / L[i][i];
      }
      for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= L[k][i] * f.c[k];
        f.c[i] = s / L[i][i];
      }
      f.active = 1.0;
      f.used_degree = (double)(n - 1);
    }
  }
  *out = f;
}

__global__ void lsm_fit_kernel(const double *moments, int deg, LsmFit *out) {
  if (threadIdx.x == 0) lsm_fit(moments, deg, out);
}

template <int DEG>
__global__ void __launch_bounds__(kLsmThreads) lsm_pass_kernel(const LsmPassArgs a) {
  constexpr int NACC = lsm_nacc<DEG>();
  __shared__ double s_red[NACC][kLsmThreads / 32];
  __shared__ bool s_last;
  double fit[DEG + 1];
  bool fit_active = false;
  if (!a.first) {
    fit_active = a.fit_next->active != 0.0;
#pragma unroll
    for (int k = 0; k <= DEG; ++k) fit[k] = a.fit_next->c[k];
  }
  double acc[NACC];
#pragma unroll
  for (int c = 0; c < NACC; ++c) acc[c] = 0.0;

  // Two columns per thread per iteration (16 B loads and stores; the slices are 256 B aligned), and the loads of the
  // next iteration are issued before the arithmetic of the current one so that every thread keeps 6 x 16 B in flight.
  const int64_t npairs = a.ncols >> 1;
  const double2 *Sn2 = reinterpret_cast<const double2 *>(a.S_next);
  const double2 *Sc2 = reinterpret_cast<const double2 *>(a.S_cur);
  double2 *z2 = reinterpret_cast<double2 *>(a.z);
  const int64_t step = (int64_t)gridDim.x * kLsmThreads;
  int64_t q = (int64_t)blockIdx.x * kLsmThreads + threadIdx.x;
  // passes alternate their direction over the columns: what the previous pass touched last (z and the shared date
  // slice) is what this pass touches first, while it is still in L2
  const bool rev = a.reverse != 0;
  auto at = [&](int64_t k) { return rev ? npairs - 1 - k : k; };
  double2 sn = make_double2(0.0, 0.0), sc = sn, zi = sn;
  if (q < npairs) {
    sn = __ldcs(Sn2 + at(q));
    if (!a.last) sc = __ldg(Sc2 + at(q));
    if (!a.first) zi = z2[at(q)];
  }
  while (q < npairs) {
    const int64_t qn = q + step;
    double2 sn_n = make_double2(0.0, 0.0), sc_n = sn_n, zi_n = sn_n;
    if (qn < npairs) {
      sn_n = __ldcs(Sn2 + at(qn));
      if (!a.last) sc_n = __ldg(Sc2 + at(qn));
      if (!a.first) zi_n = z2[at(qn)];
    }
    double2 zo;
    const int64_t col = 2 * at(q);
    lsm_column<DEG>(a, fit, fit_active, sn.x, sc.x, zi.x, col, zo.x, acc);
    lsm_column<DEG>(a, fit, fit_active, sn.y, sc.y, zi.y, col + 1, zo.y, acc);
    z2[at(q)] = zo;
    sn = sn_n;
    sc = sc_n;
    zi = zi_n;
    q = qn;
  }
  if ((a.ncols & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t p = a.ncols - 1;
    double zo;
    lsm_column<DEG>(a, fit, fit_active, a.S_next[p], a.last ? 0.0 : a.S_cur[p], a.first ? 0.0 : a.z[p], p, zo, acc);
    a.z[p] = zo;
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NACC; ++c) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[c][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kLsmThreads / 32; ++w) t += s_red[threadIdx.x][w];
    a.partials[(size_t)blockIdx.x * NACC + threadIdx.x] = t;
  }
  // the block that finishes last sums the partials of all blocks (fixed order) and fits the date's polynomial
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence();
    lsm_reduce_partials(a.partials, (int)gridDim.x, NACC, a.moments);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (a.fit_out) lsm_fit(a.moments, DEG, a.fit_out);
      *a.done = 0u;
    }
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

template <int DEG>
static cudaError_t launch_pass(const LsmPassArgs &a, int grid, cudaStream_t st) {
  lsm_pass_kernel<DEG><<<grid, kLsmThreads, 0, st>>>(a);
  return cudaGetLastError();
}

template <int DEG>
static int pass_occupancy() {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lsm_pass_kernel<DEG>, kLsmThreads, 0) != cudaSuccess || occ < 1) occ = 1;
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

int lsm_american(hh_ctx *ctx, const hh_model *m, const hh_sim *s, const hh_payoff *payoff, int degree,
                 double step_discount, const hh_comm *comm, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                 double *spot_paths) {
  int rc = validate_model_sim(ctx, m, s);
  if (rc) return rc;
  if (!payoff || !out) return ctx->fail(HH_ERR_ARG, "payoff/out is NULL");
  if (degree < 0 || degree > kLsmMaxDeg) return ctx->fail(HH_ERR_ARG, "degree must be in [0, %d] (got %d)", kLsmMaxDeg, degree);
  // Q7: the reference reads component 1 of the saved state as the spot (least_squares_montecarlo.jl:53), which is
  // only true for the S-space BlackScholesExact generator — the only LSM configuration it tests.
  if (m->kind != HH_MODEL_GBM || s->scheme != HH_SCHEME_EXACT_STEPS)
    return ctx->fail(HH_ERR_UNSUPPORTED, "LSM runs on LognormalDynamics + BlackScholesExact paths (SURVEY Q7)");
  if ((stop_idx == nullptr) != (stop_val == nullptr))
    return ctx->fail(HH_ERR_ARG, "stop_idx and stop_val must be both NULL or both non-NULL");
  if (comm && !comm->allreduce_sum_f64) return ctx->fail(HH_ERR_ARG, "hh_comm without an allreduce callback");

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
  pa.S0 = m->S0;
  pa.dt_drift = (m->r - 0.5 * (m->sigma * m->sigma)) * dt;
  pa.sig_sqdt = m->sigma * sqrt(dt);
  if (parity) {
    const size_t bytes = sizeof(double) * (size_t)N * (size_t)M;
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
  const int64_t path_blocks = (N + kLsmThreads - 1) / kLsmThreads;
  const int grid_paths = (int)(path_blocks < (int64_t)ctx->sm_count * 8 ? path_blocks : (int64_t)ctx->sm_count * 8);
  const int64_t pass_blocks = ((ncols >> 1) + kLsmThreads - 1) / kLsmThreads;
  const int64_t resident = (int64_t)ctx->sm_count * pass_occupancy_deg(degree);  // one wave: every block is resident
  int grid_pass = (int)(pass_blocks < resident ? pass_blocks : resident);
  if (grid_pass < 1) grid_pass = 1;
  const int nacc = 3 * degree + 3;
  HH_CUDA(ctx, ctx->d_lsm_partials.ensure(sizeof(double) * (size_t)grid_pass * nacc));
  // state: [moments (nacc)] [done counter] [fits (M+1)]
  const size_t done_off = ((size_t)nacc * sizeof(double) + 255) & ~(size_t)255;
  const size_t fit_off = done_off + 256;
  HH_CUDA(ctx, ctx->d_lsm_state.ensure(fit_off + sizeof(LsmFit) * (size_t)(M + 1)));
  double *d_moments = ctx->d_lsm_state.as<double>();
  LsmFit *d_fits = reinterpret_cast<LsmFit *>(static_cast<char *>(ctx->d_lsm_state.ptr) + fit_off);
  HH_CUDA(ctx, cudaMemsetAsync(ctx->d_lsm_state.ptr, 0, fit_off + sizeof(LsmFit) * (size_t)(M + 1), st));

  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  if (anti) lsm_paths_kernel<true><<<grid_paths, kLsmThreads, 0, st>>>(pa);
  else lsm_paths_kernel<false><<<grid_paths, kLsmThreads, 0, st>>>(pa);
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));

  if (want_stop) {
    lsm_fill_tau_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->d_tau.as<int32_t>(), ncols, M);
    HH_CUDA(ctx, cudaGetLastError());
  }

  // Chebyshev variable u = ua S + ub: the in-the-money side of the strike mapped to about [-1, 1]
  double ua, ub;
  if (payoff->cp < 0) {  // put: S in (0, K)
    ua = 2.0 / payoff->strike;
    ub = -1.0;
  } else {  // call: S in (K, Smax), Smax = a 6-sigma excursion of the terminal spot
    const double drift = fmax((m->r - 0.5 * m->sigma * m->sigma) * m->T, 0.0);
    const double smax = fmax(m->S0, payoff->strike) * exp(drift + 6.0 * fabs(m->sigma) * sqrt(m->T));
    ua = 2.0 / (smax - payoff->strike);
    ub = -1.0 - ua * payoff->strike;
  }

  LsmPassArgs a;
  memset(&a, 0, sizeof a);
  a.ncols = ncols;
  a.z = ctx->d_cash.as<double>();
  a.tau = want_stop ? ctx->d_tau.as<int32_t>() : nullptr;
  a.partials = ctx->d_lsm_partials.as<double>();
  a.D = step_discount;
  a.strike = payoff->strike;
  a.cp = payoff->cp;
  a.ua = ua;
  a.ub = ub;
  a.done = reinterpret_cast<unsigned int *>(static_cast<char *>(ctx->d_lsm_state.ptr) + done_off);
  a.moments = d_moments;
  const double *G = ctx->d_grid.as<double>();
  // pass(t), t = M-1 .. 0: decision at t+1 (with fit[t+1]), one-step discount, moments of date t (t >= 1)
  for (int t = M - 1; t >= 0; --t) {
    a.S_next = G + (size_t)(t + 1) * stride;
    a.S_cur = G + (size_t)t * stride;
    a.fit_next = d_fits + (t + 1);
    a.t_next = t + 1;
    a.first = (t + 1 == M);
    a.last = (t == 0);
    const bool exchange = t >= 1 && comm && comm->world > 1;
    a.fit_out = (t >= 1 && !exchange) ? d_fits + t : nullptr;
    a.reverse = (M - 1 - t) & 1;
    HH_CUDA(ctx, launch_pass_deg(degree, a, grid_pass, st));
    if (exchange) {
      if (comm->allreduce_sum_f64(comm->user, d_moments, (size_t)nacc, (void *)st) != 0)
        return ctx->fail(HH_ERR_COMM, "allreduce callback failed at date %d", t);
      lsm_fit_kernel<<<1, 32, 0, st>>>(d_moments, degree, d_fits + t);
      HH_CUDA(ctx, cudaGetLastError());
    }
  }
  HH_CUDA(ctx, cudaEventRecord(ctx->ev2, st));

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
    HH_CUDA(ctx, cudaMemcpyAsync(stop_idx, ctx->d_tau.ptr, sizeof(int32_t) * (size_t)ncols, cudaMemcpyDeviceToHost, st));
    HH_CUDA(ctx, cudaMemcpyAsync(stop_val, ctx->d_misc.ptr, sizeof(double) * (size_t)ncols, cudaMemcpyDeviceToHost, st));
  }
  HH_CUDA(ctx, cudaStreamSynchronize(st));
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
      HH_CUDA(ctx, cudaMemcpyAsync(spot_paths + (size_t)p0 * nrows, ctx->d_terminal.ptr, sizeof(double) * (size_t)np * nrows,
                                   cudaMemcpyDeviceToHost, st));
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
