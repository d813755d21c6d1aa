// hh_bk.cu — Broadie-Kaya exact simulation of the Heston model: MonteCarlo(HestonDynamics(), HestonBroadieKaya(), cfg).
//
// Replaces (reference paths):
//   rand(rng, ::LogHestonDistribution)        src/distributions/heston.jl:246-259     one exact transition
//   sample_V_T                                :125-133     V' = c * NoncentralChisq(d, lambda)
//   sample_integral_V -> sample_from_cf       :140-143, src/distributions/sample_from_cf.jl:27-41
//       moments_from_cf :50-64, cdf_from_cf :75-96, inverse_cdf :105-135
//   sample_log_S_T                            heston.jl:278-300
//   HestonNoise stepping (multi-date)         :82-91  (restart from (S, V) each date)
//
// A transition (V, S) -> (V', S') over one date interval is
//   1. V' (noncentral chi-square: chi2(d-1) + (Z + sqrt(lambda))^2 for d > 1, Poisson mixture otherwise; gamma variates by
//      Marsaglia-Tsang) from the trajectory's own Philox counter stream,
//   2. the characteristic function of int V at +h0 (the finite-difference moments of the reference through Phi(0) = 1,
//      Phi(-a) = conj Phi(a)),
//   3. Phi(h j) ONCE per term into a shared-memory table c_j = (2/pi) Re Phi(h j) / j — the reference recomputes the whole
//      series (one complex Bessel function per term) for every root-finder iteration although it does not depend on x
//      (sample_from_cf.jl:38, 84-93); terms beyond the table spill to a global slab,
//   4. the inversion of F(x) = h x / pi + sum_j c_j sin(h j x) at its uniform by safeguarded Newton inside [0, mean + 11 sd],
//      to machine precision (the reference accepts |F(x) - u| <= 1e-4; any root of the same F satisfies that),
//   5. log S'.
// Step 1 of ALL dates runs first, one trajectory per thread (the variance chain depends neither on the integrals nor on the
// spot); steps 2-4 then run over all transitions of the job in an order sorted by log2(V V'), so that the lanes of a warp
// take the same Bessel branch; step 5 walks each trajectory through its dates ("the transition-sorted pipeline" below).
// Payoffs are then reduced from the terminal spots by hh_european.cu's terminal_payoff kernel.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hh_bessel.cuh"
#include "hh_ctx.h"
#include "hh_device.cuh"
#include "hh_fastnormal.cuh"

namespace hh {

constexpr int kBkThreads = 128;
constexpr int kBkDefaultMinb = 3;
constexpr int kBkTable = 32;  // table entries per thread in shared memory (32 KB per block); longer series (mean 12.3 at C4) spill to the HBM slab

// terminal spots -> payoff sums (hh_european.cu)
int terminal_payoffs_launch(hh_ctx *ctx, const double *d_terminal, int64_t n, const hh_payoff *payoffs, int npay,
                            int64_t extra_nonfinite);

struct BkRng {  // Philox counter stream of one trajectory: counter = (idx_lo, idx_hi, date, draw)
  uint32_t c0, c1, c2, draw, k0, k1;
  const FastNormalTables *tb;  // shared-memory Box-Muller tables (hh_fastnormal.cuh): the same pair as libm's to ~4e-16
  __device__ __forceinline__ u32x4 block() { return philox4x32_10(c0, c1, c2, draw++, k0, k1); }
  __device__ __forceinline__ void normals(double &z1, double &z2) {
    const u32x4 w = block();
    fast_normal_pair(tb, w.x, w.y, w.z, w.w, z1, z2);
  }
  __device__ __forceinline__ void uniforms(double &u1, double &u2) {  // both in (0, 1)
    const u32x4 w = block();
    u1 = u01_for_log(w.x, w.y);
    u2 = u01_for_log(w.z, w.w);
  }
};

// Gamma(alpha, 1), alpha > 0 (Marsaglia & Tsang 2000; alpha < 1 through Gamma(alpha + 1) U^(1/alpha)).
__device__ double bk_gamma(BkRng &rng, double alpha) {
  double boost = 1.0;
  if (alpha < 1.0) {
    double u, u_;
    rng.uniforms(u, u_);
    boost = pow(u, 1.0 / alpha);
    alpha += 1.0;
  }
  const double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 1000; ++it) {
    double z, z_, u, u_;
    rng.normals(z, z_);
    const double t = 1.0 + c * z;
    if (t <= 0.0) continue;
    const double v = t * t * t;
    rng.uniforms(u, u_);
    const double z2 = z * z;
    if (u < 1.0 - 0.0331 * z2 * z2 || log(u) < 0.5 * z2 + d * (1.0 - v + log(v))) return boost * d * v;
  }
  return boost * d;  // unreachable in practice (acceptance > 95% per round)
}

// Poisson(mu): sequential inversion for mu < 30, PTRS transformed rejection (Hoermann 1993) above.
__device__ double bk_poisson(BkRng &rng, double mu) {
  if (mu <= 0.0) return 0.0;
  if (mu < 30.0) {
    double u, u_;
    rng.uniforms(u, u_);
    double p = exp(-mu), s = p, k = 0.0;
    while (u > s && k < 500.0) {
      k += 1.0;
      p *= mu / k;
      s += p;
    }
    return k;
  }
  const double slam = sqrt(mu), loglam = log(mu);
  const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int it = 0; it < 1000; ++it) {
    double U, V;
    rng.uniforms(U, V);
    U -= 0.5;
    const double us = 0.5 - fabs(U);
    const double k = floor((2.0 * a / us + b) * U + mu + 0.43);
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -mu + k * loglam - lgamma(k + 1.0)) return k;
  }
  return floor(mu);
}

// NoncentralChisq(d, lambda): heston.jl:131 (Distributions.jl [upstream]; any exact sampler has the same law)
__device__ double bk_ncx2(BkRng &rng, double dof, double lam) {
  if (dof > 1.0) {
    double z, z_;
    rng.normals(z, z_);
    const double s = z + sqrt(lam);
    return 2.0 * bk_gamma(rng, 0.5 * (dof - 1.0)) + s * s;
  }
  const double n = bk_poisson(rng, 0.5 * lam);
  return 2.0 * bk_gamma(rng, 0.5 * dof + n);
}

struct BkInversion {
  double mean, var, h, x, resid;
  int J, status, iters;  // status: 0 Newton inside the bracket, 1 unbracketed secant accepted, 2 fell back to max_guess
};

// Coefficient table: the first `cap` entries in shared memory (stride = blockDim.x doubles; read and written through the
// shared window, see FastRef), the rest in a global slab.
struct BkTable {
  unsigned sh_saddr;  // shared-window address of this thread's first entry
  unsigned sh_stride_bytes;
  double *slab;
  int64_t slab_stride;
  int cap;
  __device__ __forceinline__ double get(int j) const {
    return j < cap ? lds_f64(sh_saddr + (unsigned)j * sh_stride_bytes) : slab[(int64_t)(j - cap) * slab_stride];
  }
  __device__ __forceinline__ void set(int j, double v) const {
    if (j < cap) asm volatile("st.shared.f64 [%0], %1;" ::"r"(sh_saddr + (unsigned)j * sh_stride_bytes), "d"(v) : "memory");
    else slab[(int64_t)(j - cap) * slab_stride] = v;
  }
};

// F(x) - u and F'(x) for F(x) = h x / pi + sum_j c_j sin(j h x)   (cdf_from_cf, sample_from_cf.jl:75-96)
// sin(j th), cos(j th) by the Chebyshev recurrence, two terms per iteration (ping-pong registers, no moves); the
// shared-memory part of the table is read through the shared window without a per-term range test.
__device__ __noinline__ void bk_cdf(const BkTable &tb, FastRef ft, int J, double h, double x, double &F, double &dF) {
  const double th = h * x;
  double s1, c1;
  fsincos(ft, th, s1, c1);
  const double two_c = 2.0 * c1;
  double sa = 0.0, sb = s1;  // sin((j-1) th), sin(j th)
  double ca = 1.0, cb = c1;  // cos((j-1) th), cos(j th)
  double acc = 0.0, dacc = 0.0, dj = 1.0;
  const int J1 = J < tb.cap ? J : tb.cap;
  unsigned addr = tb.sh_saddr;
  const unsigned st = tb.sh_stride_bytes;
  int j = 0;
#pragma unroll 1
  for (; j + 2 <= J1; j += 2, addr += 2 * st) {
    const double c0 = lds_f64(addr), c1j = lds_f64(addr + st);
    acc = fma(c0, sb, acc);
    dacc = fma(c0 * dj, cb, dacc);
    sa = fma(two_c, sb, -sa);  // sin((j+1) th)
    ca = fma(two_c, cb, -ca);
    acc = fma(c1j, sa, acc);
    dacc = fma(c1j * (dj + 1.0), ca, dacc);
    sb = fma(two_c, sa, -sb);  // sin((j+2) th)
    cb = fma(two_c, ca, -cb);
    dj += 2.0;
  }
  for (; j < J; ++j) {  // the odd term, and terms beyond the shared-memory table (global slab)
    const double cj = tb.get(j);
    acc = fma(cj, sb, acc);
    dacc = fma(cj * dj, cb, dacc);
    const double sn = fma(two_c, sb, -sa), cn = fma(two_c, cb, -ca);
    sa = sb; sb = sn;
    ca = cb; cb = cn;
    dj += 1.0;
  }
  constexpr double kInvPi = 1.0 / kBesselPi;
  F = fma(h * x, kInvPi, acc);
  dF = fma(h, kInvPi, h * dacc);
}

// Phi is the exponential of a difference of terms of size |log I_nu(z_kappa)| + (V0 + VT) eta_kappa / sigma^2 =: L, so it
// carries ~4 eps L of rounding, and the second difference of moments_from_cf divides that by h0^2 = 1e-4. For short horizons
// or a low vol of vol (sigma tau < ~0.01: L ~ 8 V / (sigma^2 tau), var ~ sigma^2 V tau^3 / 3) that noise EXCEEDS the
// variance: the reference's estimate is then a random number, negative half of the time (-> the 1e-12 floor, a Fourier
// grid narrower than the law, a wrapped CDF: the call of sigma = 0.1 on 52 dates came out 25 standard errors high).
// Called when the noise is above 2 % of the estimate (noise_scale = 4 eps / h0^2 / 0.02): the variance is re-read from
// the modulus at a step scaled to the law, h1 = 0.1 / mean:  log |Phi(h1)|^2 = -var h1^2 + O(kappa_4 h1^4), signal
// 0.01 (sd / mean)^2 against the same 4 eps L. Out of line: C4 never gets here.
__device__ __noinline__ double bk_variance_from_modulus(const BkParams &p, const BkCf &it, double mean, double var) {
  const double h1 = 0.1 / mean;
  if (!(h1 > p.h_fd) || !(h1 < 1e300)) return var;
  double th1 = nan("");
  const cplx q = bk_chf(p, it, h1, th1);
  const double e = 1.0 - q.re;
  const double m2 = 2.0 * e - e * e - q.im * q.im;  // 1 - |Phi(h1)|^2
  return (m2 > 0.0 && m2 < 0.5) ? -flog(p.ord.ft, 1.0 - m2) / (h1 * h1) : var;  // m2 >~ 1e-7: log(1 - m2) is good to 1e-9
}

// sample_from_cf (sample_from_cf.jl:27-41) for a given uniform u.
__device__ BkInversion bk_sample_integral(const BkParams &p, double V0, double VT, double u, const BkTable &tb) {
  BkInversion r;
  const BkCf it = bk_cf_init(p, V0, VT);
  // moments_from_cf :50-64 differences Phi at +h0, 0, -h0. A characteristic function has Phi(0) = 1 and
  // Phi(-a) = conj Phi(a), so ONE evaluation gives the same central differences: mean = Im Phi(h0) / h0,
  // var = (2 - 2 Re Phi(h0)) / h0^2 - mean^2. The reference's three evaluations carry that identity only to rounding, and
  // its second difference amplifies the rounding of Phi(0) by 1 / h0^2 = 1e4 (its variance scatters by ~1e-4 relative
  // between machines and libraries, tests/test_gpu_bk.py); this form is inside that scatter and saves two of the
  // ~16 Bessel evaluations of a transition.
  double th = nan("");
  const cplx pp = bk_chf(p, it, p.h_fd, th);
  const double mean = pp.im / p.h_fd;                                                    // real(-i (pp - pm) / 2h)
  double var = (2.0 - 2.0 * pp.re) / (p.h_fd * p.h_fd) - mean * mean;                    // :60-61
  // rounding noise of that second difference against the estimate (see bk_variance_from_modulus)
  if (p.widen_fd && !(p.noise_scale * (fabs(it.logIk.re) + it.vsum_s * p.eta_k) <= var)) var = bk_variance_from_modulus(p, it, mean, var);
  const double s2 = fmax(var, 1e-12);                                                    // :32
  const double sd = sqrt(s2);
  const double ns = mean + sd * normcdfinv(u);                                           // :33
  const double guess = ns > 0.0 ? ns : mean * 0.01;                                      // :34
  const double max_guess = mean + 11.0 * sd;                                             // :35
  const double h = kBesselPi / (mean + (double)p.n_std * sd);                            // :37
  r.mean = mean;
  r.var = var;
  r.h = h;
  if (!(fabs(mean) < 1e300) || !(fabs(var) < 1e300) || !(h > 0.0) || !(h < 1e300)) {
    // non-finite moments (non-finite inputs): the reference propagates NaN; so does this, without running the loops below
    r.x = r.resid = nan("");
    r.J = 0;
    r.status = 2;
    r.iters = 0;
    return r;
  }
  // series coefficients, x-independent (:84-93)
  th = nan("");
  int J = 0;
  const double stop = kBesselPi * p.cf_tol / 2.0;
  for (int j = 1; j <= p.max_terms; ++j) {
    const cplx phi = bk_chf(p, it, h * (double)j, th);
    tb.set(j - 1, (2.0 / kBesselPi) * phi.re * rcp_fast((double)j));
    J = j;
    if (cabs2(phi) < (stop * (double)j) * (stop * (double)j)) break;  // |phi| / j < stop
  }
  r.J = J;
  // inverse_cdf :105-135. Safeguarded Newton from the reference's initial guess inside [0, max_guess], f(0) = -u < 0.
  // The right end is evaluated only when an iterate runs into it before any f >= 0 has been seen (u ~ 1, where the
  // truncated series may never reach u): for every other u that saves one of ~5 CDF evaluations and changes no iterate.
  // Once a Newton step is shorter than 1e-9 of the bracket the next iterate is within ~1e-18 of the root (quadratic
  // convergence on a finite trigonometric sum) and is returned without a confirming evaluation.
  double F, dF;
  int iters = 0;
  double fmax_ = 0.0;
  bool no_bracket = false;
  if (u <= 0.0) {
    r.x = 0.0;
    r.resid = 0.0;
    r.status = 0;
  } else {
    double lo = 0.0, hi = max_guess;
    double x = fmin(fmax(guess, 0.0), max_guess);
    double f = 0.0;
    bool hi_known = false;  // some f(hi) >= 0 has been evaluated
    for (int k = 0; k < 100; ++k) {
      bk_cdf(tb, p.ord.ft, J, h, x, F, dF);
      ++iters;
      f = F - u;
      if (f < 0.0) {
        lo = x;
      } else {
        hi = x;
        hi_known = true;
      }
      if (fabs(f) <= 1e-15 || (hi_known && hi - lo <= 1e-15 * max_guess)) break;
      double xn = dF > 0.0 ? x - f / dF : -1.0;
      if (!(xn > lo && xn < hi)) {
        if (!hi_known) {
          bk_cdf(tb, p.ord.ft, J, h, max_guess, F, dF);
          ++iters;
          fmax_ = F - u;
          if (fmax_ < 0.0) {
            no_bracket = true;
            break;
          }
          hi_known = true;
        }
        xn = 0.5 * (lo + hi);
      } else if (fabs(xn - x) <= 1e-9 * max_guess) {
        x = xn;
        f = 0.0;  // the quadratic model's residual at xn, below rounding
        break;
      }
      if (xn == x) break;
      x = xn;
    }
    if (!no_bracket) {
      r.x = x;
      r.resid = f;
      r.status = 0;
    } else {
      // F(max_guess) < u: no sign change. The reference first lets its secant iteration run (<= maxiter evaluations) and
      // accepts x >= 0 with |F(x) - u| <= atol; otherwise it returns max_guess with a warning (:123-126).
      double x0 = guess, x1 = guess * 1.001 + 1e-12, f0, f1;
      bk_cdf(tb, p.ord.ft, J, h, x0, F, dF);
      f0 = F - u;
      bk_cdf(tb, p.ord.ft, J, h, x1, F, dF);
      f1 = F - u;
      iters += 2;
      for (int k = 2; k < 10; ++k) {
        if (f1 == f0) break;
        const double x2 = x1 - f1 * (x1 - x0) / (f1 - f0);
        x0 = x1; f0 = f1; x1 = x2;
        bk_cdf(tb, p.ord.ft, J, h, x1, F, dF);
        f1 = F - u;
        ++iters;
      }
      if (isfinite(x1) && x1 >= 0.0 && fabs(f1) <= p.atol) {
        r.x = x1;
        r.resid = f1;
        r.status = 1;
      } else {
        r.x = max_guess;
        r.resid = fmax_;
        r.status = 2;
      }
    }
  }
  r.iters = iters;
  return r;
}

// Coefficient tables of the fixed order nu, by all threads of the block (IEEE divisions: once per block).
//   rk4: groups of four, R_{k,i} = prod_{m=k}^{k+i-1} 1 / (m (nu + m)), k = 1, 5, 9, ...   (bessel_series_sum)
//   bk2: groups of four running products of b_k = (4 nu^2 - (2k - 1)^2) / (8 k), k = 1, 5, 9, ...   (bessel_hankel_sums)
__device__ __forceinline__ void bk_fill_order_tables(double nu, double *rk4, double *bk2) {
  for (int g = threadIdx.x; g < kSeriesMaxTerms / 4; g += blockDim.x) {
    double r = 1.0;
    for (int i = 0; i < 4; ++i) {
      const double k = (double)(4 * g + 1 + i);
      r *= 1.0 / (k * (nu + k));
      rk4[4 * g + i] = r;
    }
  }
  for (int g = threadIdx.x; g < kHankelMaxTerms / 4; g += blockDim.x) {
    double b = 1.0;
    for (int i = 0; i < 4; ++i) {
      const double k = (double)(4 * g + 1 + i), odd = 2.0 * k - 1.0;
      b *= (4.0 * nu * nu - odd * odd) / (8.0 * k);
      bk2[4 * g + i] = b;
    }
  }
}

// Per-block copy of the parameters (the outlined functions take them by reference: a reference to the kernel arguments
// would be copied to every thread's local memory) with the coefficient tables of the fixed order nu and the tables of the
// elementary functions. Every kernel that evaluates the characteristic function — the product kernel and the parity
// probes alike — goes through this, so the probes exercise the arithmetic the product runs.
struct alignas(16) BkShared {
  BkFastTables ft;
  double rk4[kSeriesMaxTerms], bk2[kHankelMaxTerms];
  BkParams p;
};
__device__ __forceinline__ const BkParams &bk_block_setup(const BkParams &src, BkShared &sh) {
  bk_fill_order_tables(src.ord.nu, sh.rk4, sh.bk2);
  bk_fill_fast_tables(&sh.ft);
  if (threadIdx.x == 0) {
    sh.p = src;
    sh.p.ord.series_saddr = (unsigned)__cvta_generic_to_shared(sh.rk4);
    sh.p.ord.hankel_saddr = (unsigned)__cvta_generic_to_shared(sh.bk2);
    sh.p.ord.ft.saddr = (unsigned)__cvta_generic_to_shared(&sh.ft);
  }
  __syncthreads();
  return sh.p;
}

// ---- the transition-sorted pipeline (what hh_mc_european / hh_lsm_american / hh_mc_path_dependent launch) --------------
// The variance chain V_0 -> V_1 -> ... does not depend on the integrals or on the spot, so all n_paths x n_dates
// transitions (V0, VT, u) are independent units of work. ncu on the one-thread-per-trajectory kernel above showed 49 % of
// the lanes idle: lanes of a warp sit at the same series index j but at unrelated (V0, VT), and the Bessel arguments of
// C4 straddle the series / Hankel boundary (|z| ~ 20), so half of the warps executed both branches one after the other.
// Here (1) bk_chain_kernel draws every V', u, z from the trajectory's Philox stream (same counters, same draws as above)
// and histograms the transitions by log2(V0 VT) (32 buckets per octave); (2) a counting sort orders the transitions by
// that key; (3) bk_integral_sorted_kernel inverts them in sorted order — the 32 lanes of a warp now hold near-identical
// z_kappa, hence the same Bessel branch, similar series lengths and the same J (+-1); (4) bk_assemble_kernel walks each
// trajectory through its dates. Results are the ones of the kernel above to the last bit (same device functions, and
// no lane depends on its neighbours); the extra HBM traffic (~100 B per transition) is ~1 % of the run time.
constexpr int kBkBuckets = 1024;
constexpr int kBkKeyBase = (1023 - 26) << 5;  // hi word >> 15 of 2^-26: sqrt(V0 VT) from 1.2e-4 (below: bucket 0) to 8
constexpr int kBkScatterTile = 4096;

__device__ __forceinline__ int bk_bucket(double v0, double vt) {
  const int k = (__double2hiint(v0 * vt) >> 15) - kBkKeyBase;  // exponent and 5 mantissa bits of the product
  return min(max(k, 0), kBkBuckets - 1);
}

struct BkChainArgs {
  int64_t n, path_offset;  // trajectories of this chunk, global index of its first trajectory
  uint64_t base_seed;
  const uint64_t *seeds;   // nullable, already offset to the chunk
  int n_dates;
  BkParams p;
  double v0;
  double *V;       // (n_dates + 1) x n, date-major
  double *U, *Z;   // n_dates x n
  unsigned *hist;  // kBkBuckets
};

__global__ void __launch_bounds__(256) bk_chain_kernel(const BkChainArgs a) {
  __shared__ unsigned s_hist[kBkBuckets];
  __shared__ FastNormalTables s_tables;
  load_fast_tables(&s_tables);
  for (int k = threadIdx.x; k < kBkBuckets; k += blockDim.x) s_hist[k] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    BkRng rng;
    rng.tb = &s_tables;
    uint64_t key = a.base_seed, idx = (uint64_t)(a.path_offset + i);
    if (a.seeds) {
      key = a.seeds[i];
      idx = 0;
    }
    rng.c0 = (uint32_t)idx;
    rng.c1 = (uint32_t)(idx >> 32);
    rng.k0 = (uint32_t)key;
    rng.k1 = (uint32_t)(key >> 32);
    double v = a.v0;
    a.V[i] = v;
    for (int n = 0; n < a.n_dates; ++n) {
      rng.c2 = (uint32_t)n;
      rng.draw = 0;
      const double vt = fmax(a.p.c_scale * bk_ncx2(rng, a.p.dof, a.p.lam_scale * v), 1e-300);  // sample_V_T
      double u, z, dummy;
      rng.uniforms(u, dummy);
      rng.normals(z, dummy);
      atomicAdd(&s_hist[bk_bucket(fmax(v, 1e-300), vt)], 1u);
      a.V[(int64_t)(n + 1) * a.n + i] = vt;
      a.U[(int64_t)n * a.n + i] = u;
      a.Z[(int64_t)n * a.n + i] = z;
      v = vt;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kBkBuckets; k += blockDim.x)
    if (s_hist[k]) atomicAdd(&a.hist[k], s_hist[k]);
}

// hist -> exclusive prefix sums, in place (one block of kBkBuckets threads)
__global__ void __launch_bounds__(kBkBuckets) bk_scan_kernel(unsigned *hist) {
  __shared__ unsigned s[kBkBuckets];
  const int t = threadIdx.x;
  const unsigned own = hist[t];
  s[t] = own;
  __syncthreads();
  for (int o = 1; o < kBkBuckets; o <<= 1) {
    const unsigned add = t >= o ? s[t - o] : 0u;
    __syncthreads();
    s[t] += add;
    __syncthreads();
  }
  hist[t] = s[t] - own;
}

// perm[sorted position] = transition index t = date * n + trajectory. Every block ranks a tile of transitions in a
// shared-memory histogram and reserves one range per non-empty bucket (the order inside a bucket is irrelevant).
__global__ void __launch_bounds__(256) bk_scatter_kernel(const double *V, int64_t n, int64_t items, unsigned *cursor,
                                                         unsigned *perm) {
  __shared__ unsigned s_cnt[kBkBuckets], s_base[kBkBuckets];
  constexpr int PER = kBkScatterTile / 256;
  for (int64_t tile = blockIdx.x; tile * kBkScatterTile < items; tile += gridDim.x) {
    for (int k = threadIdx.x; k < kBkBuckets; k += 256) s_cnt[k] = 0;
    __syncthreads();
    int key[PER];
    unsigned rank[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int64_t t = tile * kBkScatterTile + k * 256 + threadIdx.x;
      key[k] = -1;
      if (t < items) {
        key[k] = bk_bucket(fmax(V[t], 1e-300), V[t + n]);
        rank[k] = atomicAdd(&s_cnt[key[k]], 1u);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kBkBuckets; k += 256) {
      const unsigned c = s_cnt[k];
      s_base[k] = c ? atomicAdd(&cursor[k], c) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (key[k] >= 0) perm[s_base[key[k]] + rank[k]] = (unsigned)(tile * kBkScatterTile + k * 256 + threadIdx.x);
    __syncthreads();
  }
}

struct BkIntArgs {
  int64_t items, n;
  BkParams p;
  const double *V, *U;
  const unsigned *perm;
  double *I;
  double *slab;
  int64_t slab_stride;
  unsigned long long *counters;
};

template <int MINB>
__global__ void __launch_bounds__(kBkThreads, MINB) bk_integral_sorted_kernel(const BkIntArgs a) {
  extern __shared__ double s_tab[];
  BkTable tb;
  tb.sh_saddr = (unsigned)__cvta_generic_to_shared(s_tab + threadIdx.x);
  tb.sh_stride_bytes = kBkThreads * (unsigned)sizeof(double);
  tb.cap = kBkTable;
  tb.slab = a.slab + ((int64_t)blockIdx.x * kBkThreads + threadIdx.x);
  tb.slab_stride = a.slab_stride;
  __shared__ BkShared s_sh;
  const BkParams &p = bk_block_setup(a.p, s_sh);
  unsigned long long nfall = 0, sumJ = 0, sumIt = 0, ntr = 0, nsec = 0;
  for (int64_t k = (int64_t)blockIdx.x * kBkThreads + threadIdx.x; k < a.items; k += (int64_t)gridDim.x * kBkThreads) {
    const int64_t t = a.perm[k];
    const BkInversion inv = bk_sample_integral(p, fmax(a.V[t], 1e-300), a.V[t + a.n], a.U[t], tb);  // sample_integral_V
    a.I[t] = inv.x;
    nfall += inv.status == 2;
    nsec += inv.status == 1;
    sumJ += (unsigned)inv.J;
    sumIt += (unsigned)inv.iters;
    ++ntr;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nfall += __shfl_down_sync(0xffffffffu, nfall, o);
    sumJ += __shfl_down_sync(0xffffffffu, sumJ, o);
    sumIt += __shfl_down_sync(0xffffffffu, sumIt, o);
    ntr += __shfl_down_sync(0xffffffffu, ntr, o);
    nsec += __shfl_down_sync(0xffffffffu, nsec, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&a.counters[0], nfall);
    atomicAdd(&a.counters[1], sumJ);
    atomicAdd(&a.counters[2], sumIt);
    atomicAdd(&a.counters[3], ntr);
    atomicAdd(&a.counters[4], nsec);
  }
}

struct BkAsmArgs {
  int64_t n;        // trajectories of this chunk
  int64_t n_total;  // trajectories of the whole launch (row length of `stats`)
  int n_dates;
  BkParams p;
  double x0, s0;
  const double *V, *Z, *I;
  double *terminal, *vterm, *grid, *stats;  // already offset to the chunk's first trajectory; vterm / grid / stats nullable
  int64_t grid_stride;
  int monitor_every;
  double inv_m;
};

// sample_log_S_T (heston.jl:278-300) along the dates of each trajectory, with the HestonNoise restart (:82-91)
__global__ void __launch_bounds__(256) bk_assemble_kernel(const BkAsmArgs a) {
  const BkParams &p = a.p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    double x = a.x0, v = a.V[i];
    if (a.grid) a.grid[i] = a.s0;
    double sum_s = 0.0, sum_x = 0.0, max_x = -INFINITY, min_x = INFINITY;  // running statistics (a.stats only)
    int due = a.monitor_every;
    for (int n = 0; n < a.n_dates; ++n) {
      const int64_t t = (int64_t)n * a.n + i;
      const double vt = a.V[t + a.n], iv = a.I[t], z = a.Z[t];
      const double mu = x + p.r_tau - 0.5 * iv + p.rho_over_xi * (vt - v - p.kappa_theta_tau + p.kappa * iv);
      x = mu + sqrt(p.one_m_rho2 * iv) * z;
      v = vt;
      if (a.grid) a.grid[(int64_t)(n + 1) * a.grid_stride + i] = exp(x);
      if (a.stats && --due == 0) {  // a monitoring date: the transition is exact, so the statistics carry no time-stepping bias
        due = a.monitor_every;
        sum_x += x;
        max_x = fmax(max_x, x);
        min_x = fmin(min_x, x);
        sum_s += exp(x);
      }
    }
    a.terminal[i] = exp(x);
    if (a.vterm) a.vterm[i] = v;
    if (a.stats) {  // S_T, A, G, max S, min S
      a.stats[i] = exp(x);
      a.stats[a.n_total + i] = sum_s * a.inv_m;
      a.stats[2 * a.n_total + i] = exp(sum_x * a.inv_m);
      a.stats[3 * a.n_total + i] = exp(max_x);
      a.stats[4 * a.n_total + i] = exp(min_x);
    }
  }
}

// ---- probes (parity of the deterministic pieces) ---------------------------------------------------------------------
__global__ void bk_chf_kernel(const BkParams p_in, const double *V0, const double *VT, int n, const double *a, int na,
                              double *ore, double *oim) {
  __shared__ BkShared s_sh;
  const BkParams &p = bk_block_setup(p_in, s_sh);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const BkCf it = bk_cf_init(p, V0[i], VT[i]);
  double th = nan("");
  for (int j = 0; j < na; ++j) {
    const cplx r = bk_chf(p, it, a[(size_t)i * na + j], th);
    ore[(size_t)i * na + j] = r.re;
    oim[(size_t)i * na + j] = r.im;
  }
}

__global__ void bk_log_besseli_kernel(const BesselOrder o_in, const double *zr, const double *zi, int n, double *ore, double *oim) {
  __shared__ BkShared s_sh;
  BkParams pin;
  pin.ord = o_in;
  const BesselOrder &o = bk_block_setup(pin, s_sh).ord;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cplx r = log_besseli(o, cplx{zr[i], zi[i]});
  ore[i] = r.re;
  oim[i] = r.im;
}

__global__ void bk_elementary_kernel(int kind, const double *x, const double *y, int n, double *oa, double *ob) {
  __shared__ BkShared s_sh;
  BkParams pin;
  pin.ord = make_bessel_order(0.0);
  const FastRef ft = bk_block_setup(pin, s_sh).ord.ft;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (kind == 0) oa[i] = fexp(ft, x[i]);
  else if (kind == 1) fsincos(ft, x[i], oa[i], ob[i]);
  else if (kind == 2) oa[i] = flog(ft, x[i]);
  else oa[i] = fatan2(ft, y[i], x[i]);
}

// out[i] = {x, mean, var, h, J, status, resid, iters}
__global__ void __launch_bounds__(kBkThreads) bk_integral_kernel(const BkParams p_in, const double *V0, const double *VT,
                                                                 const double *U, int n, double *out, double *slab,
                                                                 int64_t slab_stride) {
  extern __shared__ double s_tab[];
  BkTable tb;
  tb.sh_saddr = (unsigned)__cvta_generic_to_shared(s_tab + threadIdx.x);
  tb.sh_stride_bytes = kBkThreads * (unsigned)sizeof(double);
  tb.cap = kBkTable;
  tb.slab = slab + ((int64_t)blockIdx.x * kBkThreads + threadIdx.x);
  tb.slab_stride = slab_stride;
  __shared__ BkShared s_sh;
  const BkParams &p = bk_block_setup(p_in, s_sh);
  const int i = blockIdx.x * kBkThreads + threadIdx.x;
  if (i >= n) return;
  const BkInversion r = bk_sample_integral(p, V0[i], VT[i], U[i], tb);
  double *o = out + (size_t)i * 8;
  o[0] = r.x; o[1] = r.mean; o[2] = r.var; o[3] = r.h;
  o[4] = (double)r.J; o[5] = (double)r.status; o[6] = r.resid; o[7] = (double)r.iters;
}

__global__ void bk_variance_kernel(const BkParams p, const double *V0, int n, uint64_t seed, double *VT) {
  __shared__ FastNormalTables s_tables;
  load_fast_tables(&s_tables);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  BkRng rng;
  rng.tb = &s_tables;
  rng.c0 = (uint32_t)i;
  rng.c1 = 0;
  rng.c2 = 0;
  rng.draw = 0;
  rng.k0 = (uint32_t)seed;
  rng.k1 = (uint32_t)(seed >> 32);
  VT[i] = p.c_scale * bk_ncx2(rng, p.dof, p.lam_scale * V0[i]);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static int make_params(hh_ctx *ctx, const hh_model *m, double tau, const hh_bk_config *cfg, BkParams &p) {
  if (!(m->kappa > 0.0)) return ctx->fail(HH_ERR_ARG, "Broadie-Kaya needs kappa > 0");
  if (m->xi == 0.0 || !(fabs(m->xi) < 1e300)) return ctx->fail(HH_ERR_ARG, "Broadie-Kaya needs a finite non-zero vol of vol");
  if (!(m->kappa < 1e300) || !(m->theta < 1e300)) return ctx->fail(HH_ERR_ARG, "Broadie-Kaya needs finite kappa and theta");
  if (!(m->theta > 0.0)) return ctx->fail(HH_ERR_ARG, "Broadie-Kaya needs theta > 0");
  if (!(tau > 0.0)) return ctx->fail(HH_ERR_ARG, "transition horizon must be positive");
  memset(&p, 0, sizeof p);
  p.kappa = m->kappa;
  p.xi2 = m->xi * m->xi;
  p.tau = tau;
  const double E = -expm1(-m->kappa * tau);
  p.zeta_k = E / m->kappa;                                             // heston.jl:167
  p.eta_k = m->kappa * (1.0 + exp(-m->kappa * tau)) / E;                // :168
  p.wk = 4.0 * m->kappa * exp(-0.5 * m->kappa * tau) / p.xi2 / E;      // :169
  p.dof = 4.0 * m->kappa * m->theta / p.xi2;                           // :128
  p.ord = make_bessel_order(0.5 * p.dof - 1.0);                        // :165
  p.lam_scale = 4.0 * m->kappa * exp(-m->kappa * tau) / (p.xi2 * E);   // :129
  p.c_scale = p.xi2 * E / (4.0 * m->kappa);                            // :130
  hh_bk_config c;
  if (cfg && cfg->n_std > 0) c = *cfg;
  else hh_default_bk_config(&c);
  p.h_fd = fabs(c.h_fd);
  p.widen_fd = c.h_fd > 0.0;  // h_fd < 0: the reference's plain finite differences at |h_fd| whatever their noise
  p.noise_scale = 8.9e-16 / (p.h_fd * p.h_fd) / 0.02;
  p.cf_tol = c.cf_tol;
  p.atol = c.atol;
  p.n_std = c.n_std;
  p.max_terms = c.max_terms > 0 ? c.max_terms : 4096;
  if (p.max_terms > 4096) p.max_terms = 4096;
  p.r_tau = m->r * tau;
  p.kappa_theta_tau = m->kappa * m->theta * tau;
  p.rho_over_xi = m->rho / m->xi;
  p.one_m_rho2 = 1.0 - m->rho * m->rho;
  p.rho = m->rho;
  p.theta = m->theta;
  return HH_OK;
}

static int ensure_slab(hh_ctx *ctx, int nblocks, const BkParams &p, double **slab, int64_t *stride) {
  const int64_t threads = (int64_t)nblocks * kBkThreads;
  const int64_t extra = p.max_terms > kBkTable ? p.max_terms - kBkTable : 0;
  HH_CUDA(ctx, ctx->d_bk_slab.ensure(sizeof(double) * (size_t)(threads * extra + 1)));
  *slab = ctx->d_bk_slab.as<double>();
  *stride = threads;
  return HH_OK;
}

static int bk_set_smem(hh_ctx *ctx) {
  static PerDeviceOnce done;
  if (done.done(ctx->device)) return HH_OK;
  const int bytes = kBkThreads * kBkTable * (int)sizeof(double);
  HH_CUDA(ctx, cudaFuncSetAttribute(bk_integral_sorted_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  HH_CUDA(ctx, cudaFuncSetAttribute(bk_integral_sorted_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  HH_CUDA(ctx, cudaFuncSetAttribute(bk_integral_sorted_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  HH_CUDA(ctx, cudaFuncSetAttribute(bk_integral_sorted_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  HH_CUDA(ctx, cudaFuncSetAttribute(bk_integral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.set(ctx->device);
  return HH_OK;
}

// Runs the pipeline for `s` (validated by the caller); terminal spots land in ctx->d_terminal, path statistics in
// `d_stats` when non-null. Records ev0 before the first kernel. Large jobs are cut into chunks of trajectories so that the
// work arrays (36 B per transition) stay below ~5 GB.
static int bk_paths_launch(hh_ctx *ctx, const hh_model *m, const hh_sim *s, double *d_stats, int monitor_every,
                           double *d_grid = nullptr, int64_t grid_stride = 0) {
  if (s->rng_mode != HH_RNG_PHILOX)
    return ctx->fail(HH_ERR_UNSUPPORTED, "Broadie-Kaya draws from the in-kernel Philox stream only; the deterministic "
                                         "pieces have their own parity probes (hh_bk_chf, hh_bk_integral)");
  if (s->precision != HH_PREC_F64) return ctx->fail(HH_ERR_UNSUPPORTED, "Broadie-Kaya runs in f64 only");
  const int ndates = s->n_steps > 0 ? s->n_steps : 1;
  BkParams p;
  int rc = make_params(ctx, m, m->T / ndates, &s->bk, p);
  if (rc) return rc;
  const int64_t N = s->n_paths;
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  rc = bk_set_smem(ctx);
  if (rc) return rc;
  const int smem = kBkThreads * kBkTable * (int)sizeof(double);
  static const int minb_env = getenv("HH_BK_MINB") ? atoi(getenv("HH_BK_MINB")) : kBkDefaultMinb;
  const int minb = minb_env < 3 ? 3 : (minb_env > 6 ? 6 : minb_env);
  void (*kern)(const BkIntArgs) = minb == 3   ? bk_integral_sorted_kernel<3>
                                  : minb == 4 ? bk_integral_sorted_kernel<4>
                                  : minb == 5 ? bk_integral_sorted_kernel<5>
                                              : bk_integral_sorted_kernel<6>;
  int occ = 1;
  HH_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBkThreads, smem));
  if (occ < 1) occ = 1;
  // chunks of trajectories: at most 2^27 transitions each
  // (HH_BK_MAX_ITEMS: test hook that forces small chunks)
  static const int64_t max_items_env = getenv("HH_BK_MAX_ITEMS") ? atoll(getenv("HH_BK_MAX_ITEMS")) : 0;
  const int64_t max_items = max_items_env > 0 ? max_items_env : (int64_t)1 << 27;
  int64_t chunk = max_items / ndates;
  if (chunk < 1) chunk = 1;
  if (chunk > N) chunk = N;
  const int64_t chunk_items = chunk * ndates;
  int64_t igrid = (int64_t)ctx->sm_count * occ;
  const int64_t want_blocks = (chunk_items + kBkThreads - 1) / kBkThreads;
  if (igrid > want_blocks) igrid = want_blocks;
  BkIntArgs ia;
  memset(&ia, 0, sizeof ia);
  rc = ensure_slab(ctx, (int)igrid, p, &ia.slab, &ia.slab_stride);
  if (rc) return rc;
  // work arrays: V (ndates + 1) x chunk, U, Z, I: ndates x chunk (f64), perm: ndates x chunk (u32), histogram
  const size_t nV = (size_t)(ndates + 1) * (size_t)chunk, nI = (size_t)chunk_items;
  HH_CUDA(ctx, ctx->d_bk_work.ensure(sizeof(double) * (nV + 3 * nI) + sizeof(unsigned) * (nI + kBkBuckets) + 64));
  double *dV = ctx->d_bk_work.as<double>(), *dU = dV + nV, *dZ = dU + nI, *dI = dZ + nI;
  unsigned *dperm = reinterpret_cast<unsigned *>(dI + nI), *dhist = dperm + nI;
  HH_CUDA(ctx, ctx->d_terminal.ensure(sizeof(double) * (size_t)N));
  HH_CUDA(ctx, ctx->d_counters.ensure(sizeof(unsigned long long) * 8));
  HH_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.ptr, 0, sizeof(unsigned long long) * 8, st));
  if (s->seeds) {
    const size_t bytes = sizeof(uint64_t) * (size_t)N;
    HH_CUDA(ctx, ctx->d_seeds.ensure(bytes));
    HH_CUDA(ctx, cudaMemcpyAsync(ctx->d_seeds.ptr, s->seeds, bytes, cudaMemcpyHostToDevice, st));
  }
  HH_CUDA(ctx, upload_fast_tables(ctx->device, st));
  HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  const int mon = monitor_every > 0 ? monitor_every : 1;
  for (int64_t c0 = 0; c0 < N; c0 += chunk) {
    const int64_t cn = N - c0 < chunk ? N - c0 : chunk, items = cn * ndates;
    const int64_t path_blocks = (cn + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    HH_CUDA(ctx, cudaMemsetAsync(dhist, 0, sizeof(unsigned) * kBkBuckets, st));
    BkChainArgs ca;
    memset(&ca, 0, sizeof ca);
    ca.n = cn;
    ca.path_offset = s->path_offset + c0;
    ca.base_seed = s->base_seed;
    ca.seeds = s->seeds ? ctx->d_seeds.as<uint64_t>() + c0 : nullptr;
    ca.n_dates = ndates;
    ca.p = p;
    ca.v0 = m->V0;
    ca.V = dV;
    ca.U = dU;
    ca.Z = dZ;
    ca.hist = dhist;
    bk_chain_kernel<<<(unsigned)(path_blocks < cap ? path_blocks : cap), 256, 0, st>>>(ca);
    bk_scan_kernel<<<1, kBkBuckets, 0, st>>>(dhist);
    const int64_t tiles = (items + kBkScatterTile - 1) / kBkScatterTile;
    bk_scatter_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, st>>>(dV, cn, items, dhist, dperm);
    ia.items = items;
    ia.n = cn;
    ia.p = p;
    ia.V = dV;
    ia.U = dU;
    ia.perm = dperm;
    ia.I = dI;
    ia.counters = ctx->d_counters.as<unsigned long long>();
    const int64_t wb = (items + kBkThreads - 1) / kBkThreads;
    kern<<<(unsigned)(igrid < wb ? igrid : wb), kBkThreads, smem, st>>>(ia);
    BkAsmArgs aa;
    memset(&aa, 0, sizeof aa);
    aa.n = cn;
    aa.n_total = N;
    aa.n_dates = ndates;
    aa.p = p;
    aa.x0 = log(m->S0);
    aa.s0 = m->S0;
    aa.V = dV;
    aa.Z = dZ;
    aa.I = dI;
    aa.terminal = ctx->d_terminal.as<double>() + c0;
    aa.grid = d_grid ? d_grid + c0 : nullptr;
    aa.grid_stride = grid_stride;
    aa.stats = d_stats ? d_stats + c0 : nullptr;
    aa.monitor_every = mon;
    aa.inv_m = 1.0 / (double)(ndates / mon > 0 ? ndates / mon : 1);
    bk_assemble_kernel<<<(unsigned)(path_blocks < cap ? path_blocks : cap), 256, 0, st>>>(aa);
    HH_CUDA(ctx, cudaGetLastError());
  }
  return HH_OK;
}

int bk_european_launch(hh_ctx *ctx, const hh_model *m, const hh_sim *s, const hh_payoff *payoffs, int npay,
                       int want_terminal) {
  if (npay < 1 || npay > 256 || !payoffs) return ctx->fail(HH_ERR_ARG, "npayoffs must be in [1, 256] (got %d)", npay);
  int rc = bk_paths_launch(ctx, m, s, nullptr, 1);
  if (rc) return rc;
  rc = terminal_payoffs_launch(ctx, ctx->d_terminal.as<double>(), s->n_paths, payoffs, npay, 0);
  if (rc) return rc;
  ctx->pend.want_terminal = want_terminal != 0;
  ctx->pend.bk = true;
  ctx->pend.nseg = 0;
  return HH_OK;
}

// Path statistics under exact Broadie-Kaya transitions for hh_mc_path_dependent (hh_pathdep.cu): HH_PD_NSTATS x n_paths
// into `d_stats` (device), the inversion counters into ctx->bk_stats after the caller has synchronised and called
// bk_read_counters.
int bk_path_stats_launch(hh_ctx *ctx, const hh_model *m, const hh_sim *s, int monitor_every, double *d_stats) {
  return bk_paths_launch(ctx, m, s, d_stats, monitor_every);
}

int bk_path_grid_launch(hh_ctx *ctx, const hh_model *m, const hh_sim *s, double *d_grid, int64_t grid_stride) {
  return bk_paths_launch(ctx, m, s, nullptr, 1, d_grid, grid_stride);
}

int bk_read_counters(hh_ctx *ctx, int64_t *n_fallback) {  // after the stream has been synchronised
  unsigned long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  HH_CUDA(ctx, cudaMemcpy(counters, ctx->d_counters.ptr, sizeof counters, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 5; ++i) ctx->bk_stats[i] = (double)counters[i];
  if (n_fallback) *n_fallback = (int64_t)counters[0];
  return HH_OK;
}

int bk_chf(hh_ctx *ctx, const hh_model *m, double tau, const double *V0, const double *VT, int n, const double *a, int na,
           double *out_re, double *out_im) {
  if (!m || !V0 || !VT || !a || !out_re || !out_im || n < 1 || na < 1) return ctx->fail(HH_ERR_ARG, "hh_bk_chf: bad argument");
  BkParams p;
  int rc = make_params(ctx, m, tau, nullptr, p);
  if (rc) return rc;
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nv = sizeof(double) * (size_t)n, nm = sizeof(double) * (size_t)n * na;
  HH_CUDA(ctx, ctx->d_misc.ensure(2 * nv + 3 * nm));
  double *dV0 = ctx->d_misc.as<double>(), *dVT = dV0 + n, *da = dVT + n, *dre = da + (size_t)n * na, *dim_ = dre + (size_t)n * na;
  HH_CUDA(ctx, cudaMemcpyAsync(dV0, V0, nv, cudaMemcpyHostToDevice, st));
  HH_CUDA(ctx, cudaMemcpyAsync(dVT, VT, nv, cudaMemcpyHostToDevice, st));
  HH_CUDA(ctx, cudaMemcpyAsync(da, a, nm, cudaMemcpyHostToDevice, st));
  bk_chf_kernel<<<(n + 63) / 64, 64, 0, st>>>(p, dV0, dVT, n, da, na, dre, dim_);
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaMemcpyAsync(out_re, dre, nm, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaMemcpyAsync(out_im, dim_, nm, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  return HH_OK;
}

int bk_log_besseli(hh_ctx *ctx, double nu, const double *zr, const double *zi, int n, double *out_re, double *out_im) {
  if (!zr || !zi || !out_re || !out_im || n < 1) return ctx->fail(HH_ERR_ARG, "hh_bk_log_besseli: bad argument");
  if (!(nu > -1.0)) return ctx->fail(HH_ERR_ARG, "order must be > -1 (got %g)", nu);
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nv = sizeof(double) * (size_t)n;
  HH_CUDA(ctx, ctx->d_misc.ensure(4 * nv));
  double *dzr = ctx->d_misc.as<double>(), *dzi = dzr + n, *dre = dzi + n, *dim_ = dre + n;
  HH_CUDA(ctx, cudaMemcpyAsync(dzr, zr, nv, cudaMemcpyHostToDevice, st));
  HH_CUDA(ctx, cudaMemcpyAsync(dzi, zi, nv, cudaMemcpyHostToDevice, st));
  bk_log_besseli_kernel<<<(n + 127) / 128, 128, 0, st>>>(make_bessel_order(nu), dzr, dzi, n, dre, dim_);
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaMemcpyAsync(out_re, dre, nv, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaMemcpyAsync(out_im, dim_, nv, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  return HH_OK;
}

int bk_integral(hh_ctx *ctx, const hh_model *m, double tau, const hh_bk_config *cfg, const double *V0, const double *VT,
                const double *U, int n, double *out8) {
  if (!m || !V0 || !VT || !U || !out8 || n < 1) return ctx->fail(HH_ERR_ARG, "hh_bk_integral: bad argument");
  BkParams p;
  int rc = make_params(ctx, m, tau, cfg, p);
  if (rc) return rc;
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  rc = bk_set_smem(ctx);
  if (rc) return rc;
  const int nblocks = (n + kBkThreads - 1) / kBkThreads;
  double *slab;
  int64_t stride;
  rc = ensure_slab(ctx, nblocks, p, &slab, &stride);
  if (rc) return rc;
  const size_t nv = sizeof(double) * (size_t)n;
  HH_CUDA(ctx, ctx->d_misc.ensure(3 * nv + 8 * nv));
  double *dV0 = ctx->d_misc.as<double>(), *dVT = dV0 + n, *dU = dVT + n, *dout = dU + n;
  HH_CUDA(ctx, cudaMemcpyAsync(dV0, V0, nv, cudaMemcpyHostToDevice, st));
  HH_CUDA(ctx, cudaMemcpyAsync(dVT, VT, nv, cudaMemcpyHostToDevice, st));
  HH_CUDA(ctx, cudaMemcpyAsync(dU, U, nv, cudaMemcpyHostToDevice, st));
  const int smem = kBkThreads * kBkTable * (int)sizeof(double);
  bk_integral_kernel<<<nblocks, kBkThreads, smem, st>>>(p, dV0, dVT, dU, n, dout, slab, stride);
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaMemcpyAsync(out8, dout, 8 * nv, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  return HH_OK;
}

int bk_elementary(hh_ctx *ctx, int kind, const double *x, const double *y, int n, double *out_a, double *out_b) {
  if (!x || !out_a || n < 1 || kind < 0 || kind > 3 || (kind == 3 && !y) || (kind == 1 && !out_b))
    return ctx->fail(HH_ERR_ARG, "hh_bk_elementary: bad argument");
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nv = sizeof(double) * (size_t)n;
  HH_CUDA(ctx, ctx->d_misc.ensure(4 * nv));
  double *dx = ctx->d_misc.as<double>(), *dy = dx + n, *da = dy + n, *db = da + n;
  HH_CUDA(ctx, cudaMemcpyAsync(dx, x, nv, cudaMemcpyHostToDevice, st));
  if (y) HH_CUDA(ctx, cudaMemcpyAsync(dy, y, nv, cudaMemcpyHostToDevice, st));
  bk_elementary_kernel<<<(n + 127) / 128, 128, 0, st>>>(kind, dx, dy, n, da, db);
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaMemcpyAsync(out_a, da, nv, cudaMemcpyDeviceToHost, st));
  if (out_b) HH_CUDA(ctx, cudaMemcpyAsync(out_b, db, nv, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  return HH_OK;
}

int bk_variance(hh_ctx *ctx, const hh_model *m, double tau, const double *V0, int n, uint64_t seed, double *VT) {
  if (!m || !V0 || !VT || n < 1) return ctx->fail(HH_ERR_ARG, "hh_bk_variance: bad argument");
  BkParams p;
  int rc = make_params(ctx, m, tau, nullptr, p);
  if (rc) return rc;
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nv = sizeof(double) * (size_t)n;
  HH_CUDA(ctx, ctx->d_misc.ensure(2 * nv));
  double *dV0 = ctx->d_misc.as<double>(), *dVT = dV0 + n;
  HH_CUDA(ctx, upload_fast_tables(ctx->device, st));
  HH_CUDA(ctx, cudaMemcpyAsync(dV0, V0, nv, cudaMemcpyHostToDevice, st));
  bk_variance_kernel<<<(n + 127) / 128, 128, 0, st>>>(p, dV0, n, seed, dVT);
  HH_CUDA(ctx, cudaGetLastError());
  HH_CUDA(ctx, cudaMemcpyAsync(VT, dVT, nv, cudaMemcpyDeviceToHost, st));
  HH_CUDA(ctx, cudaStreamSynchronize(st));
  return HH_OK;
}

}  // namespace hh
