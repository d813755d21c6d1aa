// hh_bessel.cuh — log I_nu(z) for complex z and real order nu > -1, and the Broadie-Kaya characteristic function of
// the integrated variance that needs it. Host + device code (the host build is only used by tools/bk_host_check.cu to
// validate the arithmetic on a CPU-only box; the product calls the device build).
//
// Replaces SpecialFunctions.besseli(nu, z) [AMOS zbesi] at the reference's call sites
//   src/distributions/heston.jl:173  log(besseli(nu, z_kappa))          real z
//   src/distributions/heston.jl:207  log(besseli(nu, z_gamma_unwrapped)) complex z
// The LOG is returned directly, so arguments the reference overflows on (|Re z| > 709) stay finite here.
//
// Method (w = z reflected into Re w >= 0;  I_nu(z) = exp(+-i pi nu) I_nu(-z) otherwise):
//   |w| <= 5, or |w| < r_asym and |w| - Re w <= 5
//                        ascending series  (w/2)^nu sum_k (w^2/4)^k / (k! Gamma(nu+k+1)): the cancellation in the sum is
//                        I_nu(|w|) / |I_nu(w)| ~ e^{|w| - Re w} <= e^5, whatever |w| (no cancellation at all on the real
//                        axis), so arguments near the real axis never need the continued fractions (~50x dearer)
//   |w| >= 20 + nu^2/2   Hankel expansion with BOTH exponentials (DLMF 10.40.5)
//   nu >= 11.5, between the two (the series would need more than its 100 tabulated terms, the Hankel terms grow first):
//                        Debye's uniform expansion in the order (DLMF 10.41.3, u_0..u_8) in a sector round the real axis —
//                        low vol of vol means a large order (nu = 2 kappa theta / sigma^2 - 1: 15 at sigma = 0.1, 89 at
//                        kappa theta = 0.45), and |z| there is a few times nu
//   otherwise            continued fractions: CF1 for I'/I (modified Lentz), Steed's CF2 for K_mu, K_mu+1, |mu| <= 1/2,
//                        and the Wronskian I K' - I' K = -1/w (Temme 1975; Thompson & Barnett 1987 for complex w);
//                        orders in (-1, 0) are reached from nu + 1 through I_(nu) = I'_(nu+1) + ((nu+1)/w) I_(nu+1).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define HH_HD __host__ __device__ __forceinline__
// One copy per kernel instead of one per call site: the Broadie-Kaya kernel inlined to 16 000 instructions (256 KB), and
// ncu showed instruction fetch ("no_instruction") as its first stall reason, ahead of every data dependency.
#define HH_HD_OUTLINE inline __host__ __device__ __noinline__
#else
#define HH_HD inline
#define HH_HD_OUTLINE inline
#endif

namespace hh {

struct cplx {
  double re, im;
};
HH_HD cplx mk(double re, double im = 0.0) { return cplx{re, im}; }
HH_HD cplx operator+(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
HH_HD cplx operator-(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
HH_HD cplx operator-(cplx a) { return cplx{-a.re, -a.im}; }
HH_HD cplx operator*(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
HH_HD cplx operator*(double s, cplx a) { return cplx{s * a.re, s * a.im}; }
HH_HD cplx operator*(cplx a, double s) { return cplx{s * a.re, s * a.im}; }
HH_HD cplx operator+(cplx a, double s) { return cplx{a.re + s, a.im}; }
HH_HD cplx operator+(double s, cplx a) { return cplx{a.re + s, a.im}; }
HH_HD cplx operator-(cplx a, double s) { return cplx{a.re - s, a.im}; }
HH_HD cplx operator-(double s, cplx a) { return cplx{s - a.re, -a.im}; }
HH_HD cplx conj(cplx a) { return cplx{a.re, -a.im}; }
HH_HD double cabs2(cplx a) { return a.re * a.re + a.im * a.im; }
HH_HD double cabs(cplx a) { return sqrt(a.re * a.re + a.im * a.im); }  // magnitudes here are far from over/underflow
// 1 / x to ~1 ulp for normal-range x: on the device MUFU.RCP64H + two Newton steps instead of the IEEE division
// sequence (the Bessel series and the complex divisions below are dependent chains of them)
HH_HD double rcp_fast(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}

// sqrt x for x well inside the normal range (the moduli here): on the device the MUFU.RSQ64H seed (2^-22.9) and one cubic
// step, s = s0 (1 + e + 3/2 e^2) with s0 = x y0, e = (1 - x y0^2) / 2 — 5 FP64 instructions instead of IEEE sqrt's ~14, ~1 ulp
#ifdef __CUDA_ARCH__
// the IEEE square root (an inline expansion of ~40 instructions with its own slow path) is kept OUT of the hot functions:
// ncu shows ~1 cycle of instruction-fetch stall per issued instruction in the inversion kernel, and more than half of the
// 1500 instructions of bk_chf were never-executed fallbacks interleaved with the hot path
static __device__ __noinline__ double sqrt_slow(double x) { return sqrt(x); }
#endif
HH_HD double sqrt_fast(double x) {
#ifdef __CUDA_ARCH__
  const int hi = __double2hiint(x);
  if (hi < 0x00200000 || hi >= 0x7fd00000) return sqrt_slow(x);
  {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double h0 = __hiloint2double(__double2hiint(y0) - 0x00100000, 0);  // y0 / 2 (the seed's low word is zero)
    const double s0 = x * y0;
    const double e = fma(-s0, h0, 0.5);
    const double pp = fma(e * 1.5, e, e);
    return fma(s0, pp, s0);
  }
#else
  return sqrt(x);
#endif
}

// 1 / z = conj(z) / |z|^2 for moduli far from the over/underflow of the square (one reciprocal; Smith's form takes two)
HH_HD cplx crecip(cplx z) {
  const double id = rcp_fast(z.re * z.re + z.im * z.im);
  return cplx{z.re * id, -(z.im * id)};
}

HH_HD cplx operator/(cplx a, cplx b) {
  // Smith's algorithm (no spurious overflow), with reciprocals instead of divisions
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im * rcp_fast(b.re), id = rcp_fast(b.re + b.im * r);
    return cplx{(a.re + a.im * r) * id, (a.im - a.re * r) * id};
  }
  const double r = b.re * rcp_fast(b.im), id = rcp_fast(b.re * r + b.im);
  return cplx{(a.re * r + a.im) * id, (a.im * r - a.re) * id};
}
HH_HD cplx operator/(double s, cplx b) { return mk(s) / b; }
HH_HD cplx operator/(cplx a, double s) {
  const double is = rcp_fast(s);
  return cplx{a.re * is, a.im * is};
}

// ---- elementary functions of the hot path --------------------------------------------------------------------------------
// ncu on the Broadie-Kaya kernel put 55-60 % of the executed instructions inside the CUDA math library's exp / sincos / log /
// atan2 (~125 instructions per call, serial Horner chains, slow-path tests). The versions below read small shared-memory
// tables (11 KB per block, filled by bk_fill_fast_tables) and finish with short polynomials: exp 10, sincos ~20, log ~11,
// atan2 ~26 FP64 instructions, branch free on the fast path, |error| <= ~2e-16 (absolute for log / atan2 / sincos, relative
// for exp). With a null table pointer (host build, tools) they are the library functions.
struct alignas(16) BkFastTables {
  double exp2t[256];  // 2^(j/256)
  double logt[512];   // {r_j, -log r_j}, r_j = 1 / (1 + (j + 1/2) / 256) rounded to double
  double trig[512];   // {cos, sin}(2 pi j / 256)
  double atant[130];  // atan(j / 128), j = 0..128
};
constexpr unsigned kFtExp = 0, kFtLog = 2048, kFtTrig = 2048 + 4096, kFtAtan = 2048 + 4096 + 4096;  // byte offsets

// Shared-window address of a block's tables (0: none — the host build, and callers without tables). The tables are read
// with ld.shared through this address: a generic pointer handed to an out-of-line function compiles to LD.E (generic
// loads, ~10 cycles more than LDS and on the global-memory path), which ncu showed as the first stall of the series loops.
struct FastRef {
  unsigned saddr;
};

#ifdef __CUDACC__
__device__ __forceinline__ double lds_f64(unsigned saddr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void lds_f64x2(unsigned saddr, double &a, double &b) {
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(saddr));
}
#endif

#ifdef __CUDACC__
// all threads of the block; the caller synchronises
__device__ __forceinline__ void bk_fill_fast_tables(BkFastTables *t) {
  for (int j = threadIdx.x; j < 256; j += blockDim.x) {
    t->exp2t[j] = exp2((double)j * (1.0 / 256.0));
    const double r = 1.0 / (1.0 + ((double)j + 0.5) * (1.0 / 256.0));
    t->logt[2 * j] = r;
    t->logt[2 * j + 1] = -log(r);
    double sn, cs;
    sincospi((double)j * (1.0 / 128.0), &sn, &cs);
    t->trig[2 * j] = cs;
    t->trig[2 * j + 1] = sn;
  }
  for (int j = threadIdx.x; j <= 128; j += blockDim.x) t->atant[j] = atan((double)j * (1.0 / 128.0));
}
#endif

#ifdef __CUDACC__
// Constants that need all 64 bits live in constant memory: a DFMA takes them as a c[bank][offset] operand, where an
// immediate costs two UMOV / MOV instructions per use (ncu: 56 UMOV + 58 IMAD.MOV of the 518 straight-line instructions of
// one characteristic-function evaluation).
struct BkConsts {
  double inv_ln2_256, ln2_256_hi, ln2_256_lo, e4, e3;            // fexp
  double inv_2pi_256, twopi_256_hi, twopi_256_lo, s7, s5, s3, c6, c4;  // fsincos
  double l6, l5, l3, ln2_hi, ln2_lo;                              // flog
  double a7, a5, a3, half_pi, pi;                                 // fatan2
};
static __constant__ BkConsts kBkC = {
    369.3299304675746, -2.7076061740622863e-03, -9.058776616587108e-20, 4.1666666666666664e-02, 1.6666666666666666e-01,
    40.74366543152521, -2.454369260617026e-02, -9.567553118338697e-19, -1.9841269841269841e-04, 8.3333333333333332e-03,
    -1.6666666666666666e-01, -1.3888888888888889e-03, 4.1666666666666664e-02,
    -1.6666666666666666e-01, 0.2, 3.3333333333333331e-01, 6.9314718055994529e-01, 2.3190468138462996e-17,
    -1.4285714285714285e-01, 0.2, -3.3333333333333331e-01, 1.5707963267948966, 3.1415926535897931};
#endif

#ifdef __CUDA_ARCH__
// library fallbacks (arguments outside the fast paths, callers without tables): out of line, so that the inline expansions of
// the CUDA math library (~140 instructions per exp + sincos) do not sit between the hot instructions of every call site
static __device__ __noinline__ double fexp_slow(double x) { return exp(x); }
static __device__ __noinline__ void fsincos_slow(double x, double *s, double *c) { sincos(x, s, c); }
static __device__ __noinline__ double flog_slow(double x) { return log(x); }
static __device__ __noinline__ double fatan2_slow(double y, double x) { return atan2(y, x); }
#endif

// exp(x): x = (256 k + j) ln2 / 256 + r, exp(x) = 2^k T_j exp(r), |r| <= ln2 / 512
HH_HD double fexp(FastRef ft, double x) {
#ifdef __CUDA_ARCH__
  if (ft.saddr) {
    if ((unsigned)(__double2hiint(x) & 0x7fffffff) >= 0x4085E000u) return fexp_slow(x);  // |x| >= 700, inf, nan
    const double magic = 6755399441055744.0;                    // 1.5 2^52: the nearest integer lands in the low word
    const double t = fma(x, kBkC.inv_ln2_256, magic);           // 256 / ln 2
    const int n = __double2loint(t);
    const double nf = t - magic;
    double r = fma(nf, kBkC.ln2_256_hi, x);                     // -(ln 2 / 256), high part
    r = fma(nf, kBkC.ln2_256_lo, r);                            // low part  (hi + lo = ln2/256 to 1e-36)
    const double e = lds_f64(ft.saddr + kFtExp + 8u * (unsigned)(n & 255));
    double q = fma(r, kBkC.e4, kBkC.e3);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double v = fma(e * r, q, e);
    return __hiloint2double(__double2hiint(v) + ((n >> 8) << 20), __double2loint(v));
  }
  return fexp_slow(x);
#else
  (void)ft;
  return exp(x);
#endif
}

// sin x, cos x: x = (256 k + j) 2 pi / 256 + r, |r| <= pi / 256; rotation of the tabulated (cos, sin) by r
HH_HD void fsincos(FastRef ft, double x, double &sn, double &cs) {
#ifdef __CUDA_ARCH__
  if (ft.saddr) {
    if ((unsigned)(__double2hiint(x) & 0x7fffffff) >= 0x41300000u) {  // |x| >= 2^20, inf, nan
      fsincos_slow(x, &sn, &cs);
      return;
    }
    const double magic = 6755399441055744.0;
    const double t = fma(x, kBkC.inv_2pi_256, magic);         // 256 / (2 pi)
    const int n = __double2loint(t);
    const double nf = t - magic;
    double r = fma(nf, kBkC.twopi_256_hi, x);                   // -(2 pi / 256), high part
    r = fma(nf, kBkC.twopi_256_lo, r);                          // low part
    double2 cs0;
    lds_f64x2(ft.saddr + kFtTrig + 16u * (unsigned)(n & 255), cs0.x, cs0.y);
    const double r2 = r * r;
    double ps = fma(r2, kBkC.s7, kBkC.s5);
    ps = fma(ps, r2, kBkC.s3);
    const double ds = (r * r2) * ps;                            // sin r - r
    double pc = fma(r2, kBkC.c6, kBkC.c4);
    pc = fma(pc, r2, -0.5);
    const double dc = r2 * pc;                                  // cos r - 1
    const double sr = r + ds;
    cs = cs0.x + fma(cs0.x, dc, -(cs0.y * sr));
    sn = cs0.y + fma(cs0.y, dc, cs0.x * sr);
    return;
  }
  fsincos_slow(x, &sn, &cs);
#else
  (void)ft;
  sn = sin(x);
  cs = cos(x);
#endif
}

// log x for normal x > 0: x = 2^e m, m r_j = 1 + s with |s| <= 2^-9: log x = e ln2 - log r_j + log1p(s)
HH_HD double flog(FastRef ft, double x) {
#ifdef __CUDA_ARCH__
  if (ft.saddr) {
    const int hi = __double2hiint(x);
    if (hi >= 0x00100000 && hi < 0x7ff00000) {
      const int e = (hi >> 20) - 1023, j = (hi >> 12) & 255;
      const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
      double2 rl;
      lds_f64x2(ft.saddr + kFtLog + 16u * (unsigned)j, rl.x, rl.y);
      const double sft = fma(m, rl.x, -1.0);
      double q = fma(sft, kBkC.l6, kBkC.l5);
      q = fma(q, sft, -0.25);
      q = fma(q, sft, kBkC.l3);
      q = fma(q, sft, -0.5);
      const double s2 = sft * sft;
      const double l1p = fma(s2, q, sft);
      const double ef = (double)e;
      return fma(ef, kBkC.ln2_hi, rl.y + fma(ef, kBkC.ln2_lo, l1p));
    }
  }
  return flog_slow(x);
#else
  (void)ft;
  return log(x);
#endif
}

// atan2(y, x) in (-pi, pi]: t = min/max in [0, 1], c = nearest multiple of 1/128, atan t = atan c + atan((t - c)/(1 + t c))
HH_HD double fatan2(FastRef ft, double y, double x) {
#ifdef __CUDA_ARCH__
  if (ft.saddr) {
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    const int hm = __double2hiint(mx);
    if (hm >= 0x00200000 && hm < 0x7fd00000) {  // otherwise (0, tiny, huge, inf, nan): the library
      double ir;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ir) : "d"(mx));
      const double magic = 6755399441055744.0;
      const double tt = fma(mn * ir, 128.0, magic);
      const int j = __double2loint(tt);
      const double c = (tt - magic) * 0.0078125;
      const double num = fma(-c, mx, mn), den = fma(c, mn, mx);
      double yd;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(yd) : "d"(den));
      double e = fma(-den, yd, 1.0);
      yd = fma(yd, e, yd);
      e = fma(-den, yd, 1.0);
      yd = fma(yd, e, yd);
      const double u = num * yd, u2 = u * u;
      double q = fma(u2, kBkC.a7, kBkC.a5);
      q = fma(q, u2, kBkC.a3);
      double a = lds_f64(ft.saddr + kFtAtan + 8u * (unsigned)j) + fma(u * u2, q, u);
      if (ay > ax) a = kBkC.half_pi - a;
      if (x < 0.0) a = kBkC.pi - a;
      return copysign(a, y);
    }
  }
  return fatan2_slow(y, x);
#else
  (void)ft;
  return atan2(y, x);
#endif
}

HH_HD cplx cexp_(FastRef ft, cplx a) {
  const double e = fexp(ft, a.re);
  double s, c;
  fsincos(ft, a.im, s, c);
  return cplx{e * c, e * s};
}
// log|a| = log(re^2 + im^2) / 2: no hypot, no sqrt (the arguments here are far from the over/underflow of the squares)
HH_HD cplx clog_(FastRef ft, cplx a) { return cplx{0.5 * flog(ft, cabs2(a)), fatan2(ft, a.im, a.re)}; }
HH_HD cplx csqrt_(cplx a) {
  // principal branch, Re >= 0
  const double m = sqrt_fast(cabs2(a));
  if (m == 0.0) return cplx{0.0, 0.0};
  if (a.re >= 0.0) {
    const double t = sqrt_fast(0.5 * (m + a.re));
    return cplx{t, 0.5 * a.im * rcp_fast(t)};
  }
  const double t = sqrt_fast(0.5 * (m - a.re));
  return cplx{0.5 * fabs(a.im) * rcp_fast(t), a.im >= 0.0 ? t : -t};
}

// sqrt(a) for Re a > 0 (gamma^2 = kappa^2 - 2 sigma^2 a i in the characteristic function): the first branch of csqrt_ alone
HH_HD cplx csqrt_pos_(cplx a) {
  const double m = sqrt_fast(cabs2(a));
  const double t = sqrt_fast(0.5 * (m + a.re));
  return cplx{t, 0.5 * a.im * rcp_fast(t)};
}

constexpr double kBesselPi = 3.14159265358979323846;

// I_nu = exp(E) F: the series and the Hankel expansion produce an algebraic factor F (the sum) next to an exponent that
// only needs log|w| and arg w, which the caller often has already (bk_chf unwraps arg z_gamma anyway). Returning
// (E, F) instead of log I avoids a complex logarithm of F per evaluation, and the caller exponentiates once.
struct BesselEF {
  cplx E, F;
};

constexpr int kSeriesMaxTerms = 100, kHankelMaxTerms = 60;
// Orders from which Debye's expansion fills the gap between the series and the Hankel expansion. Below it the tabulated
// series reaches r_asym (it does up to nu = 11.9), and r_near == r_asym exactly: the hot region test relies on that.
constexpr double kDebyeMinOrder = 11.5;

// Host-prepared constants of one order, plus the shared-window addresses of the block's coefficient tables (device).
struct BesselOrder {
  double nu;        // order, > -1
  double lgam_nu1;  // lgamma(nu + 1)
  double r_asym;    // |w| from which the Hankel expansion is used
  double r_near;    // min(r_asym, largest |w| whose series reaches 1e-17 within the tabulated terms)
  double r_asym2, r_near2;  // their squares (region tests on |w|^2)
  // device: tables filled by bk_fill_order_tables in shared memory (the order is fixed per launch), 0 on the host
  //   series: groups of four, R_{k,i} = prod_{m=k}^{k+i-1} 1 / (m (nu + m)), k = 1, 5, 9, ...
  //   hankel: groups of four running products b_k, b_k b_{k+1}, ..., b_k = (4 nu^2 - (2k - 1)^2) / (8 k), k = 1, 5, 9, ...
  unsigned series_saddr, hankel_saddr;
  FastRef ft;       // tables of the elementary functions (saddr 0: the math library)
  unsigned pad_;
};
// Largest |w| <= r_hi for which the ascending series reaches its 1e-17 test within 92 terms (the device sums groups of
// four up to 100): term_N / term_peak <= 1e-17 e^{-6}, the e^{-6} for the cancellation allowed off the real axis
// (|w| - Re w <= 5) and for the peak term against the sum. log term_k = k log q - lgamma(k+1) - lgamma(nu+k+1) + const.
HH_HD bool bessel_series_enough(double nu, double r) {
  const double N = 92.0;
  const double q = 0.25 * r * r, lq = log(q);
  const double kp = fmax(0.0, 0.5 * (sqrt(nu * nu + 4.0 * q) - nu));  // k (k + nu) = q
  if (kp >= N - 1.0) return false;
  const double tn = N * lq - lgamma(N + 1.0) - lgamma(nu + N + 1.0);
  const double tp = kp * lq - lgamma(kp + 1.0) - lgamma(nu + kp + 1.0);
  return tn - tp <= -39.14394658089878 - 6.0;  // log 1e-17
}
HH_HD double bessel_series_radius(double nu, double r_hi) {
  if (r_hi <= 5.0 || bessel_series_enough(nu, r_hi)) return r_hi;
  double lo = 5.0, hi = r_hi;
  for (int i = 0; i < 60; ++i) {
    const double mid = 0.5 * (lo + hi);
    if (bessel_series_enough(nu, mid)) lo = mid; else hi = mid;
  }
  return lo;
}
HH_HD BesselOrder make_bessel_order(double nu) {
  BesselOrder o;
  o.nu = nu;
  o.lgam_nu1 = lgamma(nu + 1.0);
  o.r_asym = 20.0 + 0.5 * nu * nu;
  o.r_near = nu < kDebyeMinOrder ? o.r_asym : bessel_series_radius(nu, o.r_asym);
  o.r_asym2 = o.r_asym * o.r_asym;
  o.r_near2 = o.r_near * o.r_near;
  o.series_saddr = 0;
  o.hankel_saddr = 0;
  o.ft.saddr = 0;
  o.pad_ = 0;
  return o;
}

// sum_k (w^2/4)^k / (k! (nu+1)_k): the ascending series without its prefactor.
// Device: four terms per iteration from ONE previous term, t_{k+i-1} = t_{k-1} q^i R_{k,i} with q^2, q^3, q^4 formed once —
// the dependent chain is one complex product per four terms instead of four, the scalings are folded into the FMAs of
// the sum (33 instructions per four terms; the term-by-term loop with its reciprocals issued 46 per two). The
// convergence test runs once per group (three extra terms at most).
HH_HD cplx bessel_series_sum(const BesselOrder &o, cplx w) {
  const cplx q = 0.25 * (w * w);
  cplx term = mk(1.0), sum = mk(1.0);
#ifdef __CUDA_ARCH__
  const cplx q2 = q * q, q3 = q2 * q, q4 = q2 * q2;
  unsigned addr = o.series_saddr;
#pragma unroll 1
  for (int g = 0; g < kSeriesMaxTerms / 4; ++g, addr += 32) {
    double r1, r2, r3, r4;
    lds_f64x2(addr, r1, r2);
    lds_f64x2(addr + 16, r3, r4);
    const cplx u1 = term * q, u2 = term * q2, u3 = term * q3, u4 = term * q4;
    term = r4 * u4;
    double sr = fma(r1, u1.re, sum.re), si = fma(r1, u1.im, sum.im);
    double tr = fma(r2, u2.re, term.re), ti = fma(r2, u2.im, term.im);
    sr = fma(r3, u3.re, sr);
    si = fma(r3, u3.im, si);
    sum = cplx{sr + tr, si + ti};
    // |term| <= 1e-17 |sum| in the 1-norm (within sqrt 2 of the 2-norm test)
    if (fabs(term.re) + fabs(term.im) < 1e-17 * (fabs(sum.re) + fabs(sum.im))) break;
  }
#else
  const double nu = o.nu;
  for (int k = 1; k + 1 < kSeriesMaxTerms; k += 2) {
    term = (term * q) * (1.0 / ((double)k * (nu + (double)k)));
    sum = sum + term;
    term = (term * q) * (1.0 / ((double)(k + 1) * (nu + (double)(k + 1))));
    sum = sum + term;
    if (fabs(term.re) + fabs(term.im) < 1e-17 * (fabs(sum.re) + fabs(sum.im))) break;
  }
#endif
  return sum;
}

// Hankel sums s1 = sum (-1)^k a_k / w^k, s2 = sum a_k / w^k, a_k = prod (4 nu^2 - (2j-1)^2) / (8 j), stopped at the
// smallest term. Device: four terms per iteration from one previous term (1/w^2, 1/w^3, 1/w^4 formed once; the table holds the
// running products of the b_k inside each group of four), scalings folded into the FMAs of the two sums: ~41 instructions
// per four terms (the term-by-term loop issued 45 per term). A group whose last term is not smaller than the previous
// group's (the expansion is about to diverge: never for |w| >= r_asym) is left to the term-by-term tail.
HH_HD void bessel_hankel_sums(const BesselOrder &o, cplx w, cplx &s1, cplx &s2) {
  const cplx iw = crecip(w);
  const double mu4 = 4.0 * o.nu * o.nu;
  cplx t = mk(1.0);
  s1 = mk(1.0);
  s2 = mk(1.0);
  double last = 1.0;  // |t|^2 of the previous term
  int k = 1;
#ifdef __CUDA_ARCH__
  const cplx iw2 = iw * iw, iw3 = iw2 * iw, iw4 = iw2 * iw2;
  unsigned addr = o.hankel_saddr;
#pragma unroll 1
  for (; k + 3 < kHankelMaxTerms; k += 4, addr += 32) {
    double b1, b2, b3, b4;
    lds_f64x2(addr, b1, b2);
    lds_f64x2(addr + 16, b3, b4);
    const cplx u1 = t * iw, u2 = t * iw2, u3 = t * iw3, u4 = t * iw4;
    const cplx t4 = b4 * u4;  // a_{k+3} / w^{k+3}
    const double m4 = cabs2(t4);
    if (m4 > last) break;
    last = m4;
    t = t4;
    const cplx ev = cplx{fma(b2, u2.re, t4.re), fma(b2, u2.im, t4.im)};          // the two even-numbered terms (k odd first)
    const cplx od = cplx{fma(b3, u3.re, b1 * u1.re), fma(b3, u3.im, b1 * u1.im)};  // the two odd-numbered terms
    s2 = cplx{s2.re + (ev.re + od.re), s2.im + (ev.im + od.im)};
    s1 = cplx{s1.re + (ev.re - od.re), s1.im + (ev.im - od.im)};
    if (m4 < 1e-34) return;
  }
#endif
  for (; k < kHankelMaxTerms; ++k) {
    const double odd = (double)(2 * k - 1);
    t = (t * iw) * ((mu4 - odd * odd) * rcp_fast(8.0 * (double)k));  // a_k / w^k
    const double m = cabs2(t);
    if (m > last) break;  // the expansion has started to diverge
    last = m;
    s1 = (k & 1) ? s1 - t : s1 + t;
    s2 = s2 + t;
    if (m < 1e-34) break;
  }
}

HH_HD BesselEF besseli_series_ef(const BesselOrder &o, cplx w, double log_aw, double arg_w) {
  const cplx sum = bessel_series_sum(o, w);
  // (w/2)^nu / Gamma(nu+1) * sum
  return BesselEF{cplx{o.nu * (log_aw - 0.6931471805599453) - o.lgam_nu1, o.nu * arg_w}, sum};
}

HH_HD BesselEF besseli_asymptotic_ef(const BesselOrder &o, cplx w, double log_aw, double arg_w) {
  cplx s1, s2;
  bessel_hankel_sums(o, w, s1, s2);
  // I = e^w / sqrt(2 pi w) [ s1 + e^{-2w +- i pi (nu + 1/2)} s2 ],  + for Im w >= 0
  const double sgn = w.im >= 0.0 ? 1.0 : -1.0;
  const cplx e2 = cexp_(o.ft, cplx{-2.0 * w.re, -2.0 * w.im + sgn * kBesselPi * (o.nu + 0.5)});
  return BesselEF{cplx{w.re - 0.5 * (1.8378770664093453 + log_aw), w.im - 0.5 * arg_w}, s1 + e2 * s2};  // log(2 pi)
}

// ---- ascending series: |w| - Re w <= 5, |w| <~ 25 ---------------------------------------------------------------
HH_HD cplx log_besseli_series(const BesselOrder &o, cplx w) {
  const cplx sum = bessel_series_sum(o, w);
  // (w/2)^nu / Gamma(nu+1) * sum
  return o.nu * clog_(o.ft, 0.5 * w) - o.lgam_nu1 + clog_(o.ft, sum);
}

// ---- Hankel expansion: |w| large, Re w >= 0 -----------------------------------------------------------------
HH_HD cplx log_besseli_asymptotic(const BesselOrder &o, cplx w) {
  cplx s1, s2;
  bessel_hankel_sums(o, w, s1, s2);
  // I = e^w / sqrt(2 pi w) [ s1 + e^{-2w +- i pi (nu + 1/2)} s2 ],  + for Im w >= 0
  const double sgn = w.im >= 0.0 ? 1.0 : -1.0;
  const cplx e2 = cexp_(o.ft, cplx{-2.0 * w.re, -2.0 * w.im + sgn * kBesselPi * (o.nu + 0.5)});
  return w - 0.5 * clog_(o.ft, (2.0 * kBesselPi) * w) + clog_(o.ft, s1 + e2 * s2);
}

// ---- Debye's uniform expansion for large orders (DLMF 10.41.3, 10.41.7-9): Re w > 0 ---------------------------------
//   I_nu(nu z) ~ e^{nu eta} / sqrt(2 pi nu s) sum_k u_k(t) / nu^k,  s = sqrt(1 + z^2), t = 1/s, eta = s + log(z / (1 + s))
// u_k(t) = t^k P_k(t^2), P_k of degree k (generated by the recurrence 10.41.9 in rational arithmetic,
// tools/gen_debye.py). Used where kDebyeMinOrder <= nu and r_near <= |w| < r_asym, inside the sector bessel_debye_sector: there
// |z| >= r_near / nu keeps |t| small and nine terms are exact to rounding (measured against 30-digit values: 5e-16
// relative for nu from 12.7 to 1000; six terms already are) and the recessive exponential is below e^{-160}.
constexpr int kDebyeTerms = 8;
HH_HD bool bessel_debye_sector(double nu, double aw, cplx w) {  // Re w >= 0
  if (aw - w.re <= 5.0) return true;                            // the strip the series uses below r_near
  if (nu >= 100.0) return fabs(w.im) <= 3.6 * w.re;             // |arg w| <= 1.3
  return nu >= 30.0 && fabs(w.im) <= 0.7 * w.re;                // |arg w| <= 0.61
}
#define HH_DEBYE_COEFFS                                                                                                   \
  {1.0,                                                                                                                   \
   0.125, -0.20833333333333334,                                                                                           \
   0.0703125, -0.4010416666666667, 0.3342013888888889,                                                                    \
   0.0732421875, -0.8912109375, 1.8464626736111112, -1.0258125964506173,                                                  \
   0.112152099609375, -2.3640869140625, 8.78912353515625, -11.207002616222994, 4.669584423426247,                         \
   0.22710800170898438, -7.368794359479632, 42.53499874538846, -91.81824154324002, 84.63621767460073, -28.212072558200244, \
   0.5725014209747314, -26.491430486951554, 218.1905117442116, -699.5796273761325, 1059.9904525279999, -765.2524681411817, \
   212.57013003921713,                                                                                                    \
   1.7277275025844574, -108.09091978839466, 1200.9029132163525, -5305.646978613403, 11655.393336864534, -13586.550006434138, \
   8061.722181737309, -1919.457662318407,                                                                                 \
   6.074042001273483, -493.915304773088, 7109.514302489364, -41192.65496889755, 122200.46498301746, -203400.17728041555,  \
   192547.00123253153, -96980.59838863752, 20204.29133096615}
#ifdef __CUDACC__
__device__ __constant__ double kDebyeDev[(kDebyeTerms + 1) * (kDebyeTerms + 2) / 2] = HH_DEBYE_COEFFS;
#endif
HH_HD_OUTLINE cplx log_besseli_debye(const BesselOrder &o, cplx w) {
#ifdef __CUDA_ARCH__
  const double *c = kDebyeDev;  // rolled loops over a constant-memory table: cold code, kept small (instruction cache)
#else
  static const double c[(kDebyeTerms + 1) * (kDebyeTerms + 2) / 2] = HH_DEBYE_COEFFS;
#endif
  const double inu = 1.0 / o.nu;
  const cplx z = inu * w;
  const cplx s = csqrt_(1.0 + z * z);  // principal branch: Re s > 0 for Re z > 0
  const cplx t = crecip(s), t2 = t * t, x = inu * t;
  cplx acc = mk(0.0);
#pragma unroll 1
  for (int k = kDebyeTerms; k >= 0; --k) {
    const int base = k * (k + 1) / 2;
    cplx pk = mk(c[base + k]);
#pragma unroll 1
    for (int j = k - 1; j >= 0; --j) pk = pk * t2 + c[base + j];
    acc = acc * x + pk;
  }
  const cplx eta = s + clog_(o.ft, z / (1.0 + s));
  // -1/2 log(2 pi nu s)
  return o.nu * eta - 0.5 * (clog_(o.ft, s) + (1.8378770664093453 + log(o.nu))) + clog_(o.ft, acc);
}

// ---- continued fractions + Wronskian: 2 <= |w|, Re w >= 0, order xnu >= 0 ------------------------------------------
// Returns log I_xnu(w); if dlog is non-null also I'_xnu / I_xnu.
HH_HD_OUTLINE cplx log_besseli_cf(double xnu, cplx x, cplx *ratio_deriv) {
  const double EPS = 1e-16, FPMIN = 1e-200;
  const int nl = (int)(xnu + 0.5);
  const double xmu = xnu - (double)nl, xmu2 = xmu * xmu;
  const cplx xi = 1.0 / x, xi2 = 2.0 * xi;
  // CF1: h = I'_xnu / I_xnu
  cplx h = xnu * xi;
  if (fabs(h.re) + fabs(h.im) < FPMIN) h = mk(FPMIN);
  cplx b = xnu * xi2, d = mk(0.0), c = h;
  const double ax = cabs(x);
  const int maxit = 400 + 2 * (ax < 1e5 ? (int)ax : (ax == ax ? 100000 : 0));  // NaN in, NaN out, without the long loop
#pragma unroll 1
  for (int i = 0; i < maxit; ++i) {
    b = b + xi2;
    d = 1.0 / (b + d);
    c = b + 1.0 / c;
    const cplx del = c * d;
    h = del * h;
    if (fabs(del.re - 1.0) + fabs(del.im) < EPS) break;
  }
  // downward recurrence to order xmu
  cplx ril = mk(1.0), ripl = h;
  double lscale = 0.0;  // log of the factors taken out of (ril, ripl): I_mu / I_xnu = ril e^{lscale}, which leaves the
  cplx fact = xnu * xi; // binary64 range for large orders (e^{nu^2 / 2|x|})
#pragma unroll 1
  for (int l = nl; l >= 1; --l) {
    const cplx ritemp = fact * ril + ripl;
    fact = fact - xi;
    ripl = fact * ritemp + ril;
    ril = ritemp;
    const double m = cabs2(ril);
    if (m > 1e200) {
      const double sc = 1e-100;
      ril = sc * ril; ripl = sc * ripl;
      lscale += 230.25850929940458;  // 100 log 10
    }
  }
  const cplx f = ripl / ril;
  // CF2 (Steed): K_mu and K_mu+1, scaled by e^{x}
  cplx bb = 2.0 * (1.0 + x);
  cplx dd = 1.0 / bb, hh2 = dd, delh = dd;
  cplx q1 = mk(0.0), q2 = mk(1.0);
  const double a1 = 0.25 - xmu2;
  cplx q = mk(a1), cc = mk(a1);
  double a = -a1;
  cplx s = 1.0 + q * delh;
#pragma unroll 1
  for (int i = 2; i < 2000; ++i) {
    a -= 2.0 * (double)(i - 1);
    cc = (-a / (double)i) * cc;
    const cplx qnew = (q1 - bb * q2) / a;
    q1 = q2;
    q2 = qnew;
    q = q + cc * qnew;
    bb = bb + 2.0;
    dd = 1.0 / (bb + a * dd);
    delh = (bb * dd - 1.0) * delh;
    hh2 = hh2 + delh;
    const cplx dels = q * delh;
    s = s + dels;
    if (cabs2(dels) < EPS * EPS * cabs2(s)) break;
  }
  hh2 = a1 * hh2;
  const cplx rkmu = csqrt_((0.5 * kBesselPi) * xi) / s;          // K_mu e^{x}
  const cplx rk1 = rkmu * (xmu + x + 0.5 - hh2) * xi;             // K_mu+1 e^{x}
  const cplx rkmup = xmu * xi * rkmu - rk1;                        // K'_mu e^{x}
  const cplx rimu_scaled = xi / (f * rkmu - rkmup);                // I_mu e^{-x}
  if (ratio_deriv) *ratio_deriv = h;
  // I_xnu = I_mu / (ril e^{lscale})
  return x + clog_(FastRef{0}, rimu_scaled) - clog_(FastRef{0}, ril) - lscale;
}

// log I_nu(z), any complex z != 0. The imaginary part is a valid argument of I_nu(z) (branch unspecified).
HH_HD_OUTLINE cplx log_besseli(const BesselOrder &o, cplx z) {
  const double nu = o.nu;
  cplx w = z;
  double rot = 0.0;  // I_nu(z) = e^{i rot} I_nu(w)
  if (z.re < 0.0) {
    w = -z;
    rot = (z.im >= 0.0 ? 1.0 : -1.0) * kBesselPi * nu;
  }
  if (fabs(w.re) < 1e-140 && fabs(w.im) < 1e-140) {
    // |w|^2 would underflow. I_nu(w) = (w/2)^nu / Gamma(nu+1) (1 + O(|w|^2)): the first term, with the logarithm taken of
    // the rescaled argument (exact power of two)
    const cplx l = clog_(o.ft, cplx{w.re * 0x1p+600, w.im * 0x1p+600});  // log 0 = -inf: I_nu(0) = 0 or Inf, as it should
    return cplx{nu * (l.re - 601.0 * 0.6931471805599453) - o.lgam_nu1, nu * l.im + rot};
  }
  const double aw = cabs(w);
  if (aw != aw) return cplx{aw, aw};  // NaN in, NaN out (no branch below would terminate early)
  cplx r;
  if (aw <= 5.0 || (aw < o.r_near && aw - w.re <= 5.0)) {
    r = log_besseli_series(o, w);
  } else if (aw >= o.r_asym) {
    r = log_besseli_asymptotic(o, w);
  } else if (nu >= kDebyeMinOrder && aw >= o.r_near && bessel_debye_sector(nu, aw, w)) {
    r = log_besseli_debye(o, w);
  } else if (nu >= 0.0) {
    r = log_besseli_cf(nu, w, nullptr);
  } else {
    // I_nu = I'_(nu+1) + ((nu+1)/w) I_(nu+1)
    cplx dl;
    const cplx l1 = log_besseli_cf(nu + 1.0, w, &dl);
    r = l1 + clog_(o.ft, dl + (nu + 1.0) / w);
  }
  r.im += rot;
  return r;
}

// I_nu(z) = exp(E) F for z != 0 with log|z| and arg z supplied by the caller (arg z in (-pi, pi]).
HH_HD BesselEF besseli_ef(const BesselOrder &o, cplx z, double log_az, double arg_z) {
  const double nu = o.nu;
  cplx w = z;
  double rot = 0.0, arg_w = arg_z;
  if (z.re < 0.0) {
    w = -z;
    const double sg = z.im >= 0.0 ? 1.0 : -1.0;
    rot = sg * kBesselPi * nu;
    arg_w = arg_z - sg * kBesselPi;  // arg(-z) in (-pi/2, pi/2)
  }
  // region tests on |w|^2 (Re w >= 0 here: |w| - Re w <= 5  <=>  |w|^2 <= (5 + Re w)^2): no square root
  const double aw2 = cabs2(w), rn2 = o.r_near2, edge = 5.0 + w.re;
  BesselEF r;
  if (aw2 <= 25.0 || (aw2 < rn2 && aw2 <= edge * edge)) {
    r = besseli_series_ef(o, w, log_az, arg_w);
  } else if (aw2 >= rn2 && (o.nu < kDebyeMinOrder || aw2 >= o.r_asym2)) {
    // r_near == r_asym below kDebyeMinOrder: the order is the same for every lane, and the second load and test are skipped
    r = besseli_asymptotic_ef(o, w, log_az, arg_w);
  } else {
    r.E = log_besseli(o, w);  // Debye (large orders) or continued fractions (rare: strongly rotated arguments)
    r.F = mk(1.0);
  }
  r.E.im += rot;
  return r;
}

// ---- Broadie-Kaya: characteristic function of int_0^tau V ds given (V0, VT) ------------------------------------------
// HestonCFIterator + evaluate_chf, src/distributions/heston.jl:150-212.
struct BkParams {
  double kappa, xi2, tau;  // xi2 = sigma^2 (vol of vol squared)
  double zeta_k, eta_k;    // :167-168
  double wk;               // z_kappa = sqrt(V0 VT) wk  (:169)
  double h_fd, cf_tol, atol;
  int n_std, max_terms;
  int widen_fd, pad_;      // re-read a noise-dominated variance at a wider step (hh_bk.cu: bk_variance_from_modulus)
  double noise_scale;      // 4 eps / h_fd^2 / 0.02
  BesselOrder ord;         // nu = d/2 - 1  (:164-165)
  // transition constants (sample_V_T :128-131, sample_log_S_T :285-297)
  double dof, c_scale, lam_scale;  // d, c, lambda = lam_scale * V
  double r_tau, kappa_theta_tau, rho_over_xi, one_m_rho2, rho, theta;
};

struct BkCf {       // per (V0, VT) pair: HestonCFIterator
  double sv;        // sqrt(V0 VT)
  double vsum_s;    // (V0 + VT) / sigma^2
  double sv4_xi2;   // 4 sqrt(V0 VT) / sigma^2
  cplx logIk;       // log I_nu(z_kappa)
};

// log I_nu(x) for real x > 0 (z_kappa, heston.jl:169-173): the same two expansions in real arithmetic — a quarter of the
// floating-point work of the complex routine, once per transition. The e^{-2x} part of the Hankel form is below
// 2e-18 relative from r_asym on and is dropped. Orders in (-1, 0) and tiny or non-finite x take the general routine.
HH_HD double log_besseli_real(const BesselOrder &o, double x) {
  if (!(x > 1e-300) || !(x < 1e300)) return log_besseli(o, mk(x)).re;
  if (x >= o.r_near && x < o.r_asym) return log_besseli(o, mk(x)).re;  // large orders: Debye
  if (x < o.r_near) {
    const double q = 0.25 * x * x;
    double term = 1.0, sum = 1.0;
#ifdef __CUDA_ARCH__
    const double q2 = q * q, q3 = q2 * q, q4 = q2 * q2;
    unsigned addr = o.series_saddr;
#pragma unroll 1
    for (int g = 0; g < kSeriesMaxTerms / 4; ++g, addr += 32) {
      double r1, r2, r3, r4;
      lds_f64x2(addr, r1, r2);
      lds_f64x2(addr + 16, r3, r4);
      const double t4 = (term * q4) * r4;
      sum += fma(term * q, r1, (term * q3) * r3) + fma(term * q2, r2, t4);
      term = t4;
      if (term < 1e-17 * sum) break;
    }
#else
    for (int k = 1; k < kSeriesMaxTerms; ++k) {
      term *= q / ((double)k * (o.nu + (double)k));
      sum += term;
      if (term < 1e-17 * sum) break;
    }
#endif
    return o.nu * (flog(o.ft, x) - 0.6931471805599453) - o.lgam_nu1 + flog(o.ft, sum);
  }
  const double ix = rcp_fast(x), mu4 = 4.0 * o.nu * o.nu;
  double t = 1.0, s1 = 1.0, last = 1.0;
  for (int k = 1; k < kHankelMaxTerms; ++k) {
    const double odd = (double)(2 * k - 1);
    t *= -ix * ((mu4 - odd * odd) * rcp_fast(8.0 * (double)k));  // (-1)^k a_k / x^k
    const double m = fabs(t);
    if (m > last) break;
    last = m;
    s1 += t;
    if (m < 1e-17) break;
  }
  return x - 0.5 * (1.8378770664093453 + flog(o.ft, x)) + flog(o.ft, s1);  // log(2 pi)
}

HH_HD BkCf bk_cf_init(const BkParams &p, double V0, double VT) {
  BkCf it;
  // A variance that has underflowed (degrees of freedom << 1: the chi-square draw is a power 1/d of a uniform) is held at
  // 1e-100: V0 VT, |z_gamma|^2 and their logarithms stay normal numbers. The reference has sqrt(0) = 0 there, then
  // besseli(nu < 0, 0) = Inf and Inf - Inf.
  V0 = fmax(V0, 1e-100);
  VT = fmax(VT, 1e-100);
  it.sv = sqrt_fast(V0 * VT);
  it.vsum_s = (V0 + VT) / p.xi2;
  it.sv4_xi2 = it.sv * 4.0 / p.xi2;
  it.logIk = mk(p.ord.nu >= 0.0 ? log_besseli_real(p.ord, it.sv * p.wk) : log_besseli(p.ord, mk(it.sv * p.wk)).re);
  return it;
}

// Phi(a) with the unwrapped angle of z_gamma carried in theta_prev (NaN = first evaluation), heston.jl:184-212.
HH_HD_OUTLINE cplx bk_chf(const BkParams &p, const BkCf &it, double a, double &theta_prev) {
  const cplx g = csqrt_pos_(cplx{p.kappa * p.kappa, -2.0 * p.xi2 * a});    // gamma            :190  (Re gamma^2 = kappa^2 > 0)
  const FastRef ft = p.ord.ft;
  const cplx egh = cexp_(ft, (-0.5 * p.tau) * g);  // e^{-g tau / 2}
  const cplx eg = egh * egh;
  const cplx omeg = 1.0 - eg;
  // one complex reciprocal serves the three quotients of :191-193 (1 / zeta_g = g / (1 - e^{-g tau}))
  const cplx g_io = g * crecip(omeg);
  const cplx eta_g = g_io * (1.0 + eg);                                       // :192
  const cplx zg = it.sv4_xi2 * (g_io * egh);                                  // nu_gamma         :193
  const double th = fatan2(ft, zg.im, zg.re);                                 // :198
  double thu = th;
  if (!(theta_prev != theta_prev)) {                                          // :199-205
    double dlt = th - theta_prev;
    dlt -= 2.0 * kBesselPi * nearbyint(dlt * (0.5 / kBesselPi));
    thu = theta_prev + dlt;
  }
  theta_prev = thu;
  // I_nu(z_gamma) = exp(E) F with log|z_gamma| and the angle already in hand                       :206-207
  BesselEF ig = besseli_ef(p.ord, zg, 0.5 * flog(ft, cabs2(zg)), th);
  ig.E.im += p.ord.nu * (thu - th);
  // phi = exp(-(g - k) tau / 2) (zeta_k / zeta_g) exp((V0+VT)/s^2 (eta_k - eta_g)) exp(logIg - logIk)   :195-211
  const cplx ex = (-0.5 * p.tau) * (g - p.kappa) + it.vsum_s * (p.eta_k - eta_g) + (ig.E - it.logIk);
  return ((p.zeta_k * g_io) * ig.F) * cexp_(ft, ex);  // zeta_k / zeta_g (:191)
}

}  // namespace hh
