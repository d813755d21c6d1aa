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
HH_HD_OUTLINE double carg(cplx a) { return atan2(a.im, a.re); }
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
HH_HD double sqrt_fast(double x) {
#ifdef __CUDA_ARCH__
  const int hi = __double2hiint(x);
  if (hi >= 0x00200000 && hi < 0x7fd00000) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double h0 = __hiloint2double(__double2hiint(y0) - 0x00100000, 0);  // y0 / 2 (the seed's low word is zero)
    const double s0 = x * y0;
    const double e = fma(-s0, h0, 0.5);
    const double pp = fma(e * 1.5, e, e);
    return fma(s0, pp, s0);
  }
#endif
  return sqrt(x);
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
HH_HD_OUTLINE cplx cexp_(cplx a) {
  const double e = exp(a.re);
  double s, c;
  sincos(a.im, &s, &c);
  return cplx{e * c, e * s};
}
// log|a| = log(re^2 + im^2) / 2: no hypot, no sqrt (the arguments here are far from the over/underflow of the squares)
HH_HD cplx clog_(cplx a) { return cplx{0.5 * log(cabs2(a)), carg(a)}; }
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

constexpr double kBesselPi = 3.14159265358979323846;

// I_nu = exp(E) F: the series and the Hankel expansion produce an algebraic factor F (the sum) next to an exponent that
// only needs log|w| and arg w, which the caller often has already (bk_chf unwraps arg z_gamma anyway). Returning
// (E, F) instead of log I avoids a complex logarithm of F per evaluation, and the caller exponentiates once.
struct BesselEF {
  cplx E, F;
};

constexpr int kSeriesMaxTerms = 100, kHankelMaxTerms = 60;

// sum_k (w^2/4)^k / (k! (nu+1)_k): the ascending series without its prefactor. `rk` (nullable) tabulates
// 1 / (k (nu + k)); the convergence test runs on every second term (one extra term at most).
HH_HD cplx bessel_series_sum(double nu, cplx w, const double *rk) {
  const cplx q = 0.25 * (w * w);
  cplx term = mk(1.0), sum = mk(1.0);
#pragma unroll 1
  for (int k = 1; k + 1 < kSeriesMaxTerms; k += 2) {
    const double r1 = rk ? rk[k] : rcp_fast((double)k * (nu + (double)k));
    const double r2 = rk ? rk[k + 1] : rcp_fast((double)(k + 1) * (nu + (double)(k + 1)));
    term = (term * q) * r1;
    sum = sum + term;
    term = (term * q) * r2;
    sum = sum + term;
    // |term| <= 1e-17 |sum| in the 1-norm (within sqrt 2 of the 2-norm test; two instructions fewer per pair of terms)
    if (fabs(term.re) + fabs(term.im) < 1e-17 * (fabs(sum.re) + fabs(sum.im))) break;
  }
  return sum;
}

// Hankel sums s1 = sum (-1)^k a_k / w^k, s2 = sum a_k / w^k, a_k = prod (4 nu^2 - (2j-1)^2) / (8 j), stopped at the
// smallest term. `bk` (nullable) tabulates (4 nu^2 - (2k-1)^2) / (8 k).
HH_HD void bessel_hankel_sums(double nu, cplx w, const double *bk, cplx &s1, cplx &s2) {
  const double mu4 = 4.0 * nu * nu;
  const cplx iw = crecip(w);
  cplx t = mk(1.0);
  s1 = mk(1.0);
  s2 = mk(1.0);
  double last = 1.0;  // |t|^2 of the previous term
#pragma unroll 1
  for (int k = 1; k < kHankelMaxTerms; ++k) {
    const double odd = (double)(2 * k - 1);
    const double b = bk ? bk[k] : (mu4 - odd * odd) * rcp_fast(8.0 * (double)k);
    t = (t * iw) * b;  // a_k / w^k
    const double m = cabs2(t);
    if (m > last) break;  // the expansion has started to diverge
    last = m;
    s1 = (k & 1) ? s1 - t : s1 + t;
    s2 = s2 + t;
    if (m < 1e-34) break;
  }
}

HH_HD BesselEF besseli_series_ef(double nu, double lgam_nu1, cplx w, double log_aw, double arg_w, const double *rk) {
  const cplx sum = bessel_series_sum(nu, w, rk);
  // (w/2)^nu / Gamma(nu+1) * sum
  return BesselEF{cplx{nu * (log_aw - 0.6931471805599453) - lgam_nu1, nu * arg_w}, sum};
}

HH_HD BesselEF besseli_asymptotic_ef(double nu, cplx w, double log_aw, double arg_w, const double *bk) {
  cplx s1, s2;
  bessel_hankel_sums(nu, w, bk, s1, s2);
  // I = e^w / sqrt(2 pi w) [ s1 + e^{-2w +- i pi (nu + 1/2)} s2 ],  + for Im w >= 0
  const double sgn = w.im >= 0.0 ? 1.0 : -1.0;
  const cplx e2 = cexp_(cplx{-2.0 * w.re, -2.0 * w.im + sgn * kBesselPi * (nu + 0.5)});
  return BesselEF{cplx{w.re - 0.5 * (1.8378770664093453 + log_aw), w.im - 0.5 * arg_w}, s1 + e2 * s2};  // log(2 pi)
}

// ---- ascending series: |w| - Re w <= 5, |w| <~ 25 ---------------------------------------------------------------
HH_HD cplx log_besseli_series(double nu, double lgam_nu1, cplx w, const double *rk) {
  const cplx sum = bessel_series_sum(nu, w, rk);
  // (w/2)^nu / Gamma(nu+1) * sum
  return nu * clog_(0.5 * w) - lgam_nu1 + clog_(sum);
}

// ---- Hankel expansion: |w| large, Re w >= 0 -----------------------------------------------------------------
HH_HD cplx log_besseli_asymptotic(double nu, cplx w, const double *bk) {
  cplx s1, s2;
  bessel_hankel_sums(nu, w, bk, s1, s2);
  // I = e^w / sqrt(2 pi w) [ s1 + e^{-2w +- i pi (nu + 1/2)} s2 ],  + for Im w >= 0
  const double sgn = w.im >= 0.0 ? 1.0 : -1.0;
  const cplx e2 = cexp_(cplx{-2.0 * w.re, -2.0 * w.im + sgn * kBesselPi * (nu + 0.5)});
  return w - 0.5 * clog_((2.0 * kBesselPi) * w) + clog_(s1 + e2 * s2);
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
  const int maxit = 400 + 2 * (int)cabs(x);
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
  cplx ril = mk(1.0), ripl = h, ril1 = ril, rip1 = ripl;
  cplx fact = xnu * xi;
#pragma unroll 1
  for (int l = nl; l >= 1; --l) {
    const cplx ritemp = fact * ril + ripl;
    fact = fact - xi;
    ripl = fact * ritemp + ril;
    ril = ritemp;
    // rescale to stay in range (ratios are all that matter)
    const double m = cabs2(ril);
    if (m > 1e200) {
      const double sc = 1e-100;
      ril = sc * ril; ripl = sc * ripl; ril1 = sc * ril1; rip1 = sc * rip1;
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
  // I_xnu = I_mu * ril1 / ril
  return x + clog_(rimu_scaled) + clog_(ril1 / ril);
}

// Host-prepared constants of one order.
struct BesselOrder {
  double nu;        // order, > -1
  double lgam_nu1;  // lgamma(nu + 1)
  double r_asym;    // |w| from which the Hankel expansion is used
  // optional coefficient tables (the order is fixed per launch; the Broadie-Kaya kernel fills them in shared memory):
  // series_rk[k] = 1 / (k (nu + k)), hankel_bk[k] = (4 nu^2 - (2k - 1)^2) / (8 k). NULL: computed on the fly.
  const double *series_rk;
  const double *hankel_bk;
};
inline BesselOrder make_bessel_order(double nu) {
  BesselOrder o;
  o.nu = nu;
  o.lgam_nu1 = lgamma(nu + 1.0);
  o.r_asym = 20.0 + 0.5 * nu * nu;
  o.series_rk = nullptr;
  o.hankel_bk = nullptr;
  return o;
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
  const double aw = cabs(w);
  cplx r;
  if (aw <= 5.0 || (aw < o.r_asym && aw - w.re <= 5.0)) {
    r = log_besseli_series(nu, o.lgam_nu1, w, o.series_rk);
  } else if (aw >= o.r_asym) {
    r = log_besseli_asymptotic(nu, w, o.hankel_bk);
  } else if (nu >= 0.0) {
    r = log_besseli_cf(nu, w, nullptr);
  } else {
    // I_nu = I'_(nu+1) + ((nu+1)/w) I_(nu+1)
    cplx dl;
    const cplx l1 = log_besseli_cf(nu + 1.0, w, &dl);
    r = l1 + clog_(dl + (nu + 1.0) / w);
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
  const double aw2 = cabs2(w), ra2 = o.r_asym * o.r_asym, edge = 5.0 + w.re;
  BesselEF r;
  if (aw2 <= 25.0 || (aw2 < ra2 && aw2 <= edge * edge)) {
    r = besseli_series_ef(nu, o.lgam_nu1, w, log_az, arg_w, o.series_rk);
  } else if (aw2 >= ra2) {
    r = besseli_asymptotic_ef(nu, w, log_az, arg_w, o.hankel_bk);
  } else {
    r.E = log_besseli(o, w);  // continued fractions (rare: strongly rotated arguments of moderate size)
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

HH_HD BkCf bk_cf_init(const BkParams &p, double V0, double VT) {
  BkCf it;
  it.sv = sqrt(V0 * VT);
  it.vsum_s = (V0 + VT) / p.xi2;
  it.sv4_xi2 = it.sv * 4.0 / p.xi2;
  it.logIk = log_besseli(p.ord, mk(it.sv * p.wk));
  return it;
}

// Phi(a) with the unwrapped angle of z_gamma carried in theta_prev (NaN = first evaluation), heston.jl:184-212.
HH_HD_OUTLINE cplx bk_chf(const BkParams &p, const BkCf &it, double a, double &theta_prev) {
  const cplx g = csqrt_(cplx{p.kappa * p.kappa, -2.0 * p.xi2 * a});        // gamma            :190
  const cplx egh = cexp_((-0.5 * p.tau) * g);  // e^{-g tau / 2}
  const cplx eg = egh * egh;
  const cplx omeg = 1.0 - eg;
  // one complex reciprocal serves the three quotients of :191-193 (1 / zeta_g = g / (1 - e^{-g tau}))
  const cplx g_io = g * crecip(omeg);
  const cplx eta_g = g_io * (1.0 + eg);                                       // :192
  const cplx zg = it.sv4_xi2 * (g_io * egh);                                  // nu_gamma         :193
  const double th = carg(zg);                                                 // :198
  double thu = th;
  if (!(theta_prev != theta_prev)) {                                          // :199-205
    double dlt = th - theta_prev;
    dlt -= 2.0 * kBesselPi * nearbyint(dlt * (0.5 / kBesselPi));
    thu = theta_prev + dlt;
  }
  theta_prev = thu;
  // I_nu(z_gamma) = exp(E) F with log|z_gamma| and the angle already in hand                       :206-207
  BesselEF ig = besseli_ef(p.ord, zg, 0.5 * log(cabs2(zg)), th);
  ig.E.im += p.ord.nu * (thu - th);
  // phi = exp(-(g - k) tau / 2) (zeta_k / zeta_g) exp((V0+VT)/s^2 (eta_k - eta_g)) exp(logIg - logIk)   :195-211
  const cplx ex = (-0.5 * p.tau) * (g - p.kappa) + it.vsum_s * (p.eta_k - eta_g) + (ig.E - it.logIk);
  return ((p.zeta_k * g_io) * ig.F) * cexp_(ex);  // zeta_k / zeta_g (:191)
}

}  // namespace hh
