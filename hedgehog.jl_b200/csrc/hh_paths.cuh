// hh_paths.cuh — the SDE steppers, written once over a number type T.
//
// T = double          : plain pricing (the headline Heston Euler-Maruyama kernel)
// T = hh::Dual<P>     : forward-mode tangents, P directions at once — the in-kernel equivalent of the
//                       ForwardDiff.Dual the reference pushes through solve (src/greeks/greeks_problem.jl:249-262)
//
// Derivative rules that matter for parity with ForwardDiff's pathwise semantics:
//   max(x, 0) passes the tangent iff x > 0, otherwise the result is the constant 0;
//   sqrt at 0 therefore sees a zero tangent and stays finite (the reference would produce NaN there).
#pragma once
#include "hh_device.cuh"
#include "hh_fastnormal.cuh"

namespace hh {

template <int P>
struct Dual {
  double v;
  double d[P];
};

template <class T>
struct num_traits {
  static constexpr int ntan = 0;
};
template <int P>
struct num_traits<Dual<P>> {
  static constexpr int ntan = P;
};

// ---- double -----------------------------------------------------------------------------------------
__device__ __forceinline__ double value(double a) { return a; }
__device__ __forceinline__ double tangent(double, int) { return 0.0; }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double max0(double a) { return fmax(a, 0.0); }
__device__ __forceinline__ double sqrt0(double a) { return sqrt(fmax(a, 0.0)); }
__device__ __forceinline__ double exp_(double a) { return exp(a); }
__device__ __forceinline__ double expm1_fromexp(double a) { return exp(a) - 1.0; }

// ---- Dual<P> ----------------------------------------------------------------------------------------
#define HH_DUAL_LOOP _Pragma("unroll") for (int i = 0; i < P; ++i)
template <int P>
__device__ __forceinline__ double value(const Dual<P> &a) { return a.v; }
template <int P>
__device__ __forceinline__ double tangent(const Dual<P> &a, int i) { return a.d[i]; }

template <int P>
__device__ __forceinline__ Dual<P> operator+(const Dual<P> &a, const Dual<P> &b) {
  Dual<P> r;
  r.v = a.v + b.v;
  HH_DUAL_LOOP r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> operator-(const Dual<P> &a, const Dual<P> &b) {
  Dual<P> r;
  r.v = a.v - b.v;
  HH_DUAL_LOOP r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> operator-(const Dual<P> &a) {
  Dual<P> r;
  r.v = -a.v;
  HH_DUAL_LOOP r.d[i] = -a.d[i];
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> operator-(const Dual<P> &a, double b) {
  Dual<P> r = a;
  r.v = a.v - b;
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> operator*(const Dual<P> &a, const Dual<P> &b) {
  Dual<P> r;
  r.v = a.v * b.v;
  HH_DUAL_LOOP r.d[i] = fma(a.v, b.d[i], a.d[i] * b.v);
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> operator*(const Dual<P> &a, double b) {
  Dual<P> r;
  r.v = a.v * b;
  HH_DUAL_LOOP r.d[i] = a.d[i] * b;
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> operator*(double b, const Dual<P> &a) { return a * b; }

// a*b + c
template <int P>
__device__ __forceinline__ Dual<P> fma_(const Dual<P> &a, const Dual<P> &b, const Dual<P> &c) {
  Dual<P> r;
  r.v = fma(a.v, b.v, c.v);
  HH_DUAL_LOOP r.d[i] = fma(a.v, b.d[i], fma(a.d[i], b.v, c.d[i]));
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> fma_(double a, const Dual<P> &b, const Dual<P> &c) {
  Dual<P> r;
  r.v = fma(a, b.v, c.v);
  HH_DUAL_LOOP r.d[i] = fma(a, b.d[i], c.d[i]);
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> fma_(const Dual<P> &a, double b, const Dual<P> &c) { return fma_(b, a, c); }

template <int P>
__device__ __forceinline__ Dual<P> max0(const Dual<P> &a) {
  Dual<P> r;
  const bool pos = a.v > 0.0;
  r.v = pos ? a.v : 0.0;
  HH_DUAL_LOOP r.d[i] = pos ? a.d[i] : 0.0;
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> sqrt0(const Dual<P> &a) {
  Dual<P> r;
  const bool pos = a.v > 0.0;
  const double s = sqrt(pos ? a.v : 0.0);
  const double h = pos ? 0.5 / s : 0.0;
  r.v = s;
  HH_DUAL_LOOP r.d[i] = a.d[i] * h;
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> exp_(const Dual<P> &a) {
  Dual<P> r;
  r.v = exp(a.v);
  HH_DUAL_LOOP r.d[i] = r.v * a.d[i];
  return r;
}
template <int P>
__device__ __forceinline__ Dual<P> expm1_fromexp(const Dual<P> &a) {
  Dual<P> r;
  const double e = exp(a.v);
  r.v = e - 1.0;
  HH_DUAL_LOOP r.d[i] = e * a.d[i];
  return r;
}
#undef HH_DUAL_LOOP

// ---- model parameters in number type T ---------------------------------------------------------------
template <class T>
struct PathParams {
  // common
  double dt, sqdt;
  // GBM
  T x0, S0, dt_drift, sigma, sig_sqdt, mu, sd;
  // Heston (r doubles as the log-drift rate)
  T v0, r, kappa, theta, xi, a11, a12, a21, a22;
};

// ---- steppers ----------------------------------------------------------------------------------------

// LogHestonProblem drift/diffusion (reference src/distributions/heston.jl:8-16) advanced by
// Euler-Maruyama with the split step of StochasticDiffEq's EM() [upstream]: K = u + dt f(u), u' = K + g(K) dW.
template <class T>
__device__ __forceinline__ void heston_em_step(const PathParams<T> &p, bool split, T &x, T &v, const T &dW1,
                                               const T &dW2) {
  const T vplus = max0(v);
  const T K1 = fma_(p.dt, fma_(-0.5, vplus, p.r), x);
  const T K2 = fma_(p.dt, p.kappa * (p.theta - vplus), v);
  const T s = sqrt0(split ? K2 : v);
  x = fma_(s, dW1, K1);
  v = fma_(p.xi * s, dW2, K2);
}

// The same step for the native-RNG pricing kernel: constants folded on the host, xi folded into dW2, and the
// branch-free square root. sqrt(max(., 1e-300)) replaces sqrt(max(., 0)): the difference (1e-150) is far below
// one ulp of the state.
struct HestonFolded {
  double rdt, neg_half_dt, neg_kdt, ktdt, b21, b22;
};
__device__ __forceinline__ void heston_em_step_fast(const HestonFolded &f, bool split, double &x, double &v, double dW1,
                                                    double xi_dW2) {
  const double vplus = max0_bits(v);
  const double K1 = fma(f.neg_half_dt, vplus, x + f.rdt);
  const double K2 = fma(f.neg_kdt, vplus, v + f.ktdt);
  const double s = fast_sqrt_pos(max_tiny_bits(split ? K2 : v));
  x = fma(s, dW1, K1);
  v = fma(s, xi_dW2, K2);
}

// The step with the Box-Muller radius folded into the diffusion's square root: dW1 = rad c1, xi dW2 = rad c2 with
// rad = sqrt(R2), so s dW = sqrt(max(K2, 0) R2) c — one square root per step instead of two.
__device__ __forceinline__ void heston_em_step_folded(const HestonFolded &f, bool split, double &x, double &v, double R2,
                                                      double c1, double c2) {
  const double vplus = max0_bits(v);
  const double K1 = fma(f.neg_half_dt, vplus, x + f.rdt);
  const double K2 = fma(f.neg_kdt, vplus, v + f.ktdt);
  const double sr = fast_sqrt_pos(max_tiny_bits((split ? max0_bits(K2) : vplus) * R2));
  x = fma(sr, c1, K1);
  v = fma(sr, c2, K2);
}

// LogGBMProblem under EM (heston.jl:33-52): x' = (x + dt (r - sigma^2/2)) + sigma dW
template <class T>
__device__ __forceinline__ void gbm_em_step(const PathParams<T> &p, T &x, double dW) {
  x = fma_(p.sigma, dW, x + p.dt_drift);
}

// GeometricBrownianMotionProcess increment [upstream]: S += S (exp((r - s^2/2) dt + s sqrt(dt) Z) - 1)
template <class T>
__device__ __forceinline__ void gbm_exact_step(const PathParams<T> &p, T &S, double z, double sign) {
  const T y = fma_(p.sig_sqdt, sign * z, p.dt_drift);
  S = fma_(S, expm1_fromexp(y), S);
}

}  // namespace hh
