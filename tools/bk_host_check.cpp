// Host build of hh_bessel.cuh for validating the Bessel / characteristic-function arithmetic on a CPU-only box
// (development tool, not product and not oracle): g++ -O2 -shared -fPIC -o tools/_build/libbk_host.so tools/bk_host_check.cpp
#include "../hedgehog.jl_b200/csrc/hh_bessel.cuh"
using namespace hh;
extern "C" {
void bkh_log_besseli(double nu, const double *zr, const double *zi, int n, double *ore, double *oim) {
  BesselOrder o = make_bessel_order(nu);
  for (int i = 0; i < n; ++i) {
    cplx r = log_besseli(o, cplx{zr[i], zi[i]});
    ore[i] = r.re;
    oim[i] = r.im;
  }
}
}

static BkParams make_params(double kappa, double theta, double sigma, double tau) {
  BkParams p{};
  p.kappa = kappa; p.xi2 = sigma * sigma; p.tau = tau;
  const double E = -expm1(-kappa * tau);
  p.zeta_k = E / kappa;
  p.eta_k = kappa * (1 + exp(-kappa * tau)) / E;
  p.wk = 4 * kappa * exp(-0.5 * kappa * tau) / p.xi2 / E;
  p.ord = make_bessel_order(0.5 * (4 * kappa * theta / p.xi2) - 1);
  return p;
}
extern "C" {
// phi(a_j), j = 0..na-1, angle unwrapped along j (theta_prev starts NaN)
void bkh_chf(double kappa, double theta, double sigma, double tau, double V0, double VT, const double *a, int na,
             double *ore, double *oim) {
  BkParams p = make_params(kappa, theta, sigma, tau);
  BkCf it = bk_cf_init(p, V0, VT);
  double th = NAN;
  for (int j = 0; j < na; ++j) {
    cplx r = bk_chf(p, it, a[j], th);
    ore[j] = r.re;
    oim[j] = r.im;
  }
}
}
