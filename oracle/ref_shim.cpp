// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Thin C shim over the REFERENCE's own plan-time maths, compiled by oracle/build.py
// straight from the sources where they lie:
//   /root/reference/src/common/{kernel,pswf,utils}.cpp  + include/finufft_common/*.h
// into oracle/_ref/libfinufft_ref_common.so (git-ignored, travels to the GPU box).
// Nothing from the reference is copied into this repo; this file only calls it.
// It pins oracle/finufft_oracle.cpp's restatement of: next235, gaussquad, PSWF0,
// theoretical_kernel_ns + clamp_kernel_ns, set_kernel_shape_given_ns, poly_fit<T>,
// lowest_sigma, nhg_type3, pswf_selfft_params.
//
// The rest of the reference CPU library (spread/interp/sort/execute) needs xsimd 14.3.0,
// POET v0.0.1 and FFTW 3.3.10 or ducc0_0_41_1, fetched by CPM at configure time
// (CMakeLists.txt:72-78) and absent from this image: unbuildable here.

#include <finufft_common/common.h>
#include <finufft_common/kernel.h>
#include <finufft_common/pswf.h>
#include <finufft_common/spread_opts.h>
#include <finufft_common/utils.h>

#include <cstdint>
#include <vector>

extern "C" {

int64_t ref_next235(int64_t n, int64_t fac) { return finufft::common::next235(n, fac); }

void ref_gaussquad(int n, double *x, double *w) { finufft::common::gaussquad(n, x, w); }

int ref_pswf(double c, int64_t n, const double *x, double *out) {
  try {
    finufft::common::PSWF0 psi(c);
    for (int64_t i = 0; i < n; ++i) out[i] = psi(x[i]);
  } catch (...) {
    return 27;
  }
  return 0;
}

// ns as the plan would choose it (theoretical + clamp) and beta for kerformula 8.
void ref_kernel_ns_beta(double tol, int dim, int type, double sigma, int is_float, int *ns,
                        double *beta) {
  finufft_spread_opts so{};
  so.upsampfac  = sigma;
  so.kerformula = 8;
  int nst       = finufft::kernel::theoretical_kernel_ns(tol, dim, type, so);
  *ns           = finufft::kernel::clamp_kernel_ns(nst, sigma, is_float ? 12 : 16, is_float != 0);
  so.nspread    = *ns;
  finufft::kernel::set_kernel_shape_given_ns(so, 0);
  *beta = so.beta;
}

// poly_fit<T> of panel `panel` of the width-ns PSWF kernel (what
// precompute_horner_coeffs feeds it, include/finufft/makeplan.hpp:236-248).
void ref_polyfit_pswf_f32(int ns, double beta, int panel, int n, float *out) {
  finufft_spread_opts so{};
  so.beta       = beta;
  so.kerformula = 8;
  so.nspread    = ns;
  auto ker      = finufft::kernel::kernel_definition_lambda(so);
  const float shift = float(2 * panel + 1 - ns);
  auto f            = [&](float x) -> float {
    const float z = (x + shift) / (float)ns;
    return (float)ker((double)z);
  };
  std::vector<float> c = finufft::kernel::poly_fit<float>(f, n);
  for (int i = 0; i < n; ++i) out[i] = c[i];
}
void ref_polyfit_pswf_f64(int ns, double beta, int panel, int n, double *out) {
  finufft_spread_opts so{};
  so.beta       = beta;
  so.kerformula = 8;
  so.nspread    = ns;
  auto ker      = finufft::kernel::kernel_definition_lambda(so);
  const double shift = double(2 * panel + 1 - ns);
  auto f             = [&](double x) -> double {
    const double z = (x + shift) / (double)ns;
    return (double)ker((double)z);
  };
  std::vector<double> c = finufft::kernel::poly_fit<double>(f, n);
  for (int i = 0; i < n; ++i) out[i] = c[i];
}

double ref_lowest_sigma(double tol, int dim, int ns, double eps_mach, double gridlen) {
  return finufft::common::lowest_sigma(tol, dim, ns, eps_mach, gridlen);
}

// the reference's own sigma search (src/common/kernel.cpp:203-257)
double ref_analytic_upsampfac(double tol, int dim, int type, int is_float, double maxN) {
  return finufft::common::analytic_upsampfac(tol, dim, type, is_float ? 1.1920929e-07 : 2.220446049250313e-16,
                                             is_float ? 12 : 16, is_float != 0, maxN);
}
int ref_upsampfac_feasible(double sigma, double tol, int dim, int type, int is_float, double maxN) {
  return finufft::common::upsampfac_feasible(sigma, tol, dim, type,
                                             is_float ? 1.1920929e-07 : 2.220446049250313e-16,
                                             is_float ? 12 : 16, is_float != 0, maxN)
             ? 1 : 0;
}

void ref_nhg_type3(double sigma, double X, double S, int ns, int64_t *nf, double *h,
                   double *gam) {
  auto [a, b, c] = finufft::common::nhg_type3(sigma, X, S, ns, (int64_t)1e12);
  *nf  = a;
  *h   = b;
  *gam = c;
}

void ref_selfft_params_f32(int ns, double beta, const float *coef, int nc, int stride,
                           double *grid_scale, double *prefac) {
  auto [gs, pf] = finufft::kernel::pswf_selfft_params<float>(ns, beta, coef, nc, stride);
  *grid_scale = gs;
  *prefac     = pf;
}
void ref_selfft_params_f64(int ns, double beta, const double *coef, int nc, int stride,
                           double *grid_scale, double *prefac) {
  auto [gs, pf] = finufft::kernel::pswf_selfft_params<double>(ns, beta, coef, nc, stride);
  *grid_scale = gs;
  *prefac     = pf;
}

}  // extern "C"
