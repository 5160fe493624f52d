// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference FINUFFT (flatironinstitute/finufft @ 9810998d,
// "2.6.0-dev") type-1/2/3 hot path: plan-time kernel maths, setpts (fold-rescale +
// stable bin sort), spread, interp, FFT, deconvolve/shuffle.  Scalar, strict IEEE
// (build with -ffp-contract=off; every fused multiply-add the reference writes as
// fma is an explicit std::fma here).  Every function cites the reference file:line
// it follows (paths relative to /root/reference).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (finufft_b200/) never links or calls it.
//
// Parity status: PINNED to the reference's own code run here.
//  * Plan-time maths (ns, beta, PSWF, Horner fit, Gauss-Legendre, next235): against the
//    reference's src/common/*.cpp compiled into oracle/_ref (oracle/ref_shim.cpp) and the
//    known-answer table test/testutils.cpp:39-54.
//  * Sort permutation, spread, interp, deconvolve, type-3 set-up, the guru driver: against
//    oracle/_ref/libfinufft_ref.so = the reference's src/*.cpp + include/finufft/*.hpp
//    compiled where they lie (oracle/build.py::build_ref_library), whose outputs are committed
//    as tests/golden/reference_vectors.npz.  tests/test_reference_pin_cpu.py: permutation bit
//    for bit, outputs to 1e-13 (double) / 5e-6 (single) relative l2, every golden case.
//    The reference's three un-vendored third-party dependencies are replaced in that build by
//    the stand-ins under oracle/shim/ (xsimd: SIMD wrapper, lane order only; POET: dispatch,
//    no arithmetic; FFTW: the FFT below), so what is pinned is the reference's arithmetic up
//    to SIMD summation order and FFT rounding - the same caveat the reference's own builds
//    have between FFTW and DUCC0.
//  * Direct sums: the reference's test/utils/dirft*.hpp (oracle/ref_dirft_shim.cpp).
//
// The FFT is a third-party dependency of the reference (FFTW 3.3.10 / ducc0_0_41_1,
// CMakeLists.txt:73-76), absent from /root/reference; a plain mixed-radix (2,3,5)
// complex FFT is used instead (same unnormalised definition, src/fft.cpp:266-371).

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>

#include "fft_standin.hpp"
#endif

namespace orc {

using i64 = int64_t;
static constexpr double PI      = 3.141592653589793238462643383279502884;  // constants.h:25
static constexpr double INV_2PI = 0.159154943091895335768883763372514362;  // constants.h:27
static constexpr int MAX_NQUAD  = 100;                                      // constants.h:22
static constexpr int MIN_NC = 4, MAX_NC = 19;                               // constants.h:30-31

// error codes, include/finufft_errors.h:9-44
enum {
  ERR_MAXNALLOC = 2, ERR_SPREAD_BOX_SMALL = 3, ERR_UPSAMPFAC_TOO_SMALL = 7,
  ERR_NTRANS = 9, ERR_TYPE = 10, ERR_DIM = 12, ERR_NUM_NU_PTS = 20,
  ERR_EPS_TOO_SMALL = 26, ERR_PSWF_SETUP = 27
};

// ----------------------------------------------------------------------------------
// next235: smallest 2,3,5-smooth multiple of `fac` that is >= n.
// src/common/utils.cpp:84-122 (the reference borrows ducc0's good_size search; any
// correct search gives the same integer, so a direct scan is used here).
static bool smooth235(i64 n) {
  for (int p : {2, 3, 5})
    while (n % p == 0) n /= p;
  return n == 1;
}
i64 next235(i64 n, i64 fac) {
  n   = std::max<i64>(n, 1);
  fac = std::max<i64>(fac, 1);
  i64 q = (n + fac - 1) / fac;
  if (q < 1) q = 1;
  while (!smooth235(q)) ++q;
  return q * fac;
}

// ----------------------------------------------------------------------------------
// Gauss-Legendre nodes/weights. src/common/utils.cpp:18-79 (Newton from Chebyshev
// guesses; stop once |dx|<1e-14 has been seen three times).
static void leg_eval(int n, double x, double &p, double &dp) {
  if (n == 0) { p = 1.0; dp = 0.0; return; }
  if (n == 1) { p = x; dp = 1.0; return; }
  double p0 = 0.0, p1 = 1.0, p2 = x;
  for (int i = 1; i < n; i++) {
    p0 = p1;
    p1 = p2;
    p2 = ((2 * i + 1) * x * p1 - i * p0) / (i + 1);
  }
  p  = p2;
  dp = n * (x * p2 - p1) / (x * x - 1);
}
void gaussquad(int n, double *xgl, double *wgl) {
  xgl[n / 2] = 0;
  for (int i = 0; i < n / 2; i++) {
    int hits = 0;
    double x = std::cos((2 * i + 1) * PI / (2 * n));
    for (;;) {
      double p, dp;
      leg_eval(n, x, p, dp);
      double dx = -p / dp;
      x += dx;
      if (std::abs(dx) < 1e-14) ++hits;
      if (hits == 3) break;
    }
    xgl[i]         = -x;
    xgl[n - i - 1] = x;
  }
  for (int i = 0; i < n / 2 + 1; i++) {
    double p, dp, junk;
    leg_eval(n, xgl[i], junk, dp);
    leg_eval(n + 1, xgl[i], p, junk);
    wgl[i]         = -2 / ((n + 1) * dp * p);
    wgl[n - i - 1] = wgl[i];
  }
}

// ----------------------------------------------------------------------------------
// Order-zero prolate spheroidal wavefunction Psi_0^c on [-1,1], Psi(0)=1.
// src/common/pswf.cpp:20-191 and include/finufft_common/pswf.h:25-61: even-degree
// Legendre expansion; symmetric tridiagonal eigenproblem (QL with implicit shifts) for
// the eigenvalue, then 4 steps of inverse iteration for the coefficient vector.
struct Pswf0 {
  std::vector<double> w;                       // Legendre coefficients (workdata)
  std::vector<std::array<double, 3>> rec;      // recurrence factors (coef)
  double inv0 = 1.0;
  int err     = 0;

  static void tri_entries(double lam, int k, double c, double &a, double &b, double &g) {
    // pswf.cpp:22-33 (prolcoef)
    double kf = k;
    double a0 = kf * (kf - 1.) / ((2. * kf + 1.) * (2. * kf - 1.));
    double b0 = ((kf + 1.) * (kf + 1.) / (2. * kf + 3.) + kf * kf / (2. * kf - 1.)) /
                (2. * kf + 1.);
    double g0 = (kf + 1.) * (kf + 2.) / ((2. * kf + 1.) * (2. * kf + 3.));
    a = -c * c * a0;
    b = lam - kf * (kf + 1.) - c * c * b0;
    g = -c * c * g0;
  }
  static void fill_matrix(std::vector<double> &as, std::vector<double> &bs,
                          std::vector<double> &cs, int n, double c, double lam) {
    // pswf.cpp:36-45 (prolmatr): symmetrising scale factors
    for (int k = 0; 2 * k <= n + 2; ++k) {
      tri_entries(lam, 2 * k, c, as[k], bs[k], cs[k]);
      if (k != 0) as[k] *= std::sqrt((2 * k + .5) / (2 * k - 1.5));
      cs[k] *= std::sqrt((2 * k + .5) / (2 * k + 2.5));
    }
  }
  int ql_eigenvalues(int n, std::vector<double> &d, std::vector<double> &e) {
    // pswf.cpp:47-99 (prolql1): EISPACK-style tql1, eigenvalues sorted ascending
    if (n == 1) return 0;
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
      int iter = 0;
      for (;;) {
        int m;
        for (m = l; m < n - 1; ++m) {
          double t1 = std::abs(d[m]) + std::abs(d[m + 1]);
          double t2 = t1 + std::abs(e[m]);
          if (t2 == t1) break;
        }
        if (m == l) break;
        if (iter == 30) return ERR_PSWF_SETUP;
        ++iter;
        double g = (d[l + 1] - d[l]) / (2. * e[l]);
        double r = std::sqrt(g * g + 1.0);
        g        = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        for (int i = m - 1; i >= l; --i) {
          double f = s * e[i];
          double b = c * e[i];
          r        = std::sqrt(f * f + g * g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s        = f / r;
          c        = g / r;
          g        = d[i + 1] - p;
          r        = (d[i] - g) * s + 2. * c * b;
          p        = s * r;
          d[i + 1] = g + p;
          g        = c * r - b;
        }
        if (r == 0.) break;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
      for (int i = l; (i > 0) && (d[i] < d[i - 1]); --i) std::swap(d[i], d[i - 1]);
    }
    return 0;
  }
  explicit Pswf0(double c) {
    // pswf.cpp:167-176 (prolps0i): expansion length from table indexed by floor(c/10)
    static const int tab[20] = {48,  64,  80,  92,  106, 120, 130, 144, 156, 168,
                                178, 190, 202, 214, 224, 236, 248, 258, 268, 280};
    int i = (int)(c / 10);
    int n = (i < 20) ? tab[i] : (int)(c * 3) / 2;
    // pswf.cpp:128-165 (prolfun0)
    const double delta = 1.0e-8, eps = 1e-16;
    int h = n / 2;
    std::vector<double> xk(h + 3, 1.0), as(h + 2), bs(h + 2), cs(h + 2), u(h + 2),
        v(h + 2), ww(h + 2);
    fill_matrix(as, bs, cs, n, c, 0.);
    err = ql_eigenvalues(h, bs, as);
    if (err) return;
    double lam = -bs[h - 1] + delta;
    fill_matrix(as, bs, cs, n, c, lam);
    // pswf.cpp:101-113 (prolfact): LU of the shifted tridiagonal; args (a=bs,b=cs,c=as)
    for (int k = 0; k + 1 < h; ++k) {
      double dd  = as[k + 1] / bs[k];
      bs[k + 1] -= cs[k] * dd;
      u[k]      = dd;
      v[k + 1]  = cs[k] / bs[k + 1];
      ww[k + 1] = 1. / bs[k + 1];
    }
    ww[0] = 1. / bs[0];
    for (int it = 0; it < 4; ++it) {
      // pswf.cpp:115-126 (prolsolv)
      for (int k = 0; k + 1 < h; ++k) xk[k + 1] -= u[k] * xk[k];
      for (int k = h - 1; k > 0; --k) {
        xk[k - 1] -= xk[k] * v[k];
        xk[k] *= ww[k];
      }
      xk[0] *= ww[0];
      double nrm = 0;
      for (int j = 0; j < h; ++j) nrm += xk[j] * xk[j];
      nrm = std::sqrt(nrm);
      for (int j = 0; j < h; ++j) xk[j] /= nrm;
    }
    int imax = 0;
    for (int k = 0; k < h; ++k) {
      if (std::abs(xk[k]) > eps) imax = k;
      xk[k] *= std::sqrt(k * 2 + .5);
    }
    xk.resize(imax + 1);
    w = xk;
    // pswf.cpp:180-191 (ctor): even-Legendre two-step recurrence factors
    rec.resize(w.size());
    for (size_t k = 1; k < rec.size(); ++k) {
      double l  = 2 * k - 1.;
      rec[k][0] = ((2. * l - 1.) * (2. * l + 1.)) / (l * (l + 1.));
      rec[k][1] = ((2. * l + 1.) * (l - 1.) * (l - 1.) + l * l * (2. * l - 3)) /
                  (l * (l + 1.) * (2. * l - 3.));
      rec[k][2] = ((2. * l + 1.) * (l - 1.) * (l - 2.)) / (l * (l + 1.) * (2. * l - 3.));
    }
    inv0 = 1. / raw(0.);
  }
  double raw(double x) const {
    // pswf.h:32-52 (eval_raw)
    const double xsq = x * x;
    double pm1 = 0, pm2 = 1, val = w[0];
    size_t i = 1;
    for (; i + 1 < rec.size(); i += 2) {
      pm1 = pm2 * (xsq * rec[i][0] - rec[i][1]) - pm1 * rec[i][2];
      val += w[i] * pm1;
      pm2 = pm1 * (xsq * rec[i + 1][0] - rec[i + 1][1]) - pm2 * rec[i + 1][2];
      val += w[i + 1] * pm2;
    }
    for (; i < rec.size(); ++i) {
      double t = pm2 * (xsq * rec[i][0] - rec[i][1]) - pm1 * rec[i][2];
      val += w[i] * t;
      pm1 = pm2;
      pm2 = t;
    }
    return val;
  }
  double operator()(double x) const {  // pswf.h:57-60 and kernel.cpp:47-51
    if (std::abs(x) > 1) return 0.;
    return raw(x) * inv0;
  }
};

// ----------------------------------------------------------------------------------
// kernel width / shape. src/common/kernel.cpp:60-146, include/finufft_common/kernel.h:85-95,
// include/finufft/makeplan.hpp:112-201.
static double tolfac(int dim, int type) {  // kernel.cpp:60-77
  double r = 0.18;
  for (int i = 0; i < dim - 1; ++i) r *= 1.4;
  return r * (type == 3 ? 1.4 : 1.0);
}
static int theory_ns(double tol, int dim, int type, double sigma) {  // kernel.cpp:79-94
  return (int)std::ceil(std::log(tolfac(dim, type) / tol) / (PI * std::sqrt(1.0 - 1.0 / sigma)) +
                        1.0);
}
template<class T> constexpr int max_ns() { return std::is_same<T, float>::value ? 12 : 16; }
template<class T>
int kernel_setup(double tol_in, int dim, int type, double sigma, int allow_small, int &ns,
                 double &beta, double &tol_used) {
  if (sigma <= 1.0) return ERR_UPSAMPFAC_TOO_SMALL;                 // makeplan.hpp:138-142
  const double eps = std::numeric_limits<T>::epsilon();
  T tol = (T)tol_in;
  if (tol < (T)eps) {                                               // makeplan.hpp:154-162
    if (allow_small) tol = (T)eps;
    else return ERR_EPS_TOO_SMALL;
  }
  int nst = theory_ns((double)tol, dim, type, sigma);
  if (nst > max_ns<T>() && !allow_small) return ERR_EPS_TOO_SMALL;  // makeplan.hpp:171-178
  ns = std::max(2, nst);                                            // kernel.h:85-95
  ns = std::min(ns, max_ns<T>());
  if (std::is_same<T, float>::value && sigma < 1.4) ns = std::min(ns, 8);
  beta     = PI * (double)ns * (1.0 - 1.0 / (2.0 * sigma)) - 0.05;  // kernel.cpp:107,122 (kf=8)
  tol_used = (double)tol;
  return 0;
}

// ----------------------------------------------------------------------------------
// Polynomial fit on Chebyshev nodes -> monomial coefficients, highest degree first.
// include/finufft_common/kernel.h:19-67, all arithmetic in T except the node cosine
// (T*double promotes to double there).
template<class T, class F> std::vector<T> poly_fit(F &&f, int n) {
  std::vector<T> t(n), y(n);
  for (int k = 0; k < n; ++k) {
    t[k] = std::cos((T(2 * k + 1) * PI) / (T(2) * T(n)));
    y[k] = static_cast<T>(f(t[k]));
  }
  std::vector<T> dd = y;
  for (int j = 1; j < n; ++j)
    for (int i = n - 1; i >= j; --i) dd[i] = (dd[i] - dd[i - 1]) / (t[i] - t[i - j]);
  std::vector<T> c(n, T(0));
  std::vector<T> basis{T(1)};
  c[0] += dd[0];
  for (int j = 1; j < n; ++j) {
    std::vector<T> r(basis.size() + 1, T(0));
    for (size_t i = 0; i < basis.size(); ++i) {
      r[i] += -t[j - 1] * basis[i];
      r[i + 1] += basis[i];
    }
    basis = r;
    for (size_t m = 0; m < basis.size(); ++m) c[m] += dd[j] * basis[m];
  }
  std::reverse(c.begin(), c.end());
  return c;
}

// Piecewise-polynomial (Horner) table of the kernel. makeplan.hpp:204-315.
// Output layout coef[k*ns + j], k=0 highest degree of panel j, nc rows.
template<class T>
int horner_table(int ns, double beta, T tol, std::vector<T> &coef, int &nc) {
  const int nc_fit = std::min(MAX_NC, ns + 3);                      // kernel.h:131-133
  Pswf0 psi(beta);
  if (psi.err) return psi.err;
  std::vector<T> tab((size_t)nc_fit * ns, T(0));
  nc = MIN_NC;
  for (int j = 0; j < ns; ++j) {
    const T shift = T(2 * j + 1 - ns);                              // makeplan.hpp:240
    auto fj       = [&](T x) -> T {
      const T z = (x + shift) / (T)ns;
      return (T)psi((double)z);
    };
    std::vector<T> cj = poly_fit<T>(fj, nc_fit);
    for (int k = 0; k < nc_fit; ++k) tab[(size_t)k * ns + j] = cj[k];
    const T cutoff = 0.05;                                          // makeplan.hpp:263
    int need       = 0;
    for (int k = 0; k < nc_fit; ++k)
      if (std::abs(cj[k]) >= tol * cutoff) {
        need = nc_fit - k;
        break;
      }
    if (need > nc) nc = need;
  }
  nc = std::max(nc, std::max(MIN_NC, ns - 4));                      // makeplan.hpp:277, kernel.h:128-130
  coef.assign((size_t)nc * ns, T(0));
  const int shift = nc_fit - nc;                                    // makeplan.hpp:283-300
  for (int k = 0; k < nc; ++k)
    for (int j = 0; j < ns; ++j) coef[(size_t)k * ns + j] = tab[(size_t)(k + shift) * ns + j];
  return 0;
}

// Scalar kernel evaluator, argument in grid units on [-ns/2, ns/2].
// include/finufft/spreadinterp.hpp:57-94.
template<class T> T eval_kernel(T x, int ns, int nc, const T *coef) {
  const T ns2 = ns / T(2.0);
  T res       = T(0.0);
  for (int i = 0; i < ns; ++i) {
    if (x > -ns2 + i && x <= -ns2 + i + 1) {
      T z = std::fma(T(2.0), x - T(i), T(ns - 1));
      for (int j = 0; j < nc; ++j) res = std::fma(res, z, coef[(size_t)j * ns + i]);
      break;
    }
  }
  return res;
}

// All ns kernel values for the stencil whose leftmost cell has offset x1 in
// [-ns/2,-ns/2+1].  Scalar semantics of include/finufft/simd.hpp:327-436 (plain Horner
// branch :424-432; the SIMD even/odd variant :373-422 differs only in rounding).
template<class T> inline void eval_stencil(T x1, int ns, int nc, const T *coef, T *ker) {
  const T z = std::fma(T(2.0), x1, T(ns - 1));
  for (int i = 0; i < ns; ++i) {
    T k = coef[i];
    for (int j = 1; j < nc; ++j) k = std::fma(k, z, coef[(size_t)j * ns + i]);
    ker[i] = k;
  }
}

// Fourier series of the kernel on the fine grid, k=0..nf/2.
// include/finufft/makeplan.hpp:39-108 (single-chunk case nt=1: phase rotators start at 1).
template<class T>
void fseries(i64 nf, int ns, int nc, const T *coef, T *out) {
  T J2  = ns / 2.0;
  int q = (int)(2 + 3.0 * J2);
  T f[MAX_NQUAD];
  double z[2 * MAX_NQUAD], w[2 * MAX_NQUAD];
  gaussquad(2 * q, z, w);
  std::complex<T> a[MAX_NQUAD], aj[MAX_NQUAD];
  for (int n = 0; n < q; ++n) {
    z[n] *= J2;
    f[n] = J2 * (T)w[n] * eval_kernel<T>(T(z[n]), ns, nc, coef);
    a[n] = -std::exp(2 * PI * std::complex<double>(0, 1) * z[n] / double(nf));
  }
  for (int n = 0; n < q; ++n) aj[n] = std::pow(a[n], (T)0);
  i64 nout = nf / 2 + 1;
  for (i64 j = 0; j < nout; ++j) {
    T x = 0.0;
    for (int n = 0; n < q; ++n) {
      x += f[n] * 2 * std::real(aj[n]);
      aj[n] *= a[n];
    }
    out[j] = x;
  }
}

// Analytic PSWF self-FT parameters (type 3).  include/finufft_common/kernel.h:114-123.
template<class T>
void selfft_params(int ns, double beta, const T *coef, int nc, double &grid_scale,
                   double &prefac) {
  prefac = 0;
  for (int i = 0; i < ns; ++i)
    for (int j = nc - 1; j >= 0; j -= 2) prefac += double(coef[(size_t)j * ns + i]) / (nc - j);
  const double J2 = ns / 2.0;
  grid_scale      = J2 * J2 / beta;
}

// Type-3 fine grid size / spacing / rescale. include/finufft_common/kernel.h:165-181.
void nhg_type3(double sigma, double X, double S, int ns, i64 max_nf, i64 &nf, double &h,
               double &gam) {
  const int nss = ns + 1;
  double Xs = X, Ss = S;
  if (Xs == 0.0) {
    if (Ss == 0.0) Xs = Ss = 1.0;
    else Xs = 1.0 / Ss;
  } else
    Ss = std::max(Ss, 1.0 / Xs);
  double nfd = 2.0 * sigma * Ss * Xs / PI + nss;
  if (!std::isfinite(nfd)) nfd = 0.0;
  nf = std::max((i64)nfd, (i64)(2 * ns));
  if (nf < max_nf) nf = next235(nf, 2);
  h   = 2.0 * PI / (double)nf;
  gam = (double)nf / (2.0 * sigma * Ss);
}

// lowest_sigma / check_sigma rule. src/common/kernel.cpp:151-201, setpts.hpp:29-53.
static double smallest_sigma_for_ns(double tol, int dim, int type, int ns_target) {
  const double tf = tolfac(dim, type);
  if (tol <= 0) return 2.5;
  if (tol >= tf) return 1.0 + 0.01;
  const double u = std::log(tf / tol) / ((ns_target - 1.0) * PI);
  if (u >= 1.0) return 2.5;
  return std::min(1.0 / (1.0 - u * u), 2.5);
}
double lowest_sigma(double tol, int dim, int ns, double eps_mach, double gridlen) {
  const double eps_round = 0.48 * eps_mach * gridlen;
  const double r         = tol / eps_round;
  if (r <= 0.5) return 2.0;
  const double pure = smallest_sigma_for_ns(tol, dim, 1, ns);
  if (r >= 10.0) return std::min(pure, 2.0);
  const double a2 = ns > 8 ? 0.014 : 0.555, a1 = ns > 8 ? 0.291 : -0.290,
               a0 = ns > 8 ? -0.043 : 0.071;
  const double ir = 1.0 / r;
  const double corr = (a2 * ir + a1) * ir + a0;
  return std::min(pure + std::max(corr, 0.0), 2.0);
}

// ----------------------------------------------------------------------------------
// fold_rescale: x -> [0,N]. include/finufft/simd.hpp:318-325.
template<class T> inline T fold_rescale(T x, i64 N) {
  const T r = std::fma(x, T(INV_2PI), T(0.5));
  return (r - std::floor(r)) * T(N);
}

// Bin index of every point (exposed for tests) and the stable counting sort.
// include/finufft/spread.hpp:459-584 with the fixed 16x4x4 bins of
// include/finufft/spreadinterp.hpp:159.
template<class T> struct BinGeom {
  i64 nb1, nb2, nb3;
  T inv1, inv2, inv3;
  BinGeom(int ndims, i64 N1, i64 N2, i64 N3) {
    const double bx = 16, by = 4, bz = 4;
    nb1  = i64(T(N1) / bx + 1);                                     // spread.hpp:515-517
    nb2  = ndims > 1 ? i64(T(N2) / by + 1) : 1;
    nb3  = ndims > 2 ? i64(T(N3) / bz + 1) : 1;
    inv1 = T(1.0 / bx);
    inv2 = T(1.0 / by);
    inv3 = T(1.0 / bz);
  }
};
template<class T>
inline i64 bin_of(const BinGeom<T> &g, int ndims, i64 N1, i64 N2, i64 N3, T x, T y, T z) {
  i64 bin = i64(fold_rescale<T>(x, N1) * g.inv1);                   // spread.hpp:544-551
  if (ndims > 1) bin += g.nb1 * i64(fold_rescale<T>(y, N2) * g.inv2);
  if (ndims > 2) bin += g.nb1 * g.nb2 * i64(fold_rescale<T>(z, N3) * g.inv3);
  return bin;
}
template<class T>
void bin_sort(i64 M, const T *kx, const T *ky, const T *kz, i64 N1, i64 N2, i64 N3, i64 *ret,
              i64 *bins_out /*may be null*/) {
  const int ndims = 1 + (N2 > 1) + (N3 > 1);                        // simd.hpp:303-309
  BinGeom<T> g(ndims, N1, N2, N3);
  const i64 nbins = g.nb1 * g.nb2 * g.nb3;
  std::vector<uint32_t> counts(nbins, 0);
  auto bin_i = [&](i64 i) {
    return bin_of<T>(g, ndims, N1, N2, N3, kx[i], ndims > 1 ? ky[i] : T(0),
                     ndims > 2 ? kz[i] : T(0));
  };
  for (i64 i = 0; i < M; i++) {
    i64 b = bin_i(i);
    if (bins_out) bins_out[i] = b;
    ++counts[b];
  }
  uint32_t run = 0;                                                 // spread.hpp:568
  for (i64 b = 0; b < nbins; ++b) {
    uint32_t c = counts[b];
    counts[b]  = run;
    run += c;
  }
  for (i64 i = 0; i < M; i++) {                                     // spread.hpp:571-583
    i64 b          = bin_i(i);
    ret[counts[b]] = i;
    ++counts[b];
  }
}

// ----------------------------------------------------------------------------------
// Type-1 spreading. include/finufft/spreadinterp.hpp:310-485 (subproblem driver),
// include/finufft/spread.hpp:53-384 (subgrid kernels), :386-451 (wrapped add),
// :786-864 (subgrid box).  One subproblem = a contiguous chunk of the sorted points.
template<class T> struct KerTab {
  int ns, nc;
  const T *coef;
};

template<class T>
static void spread_chunk(int ndims, const KerTab<T> &K, i64 M0, const T *kx, const T *ky,
                         const T *kz, const T *dd, i64 N1, i64 N2, i64 N3, T *out,
                         bool atomic_add) {
  const int ns = K.ns;
  const T ns2  = (T)ns / 2;
  // get_subgrid, spread.hpp:836-863
  auto range = [&](const T *a, i64 &off, i64 &size) {
    T lo = a[0], hi = a[0];
    for (i64 i = 1; i < M0; ++i) {
      lo = std::min(lo, a[i]);
      hi = std::max(hi, a[i]);
    }
    off  = (i64)std::ceil(lo - ns2);
    size = (i64)std::ceil(hi - ns2) - off + ns;
  };
  i64 o1, o2 = 0, o3 = 0, s1, s2 = 1, s3 = 1;
  range(kx, o1, s1);
  if (ndims > 1) range(ky, o2, s2);
  if (ndims > 2) range(kz, o3, s3);
  std::vector<T> du((size_t)2 * s1 * s2 * s3, T(0));
  std::vector<T> k1(ns), k2(ns), k3(ns), k1v(2 * ns);
  for (i64 p = 0; p < M0; ++p) {
    const T re = dd[2 * p], im = dd[2 * p + 1];
    const i64 i1 = (i64)std::ceil(kx[p] - ns2);                     // spread.hpp:328-333
    T x1         = std::ceil(kx[p] - ns2) - kx[p];
    if (ndims == 1) {                                               // spread.hpp:116-121
      if (x1 < -ns2) x1 = -ns2;
      if (x1 > -ns2 + 1) x1 = -ns2 + 1;
    }
    eval_stencil<T>(x1, ns, K.nc, K.coef, k1.data());
    for (int t = 0; t < ns; ++t) {                                  // ker1val = ker1 (x) (re,im)
      k1v[2 * t]     = k1[t] * re;
      k1v[2 * t + 1] = k1[t] * im;
    }
    if (ndims == 1) {
      T *trg = du.data() + 2 * (i1 - o1);
      for (int l = 0; l < 2 * ns; ++l) trg[l] += k1v[l];           // spread.hpp:136-198 (add of ker*dd)
    } else if (ndims == 2) {
      const i64 i2 = (i64)std::ceil(ky[p] - ns2);
      const T x2   = std::ceil(ky[p] - ns2) - ky[p];
      eval_stencil<T>(x2, ns, K.nc, K.coef, k2.data());
      for (int dy = 0; dy < ns; ++dy) {                             // spread.hpp:284-298
        T *trg     = du.data() + 2 * (s1 * (i2 - o2 + dy) + i1 - o1);
        const T kv = k2[dy];
        for (int l = 0; l < 2 * ns; ++l) trg[l] = std::fma(kv, k1v[l], trg[l]);
      }
    } else {
      const i64 i2 = (i64)std::ceil(ky[p] - ns2), i3 = (i64)std::ceil(kz[p] - ns2);
      const T x2 = std::ceil(ky[p] - ns2) - ky[p], x3 = std::ceil(kz[p] - ns2) - kz[p];
      eval_stencil<T>(x2, ns, K.nc, K.coef, k2.data());
      eval_stencil<T>(x3, ns, K.nc, K.coef, k3.data());
      for (int dz = 0; dz < ns; ++dz) {                             // spread.hpp:369-383
        const i64 oz = s1 * s2 * (i3 - o3 + dz);
        for (int dy = 0; dy < ns; ++dy) {
          T *trg     = du.data() + 2 * (oz + s1 * (i2 - o2 + dy) + i1 - o1);
          const T kv = k2[dy] * k3[dz];
          for (int l = 0; l < 2 * ns; ++l) trg[l] = std::fma(kv, k1v[l], trg[l]);
        }
      }
    }
  }
  // add_wrapped_subgrid, spread.hpp:386-451
  std::vector<i64> w2(s2), w3(s3);
  i64 y = o2, z = o3;
  for (i64 i = 0; i < s2; ++i) {
    if (y < 0) y += N2;
    if (y >= N2) y -= N2;
    w2[i] = y++;
  }
  for (i64 i = 0; i < s3; ++i) {
    if (z < 0) z += N3;
    if (z >= N3) z -= N3;
    w3[i] = z++;
  }
  auto acc = [&](T &a, T b) {
    if (atomic_add) {
#pragma omp atomic
      a += b;
    } else
      a += b;
  };
  for (i64 dz = 0; dz < s3; dz++)
    for (i64 dy = 0; dy < s2; dy++) {
      T *orow       = out + 2 * (N1 * w2[dy] + N1 * N2 * w3[dz]);
      const T *irow = du.data() + 2 * s1 * (dy + s2 * dz);
      for (i64 dx = 0; dx < s1; ++dx) {
        i64 xx = o1 + dx;
        if (xx < 0) xx += N1;
        if (xx >= N1) xx -= N1;
        acc(orow[2 * xx], irow[2 * dx]);
        acc(orow[2 * xx + 1], irow[2 * dx + 1]);
      }
    }
}

template<class T>
void spread_sorted(int ndims, i64 N1, i64 N2, i64 N3, i64 M, const T *kx, const T *ky,
                   const T *kz, const T *c, const i64 *perm, const KerTab<T> &K, T *fw,
                   int nthr, i64 max_sub) {
  const i64 N = N1 * N2 * N3;
  std::fill(fw, fw + 2 * N, T(0));                                  // spreadinterp.hpp:345-352
  if (M == 0) return;
  if (nthr < 1) nthr = 1;
  // number of subproblems, spreadinterp.hpp:374-394
  i64 nb = std::min<i64>(nthr, M);
  if (nb * max_sub < M) nb = 1 + (M - 1) / max_sub;
  if (M * 1000 < N) nb = M;
  std::vector<i64> brk(nb + 1);
  for (i64 p = 0; p <= nb; ++p) brk[p] = (M * p + nb - 1) / nb;
  const bool guard = nb > 1 && nthr > 1;
#pragma omp parallel num_threads(nthr)
  {
    std::vector<T> x0, y0, z0, d0;
#pragma omp for schedule(dynamic, 1)
    for (i64 s = 0; s < nb; ++s) {
      const i64 M0 = brk[s + 1] - brk[s];
      if (M0 <= 0) continue;
      x0.resize(M0);
      y0.resize(ndims > 1 ? M0 : 0);
      z0.resize(ndims > 2 ? M0 : 0);
      d0.resize(2 * M0);
      for (i64 j = 0; j < M0; ++j) {                                // spreadinterp.hpp:430-438
        const i64 kk = perm[j + brk[s]];
        x0[j]        = fold_rescale<T>(kx[kk], N1);
        if (ndims > 1) y0[j] = fold_rescale<T>(ky[kk], N2);
        if (ndims > 2) z0[j] = fold_rescale<T>(kz[kk], N3);
        d0[2 * j]     = c[2 * kk];
        d0[2 * j + 1] = c[2 * kk + 1];
      }
      spread_chunk<T>(ndims, K, M0, x0.data(), y0.data(), z0.data(), d0.data(), N1, N2, N3,
                      fw, guard);
    }
  }
}

// ----------------------------------------------------------------------------------
// Type-2 interpolation. include/finufft/interp.hpp:457-556 (driver) and the wrapped
// variants :11-130 (line), :132-279 (square), :281-355 (cube): accumulate a line over
// (dy,dz) with fma, then dot with ker1 by fma.
template<class T>
void interp_sorted(int ndims, i64 N1, i64 N2, i64 N3, i64 M, const T *kx, const T *ky,
                   const T *kz, T *c, const i64 *perm, const KerTab<T> &K, const T *fw,
                   int nthr) {
  const int ns = K.ns;
  const T ns2  = ns * T(0.5);
  if (nthr < 1) nthr = 1;
#pragma omp parallel num_threads(nthr)
  {
    std::vector<T> k1(ns), k2(ns), k3(ns), line(2 * ns);
    std::vector<i64> j1(ns), j2(ns), j3(ns);
#pragma omp for schedule(dynamic, 4096)
    for (i64 i = 0; i < M; ++i) {
      const i64 j = perm[i];
      const T xj  = fold_rescale<T>(kx[j], N1);
      const i64 i1 = (i64)std::ceil(xj - ns2);
      const T x1   = std::ceil(xj - ns2) - xj;
      eval_stencil<T>(x1, ns, K.nc, K.coef, k1.data());
      i64 i2 = 0, i3 = 0;
      if (ndims > 1) {
        const T yj = fold_rescale<T>(ky[j], N2);
        i2         = (i64)std::ceil(yj - ns2);
        eval_stencil<T>(std::ceil(yj - ns2) - yj, ns, K.nc, K.coef, k2.data());
      }
      if (ndims > 2) {
        const T zj = fold_rescale<T>(kz[j], N3);
        i3         = (i64)std::ceil(zj - ns2);
        eval_stencil<T>(std::ceil(zj - ns2) - zj, ns, K.nc, K.coef, k3.data());
      }
      i64 x = i1, y = i2, z = i3;                                   // interp.hpp:324-337
      for (int d = 0; d < ns; d++) {
        if (x < 0) x += N1;
        if (x >= N1) x -= N1;
        j1[d] = x++;
        if (y < 0) y += N2;
        if (y >= N2) y -= N2;
        j2[d] = y++;
        if (z < 0) z += N3;
        if (z >= N3) z -= N3;
        j3[d] = z++;
      }
      std::fill(line.begin(), line.end(), T(0));
      const int nz = ndims > 2 ? ns : 1, ny = ndims > 1 ? ns : 1;
      for (int dz = 0; dz < nz; ++dz)
        for (int dy = 0; dy < ny; ++dy) {
          const i64 row = N1 * ((ndims > 1 ? j2[dy] : 0) + N2 * (ndims > 2 ? j3[dz] : 0));
          const T k23   = (ndims > 2 ? k2[dy] * k3[dz] : (ndims > 1 ? k2[dy] : T(1)));
          for (int dx = 0; dx < ns; ++dx) {
            const T *src      = fw + 2 * (row + j1[dx]);
            line[2 * dx]     = std::fma(src[0], k23, line[2 * dx]);
            line[2 * dx + 1] = std::fma(src[1], k23, line[2 * dx + 1]);
          }
        }
      T o0 = 0, o1 = 0;
      for (int dx = 0; dx < ns; ++dx) {
        o0 = std::fma(line[2 * dx], k1[dx], o0);
        o1 = std::fma(line[2 * dx + 1], k1[dx], o1);
      }
      c[2 * j]     = o0;
      c[2 * j + 1] = o1;
    }
  }
}

// ----------------------------------------------------------------------------------
// Deconvolve + mode shuffle. include/finufft/execute.hpp:69-133 (1d), :135-184 (2d),
// :186-237 (3d).  dir=1: fw -> fk; dir=2: fk -> zero-padded fw.  prefac handed down as
// nested real divisions exactly as the reference does.
template<class T>
static void deconv1d(int dir, T prefac, const T *ker, i64 ms, i64 nf1, int modeord, T *fk,
                     std::complex<T> *fw) {
  i64 kmin = -ms / 2, kmax = (ms - 1) / 2;
  if (ms == 0) kmax = -1;
  i64 pp = -2 * kmin, pn = 0;
  if (modeord == 1) {
    pp = 0;
    pn = 2 * (kmax + 1);
  }
  if (dir == 1) {
    for (i64 k = 0; k <= kmax; ++k) {
      fk[pp++] = prefac * fw[k].real() / ker[k];
      fk[pp++] = prefac * fw[k].imag() / ker[k];
    }
    for (i64 k = kmin; k < 0; ++k) {
      fk[pn++] = prefac * fw[nf1 + k].real() / ker[-k];
      fk[pn++] = prefac * fw[nf1 + k].imag() / ker[-k];
    }
  } else {
    for (i64 k = kmax + 1; k < nf1 + kmin; ++k) fw[k] = 0.0;
    for (i64 k = 0; k <= kmax; ++k) {
      T re = prefac * fk[pp++] / ker[k];
      T im = prefac * fk[pp++] / ker[k];
      fw[k] = {re, im};
    }
    for (i64 k = kmin; k < 0; ++k) {
      T re = prefac * fk[pn++] / ker[-k];
      T im = prefac * fk[pn++] / ker[-k];
      fw[nf1 + k] = {re, im};
    }
  }
}
template<class T>
static void deconv2d(int dir, T prefac, const T *ker1, const T *ker2, i64 ms, i64 mt, i64 nf1,
                     i64 nf2, int modeord, T *fk, std::complex<T> *fw) {
  i64 k2min = -mt / 2, k2max = (mt - 1) / 2;
  if (mt == 0) k2max = -1;
  i64 pp = -2 * k2min * ms, pn = 0;
  if (modeord == 1) {
    pp = 0;
    pn = 2 * (k2max + 1) * ms;
  }
  if (dir == 2)
    for (i64 j = nf1 * (k2max + 1); j < nf1 * (nf2 + k2min); ++j) fw[j] = 0.0;
  for (i64 k2 = 0; k2 <= k2max; ++k2, pp += 2 * ms)
    deconv1d<T>(dir, prefac / ker2[k2], ker1, ms, nf1, modeord, fk + pp, &fw[nf1 * k2]);
  for (i64 k2 = k2min; k2 < 0; ++k2, pn += 2 * ms)
    deconv1d<T>(dir, prefac / ker2[-k2], ker1, ms, nf1, modeord, fk + pn,
                &fw[nf1 * (nf2 + k2)]);
}
template<class T>
static void deconv3d(int dir, T prefac, const T *ker1, const T *ker2, const T *ker3, i64 ms,
                     i64 mt, i64 mu, i64 nf1, i64 nf2, i64 nf3, int modeord, T *fk,
                     std::complex<T> *fw) {
  i64 k3min = -mu / 2, k3max = (mu - 1) / 2;
  if (mu == 0) k3max = -1;
  i64 pp = -2 * k3min * ms * mt, pn = 0;
  if (modeord == 1) {
    pp = 0;
    pn = 2 * (k3max + 1) * ms * mt;
  }
  i64 np = nf1 * nf2;
  if (dir == 2)
    for (i64 j = np * (k3max + 1); j < np * (nf3 + k3min); ++j) fw[j] = 0.0;
  for (i64 k3 = 0; k3 <= k3max; ++k3, pp += 2 * ms * mt)
    deconv2d<T>(dir, prefac / ker3[k3], ker1, ker2, ms, mt, nf1, nf2, modeord, fk + pp,
                &fw[np * k3]);
  for (i64 k3 = k3min; k3 < 0; ++k3, pn += 2 * ms * mt)
    deconv2d<T>(dir, prefac / ker3[-k3], ker1, ker2, ms, mt, nf1, nf2, modeord, fk + pn,
                &fw[np * (nf3 + k3)]);
}
template<class T>
void deconvolve(int dir, int dim, const i64 *ms, const i64 *nf, int modeord, const T *p1,
                const T *p2, const T *p3, T *fk, T *fw) {
  auto *fwc = reinterpret_cast<std::complex<T> *>(fw);
  if (dim == 1) deconv1d<T>(dir, T(1), p1, ms[0], nf[0], modeord, fk, fwc);
  else if (dim == 2) deconv2d<T>(dir, T(1), p1, p2, ms[0], ms[1], nf[0], nf[1], modeord, fk, fwc);
  else
    deconv3d<T>(dir, T(1), p1, p2, p3, ms[0], ms[1], ms[2], nf[0], nf[1], nf[2], modeord, fk,
                fwc);
}

// ----------------------------------------------------------------------------------
// FFT: unnormalised in-place complex DFT with exponent sign `sign`, sizes 2,3,5-smooth
// (any size works, O(n*p) per prime factor p).  Stands in for FFTW/DUCC0
// (src/fft.cpp:266-371; dims slowest-first, :253-261).
// ----------------------------------------------------------------------------------
// Plan object: the subset of FINUFFT_PLAN_T state the hot path needs
// (include/finufft/plan.hpp:106-145) and the guru sequence makeplan / setpts / execute
// (makeplan.hpp:317-465, setpts.hpp:107-321, execute.hpp:318-563).
template<class T> struct Plan {
  int type, dim, ntr, sign, modeord, ns = 0, nc = 0, nthr = 1, spread_only = 0;
  double sigma, beta = 0, tol = 0;
  i64 ms[3] = {1, 1, 1}, nf[3] = {1, 1, 1};
  std::vector<T> coef, phihat[3];
  i64 M = 0;
  const T *x = nullptr, *y = nullptr, *z = nullptr;
  std::vector<i64> perm;
  // type 3
  i64 nk = 0;
  std::vector<T> xp[3], sp[3];
  std::vector<std::complex<T>> prephase, deconv;
  T t3C[3] = {0, 0, 0}, t3D[3] = {0, 0, 0}, t3h[3] = {0, 0, 0}, t3gam[3] = {1, 1, 1};
  Plan<T> *inner = nullptr;
  ~Plan() { delete inner; }
  i64 nftot() const { return nf[0] * nf[1] * nf[2]; }
  i64 nmodes() const { return ms[0] * ms[1] * ms[2]; }
  KerTab<T> ktab() const { return {ns, nc, coef.data()}; }
};

template<class T>
int plan_kernel(Plan<T> &p, double tol, int allow_small) {
  double tol_used;
  int err = kernel_setup<T>(tol, p.dim, p.type, p.sigma, allow_small, p.ns, p.beta, tol_used);
  if (err) return err;
  p.tol = tol_used;
  return horner_table<T>(p.ns, p.beta, (T)tol_used, p.coef, p.nc);
}

template<class T>
int makeplan(int type, int dim, const i64 *nmodes, int iflag, int ntr, double tol, double sigma,
             int modeord, int spread_only, int allow_small, int nthr, Plan<T> **out) {
  *out = nullptr;
  if (type < 1 || type > 3) return ERR_TYPE;                        // makeplan.hpp:347-358
  if (dim < 1 || dim > 3) return ERR_DIM;
  if (ntr < 1) return ERR_NTRANS;
  auto *p        = new Plan<T>();
  p->type        = type;
  p->dim         = dim;
  p->ntr         = ntr;
  p->sign        = (iflag >= 0) ? 1 : -1;                           // makeplan.hpp:361
  p->modeord     = modeord;
  p->sigma       = sigma;
  p->nthr        = nthr;
  p->spread_only = spread_only;
  int err        = plan_kernel<T>(*p, tol, allow_small);
  if (err) {
    delete p;
    return err;
  }
  if (type != 3) {
    for (int d = 0; d < dim; ++d) {
      p->ms[d] = nmodes[d];
      if (spread_only) {                                            // fft.cpp:397-399
        p->nf[d] = nmodes[d];
        continue;
      }
      i64 nf = (i64)std::ceil(sigma * (double)nmodes[d]);           // makeplan.hpp:26-37
      if (nf < 2 * p->ns) nf = 2 * p->ns;
      if (nf >= (i64)1e12) {
        delete p;
        return ERR_MAXNALLOC;
      }
      p->nf[d] = next235(nf, 2);
      p->phihat[d].resize(p->nf[d] / 2 + 1);
      fseries<T>(p->nf[d], p->ns, p->nc, p->coef.data(), p->phihat[d].data());
    }
  }
  *out = p;
  return 0;
}

template<class T>
static bool want_sort(const Plan<T> &p, int dir) {                  // spreadinterp.hpp:161-185 (sort=2)
  return !(p.dim == 1 && (dir == 2 || p.M > 1000 * p.nf[0]));
}

template<class T>
int setpts(Plan<T> &p, i64 M, const T *x, const T *y, const T *z, i64 nk, const T *s,
           const T *t, const T *u, int allow_small, int force_sort) {
  if (M < 0 || M > (i64)1e14) return ERR_NUM_NU_PTS;                // setpts.hpp:118-124
  p.M = M;
  if (p.type != 3) {
    // check_sigma, setpts.hpp:29-53
    const double eps_mach = std::numeric_limits<T>::epsilon();
    const double gridlen  = (double)*std::max_element(p.nf, p.nf + p.dim);
    const double smin     = lowest_sigma(p.tol, p.dim, p.ns, eps_mach, gridlen);
    const bool unachievable = p.tol <= 0.5 * 0.48 * eps_mach * gridlen;
    if ((unachievable || smin > p.sigma) && !allow_small)
      return ERR_EPS_TOO_SMALL;
    p.x = x;
    p.y = y;
    p.z = z;
    for (int d = 0; d < p.dim; ++d)                                 // spreadcheck, spreadinterp.hpp:31-55
      if (p.nf[d] < 2 * p.ns) return ERR_SPREAD_BOX_SMALL;
    p.perm.resize(M);
    if (force_sort || want_sort(p, p.type))
      bin_sort<T>(M, x, y, z, p.nf[0], p.nf[1], p.nf[2], p.perm.data(), nullptr);
    else
      std::iota(p.perm.begin(), p.perm.end(), (i64)0);
    return 0;
  }
  // ---- type 3, setpts.hpp:163-319
  if (nk < 0 || nk > (i64)1e14) return ERR_NUM_NU_PTS;
  p.nk = nk;
  const T *xyz[3] = {x, y, z}, *stu[3] = {s, t, u};
  T X[3] = {0, 0, 0}, S[3] = {0, 0, 0};
  auto widcen = [](i64 n, const T *a, T &w, T &c) {                  // utils.h:77-90
    T lo = INFINITY, hi = -INFINITY;
    for (i64 m = 0; m < n; ++m) {
      if (a[m] < lo) lo = a[m];
      if (a[m] > hi) hi = a[m];
    }
    w = (hi - lo) / 2;
    c = (hi + lo) / 2;
    if (std::abs(c) < 0.1 * w) {
      w += std::abs(c);
      c = 0.0;
    }
  };
  for (int d = 0; d < p.dim; ++d) {
    widcen(M, xyz[d], X[d], p.t3C[d]);
    widcen(nk, stu[d], S[d], p.t3D[d]);
  }
  for (int d = 0; d < p.dim; ++d) {
    double h, gam;
    nhg_type3(p.sigma, X[d], S[d], p.ns, (i64)1e12, p.nf[d], h, gam);
    p.t3h[d]   = T(h);
    p.t3gam[d] = T(gam);
  }
  for (int d = p.dim; d < 3; ++d) p.t3C[d] = p.t3D[d] = 0.0;
  T ig[3] = {0, 0, 0};
  for (int d = 0; d < p.dim; ++d) {
    p.xp[d].resize(M);
    p.sp[d].resize(nk);
    ig[d] = 1.0 / p.t3gam[d];
  }
  for (i64 j = 0; j < M; ++j)
    for (int d = 0; d < p.dim; ++d) p.xp[d][j] = (xyz[d][j] - p.t3C[d]) * ig[d];
  const T isign = (p.sign >= 0) ? 1 : -1;
  p.prephase.resize(M);
  if (p.t3D[0] != 0.0 || p.t3D[1] != 0.0 || p.t3D[2] != 0.0) {
    for (i64 j = 0; j < M; ++j) {
      T ph = 0;
      for (int d = 0; d < p.dim; ++d) ph += p.t3D[d] * xyz[d][j];
      p.prephase[j] = std::polar(T(1), isign * ph);
    }
  } else
    for (i64 j = 0; j < M; ++j) p.prephase[j] = {1.0, 0.0};
  double gs, pf;
  selfft_params<T>(p.ns, p.beta, p.coef.data(), p.nc, gs, pf);
  const T grid_scale = T(gs), prefac = T(pf);
  p.deconv.resize(nk);
  const bool Cfinite  = std::isfinite(p.t3C[0]) && std::isfinite(p.t3C[1]) && std::isfinite(p.t3C[2]);
  const bool Cnonzero = p.t3C[0] != 0.0 || p.t3C[1] != 0.0 || p.t3C[2] != 0.0;
  const bool do_phase = Cfinite && Cnonzero;
  for (i64 k = 0; k < nk; ++k) {
    T ph = 0, phi = 1;
    for (int d = 0; d < p.dim; ++d) {
      T sin_ = stu[d][k];
      T sp   = p.t3h[d] * p.t3gam[d] * (sin_ - p.t3D[d]);
      // Kernel_onedim_FT::operator(), plan.hpp:255-265: prefac*phi(grid_scale*k)
      phi *= prefac * eval_kernel<T>(grid_scale * sp, p.ns, p.nc, p.coef.data());
      if (do_phase) ph += (sin_ - p.t3D[d]) * p.t3C[d];
      p.sp[d][k] = sp;
    }
    p.deconv[k] = do_phase ? std::polar(T(1) / phi, isign * ph) : std::complex<T>(T(1) / phi);
  }
  p.x = p.xp[0].data();
  p.y = p.dim > 1 ? p.xp[1].data() : nullptr;
  p.z = p.dim > 2 ? p.xp[2].data() : nullptr;
  p.perm.resize(M);
  if (force_sort || want_sort(p, 1))
    bin_sort<T>(M, p.x, p.y, p.z, p.nf[0], p.nf[1], p.nf[2], p.perm.data(), nullptr);
  else
    std::iota(p.perm.begin(), p.perm.end(), (i64)0);
  delete p.inner;
  p.inner = nullptr;
  i64 t2modes[3] = {p.nf[0], p.nf[1], p.nf[2]};
  int err = makeplan<T>(2, p.dim, t2modes, p.sign, 1, p.tol, p.sigma, 0, 0, allow_small, p.nthr,
                        &p.inner);
  if (err) return err;
  return setpts<T>(*p.inner, nk, p.sp[0].data(), p.dim > 1 ? p.sp[1].data() : nullptr,
                   p.dim > 2 ? p.sp[2].data() : nullptr, 0, nullptr, nullptr, nullptr,
                   allow_small, force_sort);
}

template<class T> int execute(Plan<T> &p, T *c, T *fk, int adjoint);

template<class T>
static int exec_type12(Plan<T> &p, T *c, T *fk, bool adjoint, int ntr) {
  const i64 G = p.nftot(), Nm = p.nmodes();
  std::vector<T> fwbuf(p.spread_only ? 0 : (size_t)2 * G);
  const bool spreading = (p.type == 1) != adjoint;
  const int fsign      = adjoint ? -p.sign : p.sign;                // fft.cpp:293,366-369
  for (int b = 0; b < ntr; ++b) {                                   // execute.hpp:376-417
    T *cb  = c + (size_t)2 * b * p.M;
    T *fkb = fk + (size_t)2 * b * Nm;
    T *fw  = p.spread_only ? fkb : fwbuf.data();
    if (spreading) {
      spread_sorted<T>(p.dim, p.nf[0], p.nf[1], p.nf[2], p.M, p.x, p.y, p.z, cb, p.perm.data(),
                       p.ktab(), fw, p.nthr, p.dim == 1 ? 10000 : 100000);
      if (p.spread_only) continue;
      fft_nd<T>(p.dim, p.nf, fsign, fw, p.nthr);
      deconvolve<T>(1, p.dim, p.ms, p.nf, p.modeord, p.phihat[0].data(), p.phihat[1].data(),
                    p.phihat[2].data(), fkb, fw);
    } else {
      if (!p.spread_only) {
        deconvolve<T>(2, p.dim, p.ms, p.nf, p.modeord, p.phihat[0].data(), p.phihat[1].data(),
                      p.phihat[2].data(), fkb, fw);
        fft_nd<T>(p.dim, p.nf, fsign, fw, p.nthr);
      }
      interp_sorted<T>(p.dim, p.nf[0], p.nf[1], p.nf[2], p.M, p.x, p.y, p.z, cb, p.perm.data(),
                       p.ktab(), fw, p.nthr);
    }
  }
  return 0;
}

template<class T> int execute(Plan<T> &p, T *c, T *fk, int adjoint) {
  if (p.type != 3) return exec_type12<T>(p, c, fk, adjoint != 0, p.ntr);
  // ---- type 3 (forward only here), execute.hpp:432-558
  if (adjoint) return ERR_TYPE;
  const i64 G = p.nftot();
  std::vector<std::complex<T>> cp(p.M), fw(G);
  auto *cc  = reinterpret_cast<std::complex<T> *>(c);
  auto *fkc = reinterpret_cast<std::complex<T> *>(fk);
  for (int b = 0; b < p.ntr; ++b) {
    for (i64 j = 0; j < p.M; ++j) cp[j] = p.prephase[j] * cc[(size_t)b * p.M + j];
    spread_sorted<T>(p.dim, p.nf[0], p.nf[1], p.nf[2], p.M, p.x, p.y, p.z,
                     reinterpret_cast<T *>(cp.data()), p.perm.data(), p.ktab(),
                     reinterpret_cast<T *>(fw.data()), p.nthr, p.dim == 1 ? 10000 : 100000);
    std::complex<T> *fkb = fkc + (size_t)b * p.nk;
    int err = exec_type12<T>(*p.inner, reinterpret_cast<T *>(fkb),
                             reinterpret_cast<T *>(fw.data()), false, 1);
    if (err) return err;
    for (i64 k = 0; k < p.nk; ++k) fkb[k] *= p.deconv[k];
  }
  return 0;
}

// ----------------------------------------------------------------------------------
// Direct sums (double precision), the reference tests' ground truth:
// test/utils/dirft1d.hpp, dirft2d.hpp, dirft3d.hpp.
void dirft_type1(int dim, i64 M, const double *x, const double *y, const double *z,
                 const double *c, int sign, const i64 *ms, double *fk, int nthr) {
  const i64 N1 = ms[0], N2 = dim > 1 ? ms[1] : 1, N3 = dim > 2 ? ms[2] : 1;
  const i64 k1min = -(N1 / 2), k2min = -(N2 / 2), k3min = -(N3 / 2);
#pragma omp parallel for num_threads(nthr < 1 ? 1 : nthr) schedule(static)
  for (i64 m = 0; m < N1 * N2 * N3; ++m) {
    const i64 k1 = k1min + m % N1, k2 = dim > 1 ? k2min + (m / N1) % N2 : 0,
              k3 = dim > 2 ? k3min + m / (N1 * N2) : 0;
    double re = 0, im = 0;
    for (i64 j = 0; j < M; ++j) {
      double ph = k1 * x[j] + (dim > 1 ? k2 * y[j] : 0.0) + (dim > 2 ? k3 * z[j] : 0.0);
      double cs = std::cos(ph), sn = sign * std::sin(ph);
      re += c[2 * j] * cs - c[2 * j + 1] * sn;
      im += c[2 * j] * sn + c[2 * j + 1] * cs;
    }
    fk[2 * m]     = re;
    fk[2 * m + 1] = im;
  }
}
void dirft_type2(int dim, i64 M, const double *x, const double *y, const double *z, double *c,
                 int sign, const i64 *ms, const double *fk, int nthr) {
  const i64 N1 = ms[0], N2 = dim > 1 ? ms[1] : 1, N3 = dim > 2 ? ms[2] : 1;
  const i64 k1min = -(N1 / 2), k2min = -(N2 / 2), k3min = -(N3 / 2);
#pragma omp parallel for num_threads(nthr < 1 ? 1 : nthr) schedule(static)
  for (i64 j = 0; j < M; ++j) {
    double re = 0, im = 0;
    for (i64 m = 0; m < N1 * N2 * N3; ++m) {
      const i64 k1 = k1min + m % N1, k2 = dim > 1 ? k2min + (m / N1) % N2 : 0,
                k3 = dim > 2 ? k3min + m / (N1 * N2) : 0;
      double ph = k1 * x[j] + (dim > 1 ? k2 * y[j] : 0.0) + (dim > 2 ? k3 * z[j] : 0.0);
      double cs = std::cos(ph), sn = sign * std::sin(ph);
      re += fk[2 * m] * cs - fk[2 * m + 1] * sn;
      im += fk[2 * m] * sn + fk[2 * m + 1] * cs;
    }
    c[2 * j]     = re;
    c[2 * j + 1] = im;
  }
}
void dirft_type3(int dim, i64 M, const double *x, const double *y, const double *z,
                 const double *c, int sign, i64 nk, const double *s, const double *t,
                 const double *u, double *fk, int nthr) {
#pragma omp parallel for num_threads(nthr < 1 ? 1 : nthr) schedule(static)
  for (i64 k = 0; k < nk; ++k) {
    double re = 0, im = 0;
    for (i64 j = 0; j < M; ++j) {
      double ph = s[k] * x[j] + (dim > 1 ? t[k] * y[j] : 0.0) + (dim > 2 ? u[k] * z[j] : 0.0);
      double cs = std::cos(ph), sn = sign * std::sin(ph);
      re += c[2 * j] * cs - c[2 * j + 1] * sn;
      im += c[2 * j] * sn + c[2 * j + 1] * cs;
    }
    fk[2 * k]     = re;
    fk[2 * k + 1] = im;
  }
}

}  // namespace orc

// =====================================================================================
// C ABI for ctypes (tests / bench cpu_baseline only).
// =====================================================================================
using orc::i64;
extern "C" {

i64 orc_next235(i64 n, i64 fac) { return orc::next235(n, fac); }
void orc_gaussquad(int n, double *x, double *w) { orc::gaussquad(n, x, w); }
int orc_pswf(double c, i64 n, const double *x, double *out) {
  orc::Pswf0 psi(c);
  if (psi.err) return psi.err;
  for (i64 i = 0; i < n; ++i) out[i] = psi(x[i]);
  return 0;
}
double orc_lowest_sigma(double tol, int dim, int ns, double eps_mach, double gridlen) {
  return orc::lowest_sigma(tol, dim, ns, eps_mach, gridlen);
}
void orc_nhg_type3(double sigma, double X, double S, int ns, i64 *nf, double *h, double *gam) {
  orc::nhg_type3(sigma, X, S, ns, (i64)1e12, *nf, *h, *gam);
}
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_dirft1(int dim, i64 M, const double *x, const double *y, const double *z,
                const double *c, int sign, const i64 *ms, double *fk, int nthr) {
  orc::dirft_type1(dim, M, x, y, z, c, sign, ms, fk, nthr);
}
void orc_dirft2(int dim, i64 M, const double *x, const double *y, const double *z, double *c,
                int sign, const i64 *ms, const double *fk, int nthr) {
  orc::dirft_type2(dim, M, x, y, z, c, sign, ms, fk, nthr);
}
void orc_dirft3(int dim, i64 M, const double *x, const double *y, const double *z,
                const double *c, int sign, i64 nk, const double *s, const double *t,
                const double *u, double *fk, int nthr) {
  orc::dirft_type3(dim, M, x, y, z, c, sign, nk, s, t, u, fk, nthr);
}

#define ORC_API(SUF, T)                                                                        \
  int orc_kernel_setup_##SUF(double tol, int dim, int type, double sigma, int allow_small,    \
                             int *ns, double *beta, double *tol_used) {                       \
    return orc::kernel_setup<T>(tol, dim, type, sigma, allow_small, *ns, *beta, *tol_used);   \
  }                                                                                            \
  int orc_horner_##SUF(int ns, double beta, double tol, T *coef /*[19*ns]*/, int *nc) {        \
    std::vector<T> c;                                                                          \
    int err = orc::horner_table<T>(ns, beta, (T)tol, c, *nc);                                  \
    if (!err) std::copy(c.begin(), c.end(), coef);                                             \
    return err;                                                                                \
  }                                                                                            \
  void orc_polyfit_pswf_##SUF(int ns, double beta, int panel, int n, T *out) {                 \
    orc::Pswf0 psi(beta);                                                                      \
    const T shift = T(2 * panel + 1 - ns);                                                     \
    auto f        = [&](T x) -> T { return (T)psi((double)((x + shift) / (T)ns)); };           \
    std::vector<T> c = orc::poly_fit<T>(f, n);                                                 \
    std::copy(c.begin(), c.end(), out);                                                        \
  }                                                                                            \
  void orc_eval_stencil_##SUF(T x1, int ns, int nc, const T *coef, T *ker) {                   \
    orc::eval_stencil<T>(x1, ns, nc, coef, ker);                                               \
  }                                                                                            \
  T orc_eval_kernel_##SUF(T x, int ns, int nc, const T *coef) {                                \
    return orc::eval_kernel<T>(x, ns, nc, coef);                                               \
  }                                                                                            \
  void orc_fseries_##SUF(i64 nf, int ns, int nc, const T *coef, T *out) {                      \
    orc::fseries<T>(nf, ns, nc, coef, out);                                                    \
  }                                                                                            \
  void orc_fold_rescale_##SUF(i64 n, const T *x, i64 N, T *out) {                              \
    for (i64 i = 0; i < n; ++i) out[i] = orc::fold_rescale<T>(x[i], N);                        \
  }                                                                                            \
  void orc_bin_sort_##SUF(i64 M, const T *x, const T *y, const T *z, i64 N1, i64 N2, i64 N3,  \
                          i64 *perm, i64 *bins) {                                              \
    orc::bin_sort<T>(M, x, y, z, N1, N2, N3, perm, bins);                                      \
  }                                                                                            \
  void orc_spread_##SUF(int dim, const i64 *nf, i64 M, const T *x, const T *y, const T *z,    \
                        const T *c, const i64 *perm, int ns, int nc, const T *coef, T *fw,     \
                        int nthr) {                                                            \
    orc::KerTab<T> K{ns, nc, coef};                                                            \
    orc::spread_sorted<T>(dim, nf[0], dim > 1 ? nf[1] : 1, dim > 2 ? nf[2] : 1, M, x, y, z, c, \
                          perm, K, fw, nthr, dim == 1 ? 10000 : 100000);                       \
  }                                                                                            \
  void orc_interp_##SUF(int dim, const i64 *nf, i64 M, const T *x, const T *y, const T *z,    \
                        T *c, const i64 *perm, int ns, int nc, const T *coef, const T *fw,     \
                        int nthr) {                                                            \
    orc::KerTab<T> K{ns, nc, coef};                                                            \
    orc::interp_sorted<T>(dim, nf[0], dim > 1 ? nf[1] : 1, dim > 2 ? nf[2] : 1, M, x, y, z, c, \
                          perm, K, fw, nthr);                                                  \
  }                                                                                            \
  void orc_deconvolve_##SUF(int dir, int dim, const i64 *ms, const i64 *nf, int modeord,      \
                            const T *p1, const T *p2, const T *p3, T *fk, T *fw) {             \
    orc::deconvolve<T>(dir, dim, ms, nf, modeord, p1, p2, p3, fk, fw);                         \
  }                                                                                            \
  void orc_fft_##SUF(int dim, const i64 *nf, int sign, T *data, int nthr) {                    \
    orc::fft_nd<T>(dim, nf, sign, data, nthr);                                                 \
  }                                                                                            \
  int orc_makeplan_##SUF(int type, int dim, const i64 *nmodes, int iflag, int ntr, double tol, \
                         double sigma, int modeord, int spread_only, int allow_small,          \
                         int nthr, void **plan) {                                              \
    orc::Plan<T> *p = nullptr;                                                                 \
    int err = orc::makeplan<T>(type, dim, nmodes, iflag, ntr, tol, sigma, modeord,             \
                               spread_only, allow_small, nthr, &p);                            \
    *plan = p;                                                                                 \
    return err;                                                                                \
  }                                                                                            \
  int orc_setpts_##SUF(void *plan, i64 M, const T *x, const T *y, const T *z, i64 nk,         \
                       const T *s, const T *t, const T *u, int allow_small, int force_sort) { \
    return orc::setpts<T>(*(orc::Plan<T> *)plan, M, x, y, z, nk, s, t, u, allow_small,         \
                          force_sort);                                                         \
  }                                                                                            \
  int orc_execute_##SUF(void *plan, T *c, T *fk, int adjoint) {                                \
    return orc::execute<T>(*(orc::Plan<T> *)plan, c, fk, adjoint);                             \
  }                                                                                            \
  void orc_destroy_##SUF(void *plan) { delete (orc::Plan<T> *)plan; }                          \
  void orc_plan_info_##SUF(void *plan, int *ns, int *nc, double *beta, i64 *nf, double *tol) { \
    auto *p = (orc::Plan<T> *)plan;                                                            \
    *ns = p->ns; *nc = p->nc; *beta = p->beta; *tol = p->tol;                                  \
    nf[0] = p->nf[0]; nf[1] = p->nf[1]; nf[2] = p->nf[2];                                      \
  }                                                                                            \
  void orc_plan_tables_##SUF(void *plan, T *coef, T *ph1, T *ph2, T *ph3) {                    \
    auto *p = (orc::Plan<T> *)plan;                                                            \
    if (coef) std::copy(p->coef.begin(), p->coef.end(), coef);                                 \
    T *ph[3] = {ph1, ph2, ph3};                                                                \
    for (int d = 0; d < 3; ++d)                                                                \
      if (ph[d]) std::copy(p->phihat[d].begin(), p->phihat[d].end(), ph[d]);                   \
  }                                                                                            \
  void orc_plan_perm_##SUF(void *plan, i64 *perm) {                                            \
    auto *p = (orc::Plan<T> *)plan;                                                            \
    std::copy(p->perm.begin(), p->perm.end(), perm);                                           \
  }

ORC_API(f32, float)
ORC_API(f64, double)

}  // extern "C"
