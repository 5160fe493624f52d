// Host-side plan-time mathematics (see planmath.hpp for the behavioural spec citations).
#include "planmath.hpp"

#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>

#include "errors.hpp"

namespace b200 {

// ------------------------------------------------------------------ width and shape
static double aliasing_prefactor(int dim, int type) {
  // tol ~ prefactor * exp(-(ns-1) pi sqrt(1-1/sigma)); empirical constants of the reference
  // (src/common/kernel.cpp:60-77): 0.18 in 1D, x1.4 per extra dimension, x1.4 for type 3.
  double f = 0.18;
  for (int d = 1; d < dim; ++d) f *= 1.4;
  if (type == 3) f *= 1.4;
  return f;
}

int choose_kernel(double tol, int dim, int type, double sigma, bool is_float, bool allow_small,
                  int &ns, double &beta, double &tol_used) {
  if (!(sigma > 1.0)) return ERR_UPSAMPFAC_TOO_SMALL;
  // tolerance is held in the plan's working precision (makeplan.hpp:154-162)
  const double eps = is_float ? (double)std::numeric_limits<float>::epsilon()
                              : std::numeric_limits<double>::epsilon();
  double t = is_float ? (double)(float)tol : tol;
  if (t < eps) {
    if (!allow_small) return ERR_EPS_TOO_SMALL;
    t = eps;
  }
  const int cap = is_float ? kMaxNsF32 : kMaxNsF64;
  const double rate = kPi * std::sqrt(1.0 - 1.0 / sigma);
  const int ideal   = (int)std::ceil(std::log(aliasing_prefactor(dim, type) / t) / rate + 1.0);
  if (ideal > cap && !allow_small) return ERR_EPS_TOO_SMALL;
  int w = std::min(std::max(ideal, 2), cap);
  if (is_float && sigma < 1.4) w = std::min(w, 8);  // single-precision cancellation guard
  ns       = w;
  beta     = kPi * w * (1.0 - 0.5 / sigma) - 0.05;  // prolate bandwidth, cutoff minus 0.05
  tol_used = t;
  return 0;
}

// ------------------------------------------------------------------ prolate window
// psi_0^c expanded in normalised even Legendre functions Pbar_k = sqrt(k+1/2) P_k, k=0,2,4,..:
// (A - chi I) b = 0 with A symmetric tridiagonal
//   A[k,k]   = k(k+1) + c^2 (2k(k+1)-1) / ((2k+3)(2k-1))
//   A[k,k+2] = c^2 (k+2)(k+1) / ((2k+3) sqrt((2k+1)(2k+5)))
// psi_0 belongs to the smallest eigenvalue chi_0.
Prolate0::Prolate0(double c) {
  const int K = 40 + (int)std::ceil(std::abs(c));
  std::vector<double> d(K), e(K, 0.0);
  const double c2 = c * c;
  for (int j = 0; j < K; ++j) {
    const double k = 2.0 * j;
    d[j] = k * (k + 1.0) + c2 * (2.0 * k * (k + 1.0) - 1.0) / ((2.0 * k + 3.0) * (2.0 * k - 1.0));
    e[j] = c2 * (k + 2.0) * (k + 1.0) /
           ((2.0 * k + 3.0) * std::sqrt((2.0 * k + 1.0) * (2.0 * k + 5.0)));  // couples j, j+1
  }
  // Gershgorin bracket, then bisection on the Sturm count for the smallest eigenvalue.
  double lo = std::numeric_limits<double>::max(), hi = -lo;
  for (int j = 0; j < K; ++j) {
    const double r = (j ? std::abs(e[j - 1]) : 0.0) + (j + 1 < K ? std::abs(e[j]) : 0.0);
    lo = std::min(lo, d[j] - r);
    hi = std::max(hi, d[j] + r);
  }
  auto count_below = [&](double lam) {
    int cnt  = 0;
    double q = d[0] - lam;
    if (q < 0) ++cnt;
    for (int j = 1; j < K; ++j) {
      if (q == 0.0) q = 1e-300;
      q = d[j] - lam - e[j - 1] * e[j - 1] / q;
      if (q < 0) ++cnt;
    }
    return cnt;
  };
  for (int it = 0; it < 300 && hi - lo > 2e-16 * std::max(1.0, std::abs(lo) + std::abs(hi)); ++it) {
    const double mid = 0.5 * (lo + hi);
    if (count_below(mid) >= 1) hi = mid;
    else lo = mid;
  }
  const double chi   = 0.5 * (lo + hi);
  const double shift = chi - 1e-9 * std::max(1.0, std::abs(chi));
  // inverse iteration with an LDL^T factorisation of A - shift I (positive definite)
  std::vector<double> piv(K), l(K, 0.0), v(K, 1.0);
  piv[0] = d[0] - shift;
  for (int j = 1; j < K; ++j) {
    l[j]   = e[j - 1] / piv[j - 1];
    piv[j] = d[j] - shift - l[j] * e[j - 1];
  }
  for (int j = 0; j < K; ++j)
    if (!(piv[j] > 0.0) || !std::isfinite(piv[j])) return;  // ok_ stays false
  for (int it = 0; it < 5; ++it) {
    for (int j = 1; j < K; ++j) v[j] -= l[j] * v[j - 1];
    for (int j = 0; j < K; ++j) v[j] /= piv[j];
    for (int j = K - 2; j >= 0; --j) v[j] -= l[j + 1] * v[j + 1];
    double nrm = 0;
    for (double t : v) nrm += t * t;
    nrm = 1.0 / std::sqrt(nrm);
    for (double &t : v) t *= nrm;
  }
  // to plain Legendre coefficients; drop the negligible tail
  double big = 0;
  for (int j = 0; j < K; ++j) {
    v[j] *= std::sqrt(2.0 * j + 0.5);
    big = std::max(big, std::abs(v[j]));
  }
  int last = 0;
  for (int j = 0; j < K; ++j)
    if (std::abs(v[j]) > 1e-18 * big) last = j;
  leg_.assign(v.begin(), v.begin() + last + 1);
  const double at0 = series(0.0);
  if (at0 == 0.0 || !std::isfinite(at0)) return;
  scale_ = 1.0 / at0;
  ok_    = true;
}

double Prolate0::series(double x) const {
  // sum_j leg_[j] P_{2j}(x) by the three-term recurrence on all degrees
  double pm = 1.0, p = x, s = leg_[0];
  const int top = 2 * ((int)leg_.size() - 1);
  for (int n = 1; n < top; ++n) {
    const double pn = ((2.0 * n + 1.0) * x * p - n * pm) / (n + 1.0);  // P_{n+1}
    pm = p;
    p  = pn;
    if (((n + 1) & 1) == 0) s += leg_[(n + 1) / 2] * p;
  }
  return s;
}

double Prolate0::operator()(double x) const {
  if (std::abs(x) > 1.0) return 0.0;
  return series(x) * scale_;
}

// ------------------------------------------------------------------ polynomial table
// Interpolate g on the n first-kind Chebyshev nodes and return monomial coefficients,
// highest degree first.  Arithmetic in T (the reference fits in the plan's precision).
template<class T, class G> static std::vector<T> cheb_fit_monomial(G &&g, int n) {
  std::vector<T> node(n), dd(n);
  for (int k = 0; k < n; ++k) {
    node[k] = (T)std::cos(((double)T(2 * k + 1) * kPi) / (double)(T(2) * T(n)));
    dd[k]   = (T)g(node[k]);
  }
  for (int lvl = 1; lvl < n; ++lvl)  // Newton divided differences, in place
    for (int i = n - 1; i >= lvl; --i) dd[i] = (dd[i] - dd[i - 1]) / (node[i] - node[i - lvl]);
  std::vector<T> mono(n, T(0)), prod(1, T(1)), next;
  mono[0] += dd[0];
  for (int lvl = 1; lvl < n; ++lvl) {  // prod <- prod * (t - node[lvl-1]), low -> high
    const T root = node[lvl - 1];
    next.assign(prod.size() + 1, T(0));
    for (size_t i = 0; i < prod.size(); ++i) {
      next[i] += -root * prod[i];
      next[i + 1] += prod[i];
    }
    prod.swap(next);
    for (size_t i = 0; i < prod.size(); ++i) mono[i] += dd[lvl] * prod[i];
  }
  std::reverse(mono.begin(), mono.end());
  return mono;
}

template<class T>
int build_horner_table(int ns, double beta, T tol, std::vector<T> &coef, int &nc) {
  Prolate0 psi(beta);
  if (!psi.ok()) return ERR_PSWF_SETUP;
  const int nfit = std::min(kMaxNc, ns + 3);
  std::vector<T> full((size_t)nfit * ns);
  int need_max = 4;
  for (int j = 0; j < ns; ++j) {
    const T centre = T(2 * j + 1 - ns);  // panel j is [-1+2j/ns, -1+2(j+1)/ns] in window units
    auto panel = [&](T u) -> T { return (T)psi((double)((u + centre) / (T)ns)); };
    const std::vector<T> cj = cheb_fit_monomial<T>(panel, nfit);
    for (int k = 0; k < nfit; ++k) full[(size_t)k * ns + j] = cj[k];
    const T small = tol * T(0.05);  // leading coefficients below this are dropped
    for (int k = 0; k < nfit; ++k)
      if (std::abs(cj[k]) >= small) {
        need_max = std::max(need_max, nfit - k);
        break;
      }
  }
  nc = std::max(need_max, std::max(4, ns - 4));
  coef.assign(full.begin() + (size_t)(nfit - nc) * ns, full.end());
  return 0;
}
template int build_horner_table<float>(int, double, float, std::vector<float> &, int &);
template int build_horner_table<double>(int, double, double, std::vector<double> &, int &);

template<class T> double eval_table(double x, int ns, int nc, const T *coef) {
  const double half = 0.5 * ns;
  if (!(std::abs(x) <= half)) return 0.0;
  int j = (int)std::ceil(x + half) - 1;  // panel with x in (-half+j, -half+j+1]
  j     = std::min(std::max(j, 0), ns - 1);
  const double z = 2.0 * (x - j) + (ns - 1);
  double r = 0.0;
  for (int k = 0; k < nc; ++k) r = r * z + (double)coef[(size_t)k * ns + j];
  return r;
}
template double eval_table<float>(double, int, int, const float *);
template double eval_table<double>(double, int, int, const double *);

// ------------------------------------------------------------------ grid sizes
static bool is_smooth235(int64_t n) {
  for (int p : {2, 3, 5})
    while (n % p == 0) n /= p;
  return n == 1;
}
int64_t next_smooth_even(int64_t n) {
  int64_t h = std::max<int64_t>((n + 1) / 2, 1);
  while (!is_smooth235(h)) ++h;
  return 2 * h;
}
int64_t fine_grid_size(double sigma, int64_t modes, int ns) {
  int64_t nf = (int64_t)std::ceil(sigma * (double)modes);
  nf         = std::max<int64_t>(nf, 2 * ns);
  if (nf >= (int64_t)1e12) return -1;
  return next_smooth_even(nf);
}

// ------------------------------------------------------------------ quadrature
void gauss_legendre(int n, double *x, double *w) {
  // Newton on P_n from Chebyshev guesses; weights 2 / ((1-x^2) P_n'(x)^2).
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double t = std::cos(kPi * (i + 0.75) / (n + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = t;
      for (int k = 1; k < n; ++k) {
        const double p2 = ((2.0 * k + 1.0) * t * p1 - k * p0) / (k + 1.0);
        p0 = p1;
        p1 = p2;
      }
      dp = n * (t * p1 - p0) / (t * t - 1.0);
      const double dt = p1 / dp;
      t -= dt;
      if (std::abs(dt) < 1e-15) {
        // refresh derivative at the converged node
        p0 = 1.0, p1 = t;
        for (int k = 1; k < n; ++k) {
          const double p2 = ((2.0 * k + 1.0) * t * p1 - k * p0) / (k + 1.0);
          p0 = p1;
          p1 = p2;
        }
        dp = n * (t * p1 - p0) / (t * t - 1.0);
        break;
      }
    }
    const double wt = 2.0 / ((1.0 - t * t) * dp * dp);
    x[i]         = -t;
    x[n - 1 - i] = t;
    w[i] = w[n - 1 - i] = wt;
  }
  if (n & 1) x[n / 2] = 0.0;
}

// window value at grid-unit argument x, all arithmetic in T with fused Horner steps
template<class T> static T eval_table_native(T x, int ns, int nc, const T *coef) {
  const T half = ns / T(2.0);
  for (int j = 0; j < ns; ++j)
    if (x > -half + j && x <= -half + j + 1) {
      const T z = std::fma(T(2.0), x - T(j), T(ns - 1));
      T r       = T(0);
      for (int k = 0; k < nc; ++k) r = std::fma(r, z, coef[(size_t)k * ns + j]);
      return r;
    }
  return T(0);
}

template<class T>
void fseries_wound(int64_t nf, int ns, int nc, const T *coef, std::vector<T> &out) {
  const T half = ns / 2.0;
  const int q  = (int)(2 + 3.0 * half);
  std::vector<double> x(2 * q), w(2 * q);
  gauss_legendre(2 * q, x.data(), w.data());
  std::vector<T> f(q);
  std::vector<std::complex<T>> rot(q), cur(q, std::complex<T>(1, 0));
  for (int n = 0; n < q; ++n) {
    const double zn = x[n] * half;  // node in (-ns/2, 0)
    f[n]            = half * (T)w[n] * eval_table_native<T>(T(zn), ns, nc, coef);
    const std::complex<double> a = -std::exp(2 * kPi * std::complex<double>(0, 1) * zn / double(nf));
    rot[n] = std::complex<T>((T)a.real(), (T)a.imag());
  }
  out.resize(nf / 2 + 1);
  for (int64_t k = 0; k <= nf / 2; ++k) {
    T s = 0.0;
    for (int n = 0; n < q; ++n) {
      s += f[n] * 2 * std::real(cur[n]);
      cur[n] *= rot[n];
    }
    out[k] = s;
  }
}
template void fseries_wound<float>(int64_t, int, int, const float *, std::vector<float> &);
template void fseries_wound<double>(int64_t, int, int, const double *, std::vector<double> &);

// ------------------------------------------------------------------ sigma feasibility
static double sigma_reaching(double tol, int dim, int type, int ns) {
  const double pre = aliasing_prefactor(dim, type);
  if (tol <= 0) return 2.5;
  if (tol >= pre) return 1.01;
  const double u = std::log(pre / tol) / ((ns - 1.0) * kPi);
  if (u >= 1.0) return 2.5;
  return std::min(1.0 / (1.0 - u * u), 2.5);
}
double least_sigma(double tol, int dim, int ns, double eps_mach, double gridlen) {
  // src/common/kernel.cpp:172-201: analytic inversion of the aliasing law, plus an
  // empirical 1/r polynomial near the rounding floor eps_round = 0.48 eps N.
  const double r = tol / (0.48 * eps_mach * gridlen);
  if (r <= 0.5) return 2.0;
  const double pure = sigma_reaching(tol, dim, 1, ns);
  if (r >= 10.0) return std::min(pure, 2.0);
  const bool wide = ns > 8;
  const double a2 = wide ? 0.014 : 0.555, a1 = wide ? 0.291 : -0.290, a0 = wide ? -0.043 : 0.071;
  const double corr = (a2 / r + a1) / r + a0;
  return std::min(pure + std::max(corr, 0.0), 2.0);
}

bool sigma_feasible(double sigma, double tol, int dim, int type, bool is_float, double maxN) {
  if (!(sigma > 1.0)) return false;
  const double eps = is_float ? (double)std::numeric_limits<float>::epsilon()
                              : std::numeric_limits<double>::epsilon();
  const int cap    = is_float ? kMaxNsF32 : kMaxNsF64;
  // width the aliasing law asks for (no clamping) against the width the plan would use
  const double rate = kPi * std::sqrt(1.0 - 1.0 / sigma);
  const int ideal   = (int)std::ceil(std::log(aliasing_prefactor(dim, type) / tol) / rate + 1.0);
  int w = std::min(std::max(ideal, 2), cap);
  if (is_float && sigma < 1.4) w = std::min(w, 8);
  if (w < ideal) return false;
  if (type == 3) return true;
  const int64_t nf = fine_grid_size(sigma, (int64_t)maxN, w);
  if (nf < 0) return false;
  return least_sigma(tol, dim, w, eps, (double)nf) <= sigma;
}

double smallest_feasible_sigma(double tol, int dim, int type, bool is_float, double maxN) {
  constexpr double lo0 = 1.15, hi0 = 2.5;
  if (sigma_feasible(lo0, tol, dim, type, is_float, maxN)) return lo0;
  if (!sigma_feasible(hi0, tol, dim, type, is_float, maxN)) return hi0;
  double lo = lo0, hi = hi0;
  for (int i = 0; i < 40; ++i) {
    const double mid = 0.5 * (lo + hi);
    (sigma_feasible(mid, tol, dim, type, is_float, maxN) ? hi : lo) = mid;
  }
  return hi;
}

namespace {
int width_at(double tol, int dim, int type, double sigma, bool is_float) {
  int ns = 0;
  double beta = 0, tu = 0;
  if (choose_kernel(tol, dim, type, sigma, is_float, true, ns, beta, tu)) return 16;
  return ns;
}
// Cost model of one execute on a B200, milliseconds (measured: profiles/r2z_bench_*.json).
// spread / interp of M points with a width-ns kernel:
double spread_ms(int dim, bool is_float, int ns, double M) {
  double stencil = 1.0;
  for (int d = 0; d < dim; ++d) stencil *= ns;
  double base, per_cell;  // ms per point: base + per_cell * ns^dim
  if (dim == 3) {
    const bool sweep = is_float && ns <= 7;            // k_sweep3: 8.05 ms per 1e8 points at ns = 7
    base = 1.0e-8, per_cell = sweep ? 2.1e-10 : 9.5e-10;  // generic kernels: 35 ms at ns = 7
    if (!is_float) per_cell *= 2.0;
  } else if (dim == 2) {
    base = is_float ? 1.2e-8 : 2.0e-8;                  // k_sweep2: 2.42 ms (f32, ns 6), 8.1 ms per 1e8 (f64, ns 10)
    per_cell = is_float ? 3.4e-10 : 6.1e-10;
  } else {
    base = 2.0e-8, per_cell = is_float ? 1.4e-9 : 2.7e-9;  // generic 1D: 4.7 ms per 1e8 (f64, ns 10)
  }
  return M * (base + per_cell * stencil);
}
// cuFFT + the zero / deconvolve passes over a fine grid: 1.08 + 0.27 ms on 512^3 (f32)
double grid_ms(double cells, bool is_float) {
  const double w = is_float ? 1.0 : 2.0;
  return cells * std::log2(std::max(cells, 2.0)) * 3.0e-10 * w + cells * 2.0e-9 * w;
}
double cost_ms(int dim, bool is_float, const int64_t *modes, double sigma, int ns, double M) {
  double cells = 1.0;
  for (int d = 0; d < dim; ++d) cells *= (double)fine_grid_size(sigma, modes[d], ns);
  return spread_ms(dim, is_float, ns, M) + grid_ms(cells, is_float);
}
double working_tol(double tol, bool is_float) {
  const double eps = is_float ? (double)std::numeric_limits<float>::epsilon()
                              : std::numeric_limits<double>::epsilon();
  const double t = is_float ? (double)(float)tol : tol;
  return t < eps ? eps : t;
}
// The minimiser over the candidate set, with this library's two house rules: candidates stop at
// sigma = 2 (the reference searches up to 2.5; above 2 the fine grid grows faster than the kernel
// narrows on this device, and 2 is where every kernel family is tuned), and sigma = 2 is kept
// unless another candidate is clearly (by the factor `margin`) cheaper.  cost(sigma, ns) in milliseconds.
template<class Cost>
double pick_sigma(double t, int dim, int type, bool is_float, double maxN, Cost &&cost,
                  double margin, double *cost_out = nullptr) {
  constexpr double smax = 2.0;
  double sig[32];
  int wid[32];
  const int n = sigma_candidates(t, dim, type, is_float, maxN, smax, sig, wid, 32);
  // tolerance out of reach, or nothing below the cap: the default
  if (n == 0 || sig[0] >= smax || !sigma_feasible(sig[0], t, dim, type, is_float, maxN)) {
    if (cost_out) *cost_out = cost(smax, width_at(t, dim, type, smax, is_float));
    return smax;
  }
  double best = smax, best_cost = std::numeric_limits<double>::infinity();
  for (int i = 0; i < n; ++i) {
    const double c = cost(sig[i], wid[i]);
    if (c < best_cost) best = sig[i], best_cost = c;
  }
  if (sigma_feasible(2.0, t, dim, type, is_float, maxN)) {
    const double c2 = cost(2.0, width_at(t, dim, type, 2.0, is_float));
    if (c2 <= margin * best_cost) best = 2.0, best_cost = c2;
  }
  if (cost_out) *cost_out = best_cost;
  return best;
}
}  // namespace

int sigma_candidates(double tol, int dim, int type, bool is_float, double maxN, double smax,
                     double *sigma_out, int *ns_out, int cap) {
  int n = 0;
  auto push = [&](double s) {
    if (n < cap) sigma_out[n] = s, ns_out[n] = width_at(tol, dim, type, s, is_float), ++n;
  };
  const double smin = smallest_feasible_sigma(tol, dim, type, is_float, maxN);
  push(smin);
  if (!sigma_feasible(smin, tol, dim, type, is_float, maxN) || smin >= smax) return n;
  const int ns_lo = width_at(tol, dim, type, smax, is_float);
  for (int w = width_at(tol, dim, type, smin, is_float) - 1; w >= ns_lo; --w) {
    // sigma_reaching sits exactly on the edge of the width law's ceil(); one part in 1e12 above
    // it the width is w whatever the rounding of the log / sqrt did
    const double edge = sigma_reaching(tol, dim, type, w) * (1.0 + 1e-12);
    const double s    = std::min(std::max(edge, smin), smax);
    if (sigma_feasible(s, tol, dim, type, is_float, maxN)) push(s);
  }
  return n;
}

double choose_sigma(double tol, int dim, int type, bool is_float, const int64_t *modes,
                    double npoints, double *cost_out) {
  const double t = working_tol(tol, is_float);
  double maxN = 1.0;
  for (int d = 0; d < dim; ++d) maxN = std::max(maxN, (double)modes[d]);
  return pick_sigma(
      t, dim, type, is_float, maxN,
      [&](double s, int ns) { return cost_ms(dim, is_float, modes, s, ns, npoints); }, 1.1,
      cost_out);
}

double choose_sigma_type3(double tol, int dim, bool is_float, double nsources, double ntargets,
                          const double *X, const double *S) {
  const double t = working_tol(tol, is_float);
  auto cost = [&](double s3, int ns3) {
    int64_t nfd[3] = {1, 1, 1};
    double cells   = 1.0;
    for (int d = 0; d < dim; ++d) {
      double h, gam;
      type3_grid(s3, X[d], S[d], ns3, nfd[d], h, gam);
      cells *= (double)nfd[d];
    }
    double inner = 0.0;  // the inner type 2 re-optimises its own sigma from nfd and the targets
    choose_sigma(t, dim, 2, is_float, nfd, ntargets, &inner);
    // outer spread + zero fill of the spreading grid (one write pass) + inner transform
    return spread_ms(dim, is_float, ns3, nsources) + cells * 1.3e-9 * (is_float ? 1.0 : 2.0) +
           inner;
  };
  // A wider margin than for types 1 / 2: the model knows cuFFT by cell count only, and at C5
  // (M = N = 1e7, tol 1e-6) the candidate it scored 11 % cheaper (sigma3 = 1.937: grids 432^3 /
  // 864^3 instead of 450^3 / 900^3) measured 11 % slower; the picks that matter (few points on a
  // wide frequency box, FFT-dominated) are cheaper by integer factors.
  return pick_sigma(t, dim, 3, is_float, 1.0, cost, 1.25);
}

// ------------------------------------------------------------------ type 3
void type3_grid(double sigma, double X, double S, int ns, int64_t &nf, double &h, double &gam) {
  double Xs = X, Ss = S;  // enforce X*S >= 1, also when either is zero
  if (Xs == 0.0) {
    if (Ss == 0.0) Xs = Ss = 1.0;
    else Xs = 1.0 / Ss;
  } else {
    Ss = std::max(Ss, 1.0 / Xs);
  }
  double want = 2.0 * sigma * Ss * Xs / kPi + (ns + 1);
  if (!std::isfinite(want)) want = 0.0;
  nf = std::max<int64_t>((int64_t)want, 2 * ns);
  if (nf < (int64_t)1e12) nf = next_smooth_even(nf);
  h   = 2.0 * kPi / (double)nf;
  gam = (double)nf / (2.0 * sigma * Ss);
}

template<class T>
void selfft_params(int ns, double beta, int nc, const T *coef, double &grid_scale,
                   double &prefac) {
  // integral of the tabulated window: per panel, int_{-1}^{1} sum_k c_k z^(nc-1-k) dz / 2;
  // odd powers vanish, even power p contributes c/(p+1).
  double total = 0.0;
  for (int j = 0; j < ns; ++j)
    for (int k = nc - 1; k >= 0; k -= 2) total += (double)coef[(size_t)k * ns + j] / (nc - k);
  prefac     = total;
  grid_scale = 0.25 * ns * ns / beta;
}
template void selfft_params<float>(int, double, int, const float *, double &, double &);
template void selfft_params<double>(int, double, int, const double *, double &, double &);

}  // namespace b200
