// Host-side plan-time mathematics for the B200 NUFFT engine: kernel width/shape selection,
// the order-zero prolate (PSWF) window, its piecewise-polynomial table, fine-grid sizing and
// Gauss-Legendre quadrature for the window's Fourier series.
//
// Behavioural spec (what must agree with the reference CPU library, finufft @ 9810998d):
//   width/shape rule      include/finufft/makeplan.hpp:112-201, src/common/kernel.cpp:60-146
//   prolate window        src/common/kernel.cpp:18-58, src/common/pswf.cpp
//   polynomial table      include/finufft/makeplan.hpp:204-315, include/finufft_common/kernel.h:19-67
//   fine grid size        include/finufft/makeplan.hpp:18-37, src/common/utils.cpp:84-122
//   Fourier series nodes  include/finufft/makeplan.hpp:72-90, src/common/utils.cpp:18-79
//   sigma feasibility     include/finufft/setpts.hpp:29-53, src/common/kernel.cpp:151-201
//   type-3 grid           include/finufft_common/kernel.h:165-181, :114-123
#pragma once
#include <cstdint>
#include <vector>

namespace b200 {

constexpr int kMaxNc     = 19;  // most polynomial coefficients per panel
constexpr int kMaxNsF32  = 12;  // widest kernel, single precision
constexpr int kMaxNsF64  = 16;  // widest kernel, double precision
constexpr int kMaxQuad   = 100; // most positive quadrature nodes
constexpr double kPi     = 3.141592653589793238462643383279502884;
constexpr double kInv2Pi = 0.159154943091895335768883763372514362;

// Width ns and shape beta for (tol, dim, type, sigma).  Returns 0 or a FINUFFT error code
// (7: sigma<=1, 26: tol below machine epsilon / width cap, unless allow_small).
int choose_kernel(double tol, int dim, int type, double sigma, bool is_float, bool allow_small,
                  int &ns, double &beta, double &tol_used);

// Order-zero prolate spheroidal wavefunction psi_0^c on [-1,1], scaled so psi(0)=1,
// zero outside.  Even-Legendre expansion; smallest eigenpair by bisection + inverse iteration.
class Prolate0 {
 public:
  explicit Prolate0(double c);
  double operator()(double x) const;
  bool ok() const { return ok_; }

 private:
  std::vector<double> leg_;  // coefficients of P_0, P_2, P_4, ...
  double scale_ = 1.0;
  bool ok_      = false;
  double series(double x) const;
};

// Piecewise polynomial table of the width-ns window: ns panels, `nc` coefficients each,
// coef[k*ns + j] with k=0 the highest degree of panel j.  Fit arithmetic is done in T.
template<class T>
int build_horner_table(int ns, double beta, T tol, std::vector<T> &coef, int &nc);

// Smallest even 2,3,5-smooth integer >= n.
int64_t next_smooth_even(int64_t n);

// Fine-grid length for a type-1/2 dimension with `modes` modes. Returns -1 if it exceeds 1e12.
int64_t fine_grid_size(double sigma, int64_t modes, int ns);

// n-point Gauss-Legendre rule on [-1,1], nodes ascending.
void gauss_legendre(int n, double *x, double *w);

// Window Fourier series phihat[k], k = 0..nf/2, computed the way the reference CPU library
// does (include/finufft/makeplan.hpp:72-105, single chunk): node weights and the phase
// rotators a_n = -exp(2 pi i z_n / nf) are held in the plan's precision T and wound by repeated
// complex multiplication, phihat[k] = sum_n 2 f_n Re(a_n^k).  In single precision this drifts
// by ~k*eps; it is reproduced deliberately so the deconvolution factors equal the reference's.
template<class T>
void fseries_wound(int64_t nf, int ns, int nc, const T *coef, std::vector<T> &out);

// Evaluate the table at grid-unit argument x in [-ns/2, ns/2] (double arithmetic).
template<class T> double eval_table(double x, int ns, int nc, const T *coef);

// Least sigma that can reach tol on a grid of this length (check_sigma rule).
double least_sigma(double tol, int dim, int ns, double eps_mach, double gridlen);

// Would the plan pipeline accept this sigma at this tolerance (kernel width not clamped, and for
// types 1/2 the rounding-floor rule on the fine grid sigma would build)?  maxN = largest mode
// count over the dimensions.  Reference src/common/kernel.cpp:203-228 (upsampfac_feasible).
bool sigma_feasible(double sigma, double tol, int dim, int type, bool is_float, double maxN);
// Smallest accepted sigma in [1.15, 2.5] (bisection), 2.5 if none.  Reference
// src/common/kernel.cpp:231-257 (analytic_upsampfac).
double smallest_feasible_sigma(double tol, int dim, int type, bool is_float, double maxN);
// Automatic upsampfac for a type-1/2 plan on this device: the candidate set of the reference's
// heuristic (include/finufft/heuristics.hpp:82-128: the smallest feasible sigma, then for every
// narrower kernel width the smallest sigma that reaches it, up to 2.5), scored with a B200 cost
// model (spread / interp time per point by dimension, width, precision and kernel family + FFT
// and grid-pass time per fine-grid cell; constants from profiles/r2z_bench_*.json).
// cost_out (optional): the model's milliseconds for the value returned.
double choose_sigma(double tol, int dim, int type, bool is_float, const int64_t *modes,
                    double npoints, double *cost_out = nullptr);
// The candidate set itself, in the reference's order (heuristics.hpp:82-107 `minimize`): the
// smallest feasible sigma, then for every narrower width down to the width at smax the smallest
// sigma reaching it (clamped to [smin, smax], infeasible ones skipped); with smax = 2.5 these are
// exactly the (sigma, ns) pairs the reference's minimiser scores.  Returns the count.
int sigma_candidates(double tol, int dim, int type, bool is_float, double maxN, double smax,
                     double *sigma_out, int *ns_out, int cap);
// Automatic upsampfac of a type-3 plan (include/finufft/heuristics.hpp:130-150 `best_type3`,
// applied at setpts: include/finufft/setpts.hpp:186-200): candidates as above with type = 3,
// cost = outer spread of the sources at the candidate's width + the inner type-2 transform on the
// grid that sigma builds from the half-widths X (sources) and S (targets), the inner transform at
// ITS best sigma for the number of targets; this device's cost model.
double choose_sigma_type3(double tol, int dim, bool is_float, double nsources, double ntargets,
                          const double *X, const double *S);

// Type-3 fine grid (nf, spacing h, rescale gam) for half-widths X (space) and S (frequency).
void type3_grid(double sigma, double X, double S, int ns, int64_t &nf, double &h, double &gam);

// Analytic self-transform parameters of the tabulated prolate (type-3 deconvolution):
// phihat(xi) = prefac * phi(grid_scale * xi).
template<class T>
void selfft_params(int ns, double beta, int nc, const T *coef, double &grid_scale,
                   double &prefac);

}  // namespace b200
