// ORACLE - TEST INFRASTRUCTURE ONLY.
//
// Read-only introspection of a plan of the REFERENCE library built into
// oracle/_ref/libfinufft_ref.so (oracle/build.py::build_ref_library): the sort permutation
// setpts produced (include/finufft/spreadinterp.hpp:120-196, spread.hpp:459-584), the kernel
// parameters and tables makeplan chose, so that tests can compare the restatement and the GPU
// library with the reference's own state, bit for bit.  Nothing is computed here, except that
// the reference's own sigma search (include/finufft/heuristics.hpp) is called for the tests of
// the automatic upsampfac: the candidate (sigma, ns) pairs its minimiser scores and its picks.
#include <array>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define private public  // the plan's state is private; layout is unaffected by access
#include <finufft/plan.hpp>
#undef private
#include <finufft/heuristics.hpp>

namespace {
template<class T> int64_t perm(void *plan, int64_t *out, int *did_sort) {
  auto *p = static_cast<FINUFFT_PLAN_T<T> *>(plan);
  const auto &s = p->m.sortIndices;
  if (did_sort) *did_sort = p->m.didSort ? 1 : 0;
  if (out) std::memcpy(out, s.data(), sizeof(int64_t) * s.size());
  return (int64_t)s.size();
}
template<class T>
void info(void *plan, int *ns, int *nc, double *beta, double *sigma, int64_t nf[3], double *tol) {
  auto *p = static_cast<FINUFFT_PLAN_T<T> *>(plan);
  *ns     = p->m.spopts.nspread;
  *nc     = p->m.nc;
  *beta   = p->m.spopts.beta;
  *sigma  = p->m.spopts.upsampfac;
  *tol    = (double)p->m.tol;
  for (int d = 0; d < 3; ++d) nf[d] = p->m.nfdim[d];
}
template<class T> int64_t phihat(void *plan, int d, T *out) {
  auto *p = static_cast<FINUFFT_PLAN_T<T> *>(plan);
  const auto &v = p->m.phiHat[d];
  if (out) std::memcpy(out, v.data(), sizeof(T) * v.size());
  return (int64_t)v.size();
}
template<class T>
int trace(double tol, int dim, int type, double maxN, double *sig, int *ns, int cap) {
  int n = 0;
  // heuristics.hpp:82-107: every call of the cost functor is one candidate, in order
  finufft::heuristics::minimize<T>(tol, dim, type, maxN, [&](double s, int w) {
    if (n < cap) sig[n] = s, ns[n] = w;
    ++n;
    return 1.0;
  });
  return n;
}
}  // namespace

extern "C" {
__attribute__((visibility("default"))) int ref_minimize_trace(double tol, int dim, int type,
                                                              int is_float, double maxN,
                                                              double *sig, int *ns, int cap) {
  return is_float ? trace<float>(tol, dim, type, maxN, sig, ns, cap)
                  : trace<double>(tol, dim, type, maxN, sig, ns, cap);
}
__attribute__((visibility("default"))) double ref_best_type3(double tol, int dim, int nthreads,
                                                             double nj, const double *X,
                                                             const double *S, double nk,
                                                             int is_float) {
  return is_float ? finufft::heuristics::best_type3<float>(tol, dim, nthreads, nj, X, S, nk)
                  : finufft::heuristics::best_type3<double>(tol, dim, nthreads, nj, X, S, nk);
}
__attribute__((visibility("default"))) int64_t ref_plan_perm_f32(void *p, int64_t *o, int *d) {
  return perm<float>(p, o, d);
}
__attribute__((visibility("default"))) int64_t ref_plan_perm_f64(void *p, int64_t *o, int *d) {
  return perm<double>(p, o, d);
}
__attribute__((visibility("default"))) void ref_plan_info_f32(void *p, int *ns, int *nc,
                                                              double *beta, double *sigma,
                                                              int64_t nf[3], double *tol) {
  info<float>(p, ns, nc, beta, sigma, nf, tol);
}
__attribute__((visibility("default"))) void ref_plan_info_f64(void *p, int *ns, int *nc,
                                                              double *beta, double *sigma,
                                                              int64_t nf[3], double *tol) {
  info<double>(p, ns, nc, beta, sigma, nf, tol);
}
__attribute__((visibility("default"))) int64_t ref_plan_phihat_f32(void *p, int d, float *o) {
  return phihat<float>(p, d, o);
}
__attribute__((visibility("default"))) int64_t ref_plan_phihat_f64(void *p, int d, double *o) {
  return phihat<double>(p, d, o);
}
}
