// ORACLE - TEST INFRASTRUCTURE ONLY.  Implementation of oracle/shim/fftw3.h over
// oracle/fft_standin.hpp.
#include "fftw3.h"

#include <omp.h>

#include "../fft_standin.hpp"

namespace {
struct PlanRec {
  int rank = 0, howmany = 1, sign = -1, nthreads = 1;
  orc::i64 n[3] = {1, 1, 1};  // fastest first
  orc::i64 dist = 0;
};
int g_threads = 1;

template<class C>
PlanRec *make(int rank, const int *n, int howmany, C *in, C *out, int istride, int idist,
              int ostride, int odist, int sign) {
  if (rank < 1 || rank > 3 || in != out || istride != 1 || ostride != 1 || idist != odist)
    return nullptr;
  auto *p   = new PlanRec;
  p->rank   = rank;
  p->howmany = howmany;
  p->sign   = sign;
  p->nthreads = g_threads;
  orc::i64 tot = 1;
  for (int d = 0; d < rank; ++d) {
    p->n[d] = n[rank - 1 - d];  // FFTW lists the slowest dimension first
    tot *= n[d];
  }
  if (howmany > 1 && idist != tot) {
    delete p;
    return nullptr;
  }
  p->dist = tot;
  return p;
}
template<class T> void run(const PlanRec *p, T *data) {
  if (!p) return;
  for (int b = 0; b < p->howmany; ++b)
    orc::fft_nd<T>(p->rank, p->n, p->sign, data + 2 * p->dist * b, p->nthreads);
}
}  // namespace

extern "C" {
#define ORC_FFTW_IMPL(P, C, PLAN, T)                                                             \
  PLAN P##plan_many_dft(int rank, const int *n, int howmany, C *in, const int *, int istride,    \
                        int idist, C *out, const int *, int ostride, int odist, int sign,        \
                        unsigned) {                                                              \
    return reinterpret_cast<PLAN>(make(rank, n, howmany, in, out, istride, idist, ostride,       \
                                       odist, sign));                                            \
  }                                                                                              \
  void P##execute_dft(const PLAN p, C *in, C *) {                                                \
    run<T>(reinterpret_cast<const PlanRec *>(p), reinterpret_cast<T *>(in));                     \
  }                                                                                              \
  void P##destroy_plan(PLAN p) { delete reinterpret_cast<PlanRec *>(p); }                        \
  int P##init_threads(void) { return 1; }                                                        \
  void P##plan_with_nthreads(int nthreads) { g_threads = nthreads < 1 ? 1 : nthreads; }          \
  void P##forget_wisdom(void) {}                                                                 \
  void P##cleanup(void) {}                                                                       \
  void P##cleanup_threads(void) {}
ORC_FFTW_IMPL(fftw_, fftw_complex, fftw_plan, double)
ORC_FFTW_IMPL(fftwf_, fftwf_complex, fftwf_plan, float)
}
