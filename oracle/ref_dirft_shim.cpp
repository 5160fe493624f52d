// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C shim over the REFERENCE's own direct-sum checkers and error norms, compiled by
// oracle/build.py straight from the headers where they lie:
//   /root/reference/test/utils/dirft1d.hpp, dirft2d.hpp, dirft3d.hpp, norms.hpp
// into oracle/_ref/libfinufft_ref_dirft.so (git-ignored, travels to the GPU box).  These are the
// functions every reference accuracy test compares against (test/finufft{1,2,3}d_test.cpp,
// test/tolsweep.cpp, test/cuda/cufinufft*_test.cu), so a transform that passes against them with
// the reference's thresholds is pinned to the reference's own acceptance criterion.  Nothing
// from the reference is copied into this repo; this file only calls it.
//
// The reference functions are serial O(M N) loops.  The shim only adds an OpenMP driver on top:
// type 1 is linear in the points, so each thread runs the reference routine on a contiguous
// chunk of points and the partial mode arrays are added; types 2 and 3 are independent per
// output, so each thread runs the reference routine on a chunk of outputs.
#include <omp.h>

#include <complex>
#include <cstdint>
#include <vector>

#include "dirft1d.hpp"
#include "dirft2d.hpp"
#include "dirft3d.hpp"
#include "norms.hpp"

using i64 = int64_t;
using cd  = std::complex<double>;

namespace {
struct Chunk {
  i64 lo, n;
};
Chunk chunk_of(i64 total, int t, int nt) {
  const i64 base = total / nt, extra = total % nt;
  const i64 lo = t * base + (t < extra ? t : extra);
  return {lo, base + (t < extra ? 1 : 0)};
}
}  // namespace

extern "C" {

// fk[ms*mt*mu] = sum_j c_j exp(i sign k.x_j), CMCL mode order (x fastest)
void ref_dirft1(int dim, i64 M, const double *x, const double *y, const double *z, const cd *c,
                int sign, const i64 *ms, cd *fk, int nthr) {
  const i64 N = ms[0] * (dim > 1 ? ms[1] : 1) * (dim > 2 ? ms[2] : 1);
  if (nthr < 1) nthr = omp_get_max_threads();
  if ((i64)nthr > M) nthr = M > 0 ? (int)M : 1;
  std::vector<std::vector<cd>> part(nthr);
#pragma omp parallel num_threads(nthr)
  {
    const int t = omp_get_thread_num();
    const Chunk ch = chunk_of(M, t, nthr);
    std::vector<cd> &f = part[t];
    f.assign((size_t)N, cd{0, 0});
    const double *xp = x + ch.lo, *yp = y ? y + ch.lo : nullptr, *zp = z ? z + ch.lo : nullptr;
    const cd *cp = c + ch.lo;
    if (dim == 1) dirft1d1<i64>(ch.n, xp, cp, sign, ms[0], f);
    else if (dim == 2) dirft2d1<i64>(ch.n, xp, yp, cp, sign, ms[0], ms[1], f);
    else dirft3d1<i64>(ch.n, xp, yp, zp, cp, sign, ms[0], ms[1], ms[2], f);
  }
#pragma omp parallel for num_threads(nthr)
  for (i64 m = 0; m < N; ++m) {
    cd s{0, 0};
    for (int t = 0; t < nthr; ++t) s += part[t][m];
    fk[m] = s;
  }
}

// c_j = sum_k fk[k] exp(i sign k.x_j)
void ref_dirft2(int dim, i64 M, const double *x, const double *y, const double *z, cd *c,
                int sign, const i64 *ms, const cd *fk, int nthr) {
  if (nthr < 1) nthr = omp_get_max_threads();
  if ((i64)nthr > M) nthr = M > 0 ? (int)M : 1;
#pragma omp parallel num_threads(nthr)
  {
    const int t = omp_get_thread_num();
    const Chunk ch = chunk_of(M, t, nthr);
    const double *xp = x + ch.lo, *yp = y ? y + ch.lo : nullptr, *zp = z ? z + ch.lo : nullptr;
    cd *cp = c + ch.lo;
    if (dim == 1) dirft1d2<i64>(ch.n, xp, cp, sign, ms[0], fk);
    else if (dim == 2) dirft2d2<i64>(ch.n, xp, yp, cp, sign, ms[0], ms[1], fk);
    else dirft3d2<i64>(ch.n, xp, yp, zp, cp, sign, ms[0], ms[1], ms[2], fk);
  }
}

// fk[k] = sum_j c_j exp(i sign (s_k x_j + t_k y_j + u_k z_j))
void ref_dirft3(int dim, i64 M, const double *x, const double *y, const double *z, const cd *c,
                int sign, i64 nk, const double *s, const double *t, const double *u, cd *fk,
                int nthr) {
  if (nthr < 1) nthr = omp_get_max_threads();
  if ((i64)nthr > nk) nthr = nk > 0 ? (int)nk : 1;
#pragma omp parallel num_threads(nthr)
  {
    const int th = omp_get_thread_num();
    const Chunk ch = chunk_of(nk, th, nthr);
    const double *sp = s + ch.lo, *tp = t ? t + ch.lo : nullptr, *up = u ? u + ch.lo : nullptr;
    cd *fp = fk + ch.lo;
    if (dim == 1) dirft1d3<i64>(M, x, c, sign, ch.n, sp, fp);
    else if (dim == 2) dirft2d3<i64>(M, x, y, c, sign, ch.n, sp, tp, fp);
    else dirft3d3<i64>(M, x, y, z, c, sign, ch.n, sp, tp, up, fp);
  }
}

// ||a - b||_2 / ||a||_2 (test/utils/norms.hpp:17-37), a = the trusted array
double ref_relerrtwonorm(i64 n, const cd *a, const cd *b) { return relerrtwonorm<i64>(n, a, b); }
double ref_twonorm(i64 n, const cd *a) { return twonorm<i64>(n, a); }
double ref_infnorm(i64 n, const cd *a) { return infnorm<i64>(n, a); }

}  // extern "C"
