// Type 3 (nonuniform -> nonuniform): setpts-time rescaling / phase tables and the execute
// composition  prephase -> spread -> inner type-2 transform -> deconvolve.
//
// Spec: include/finufft/setpts.hpp:163-319 (centres and half-widths, nhg_type3 grid choice,
// x' = (x-C)/gam, prephase e^{+-i D.x}, targets s' = h gam (s-D), deconv = e^{+-i (s-D).C} /
// prod_d phihat(s'_d) with phihat(xi) = prefac * phi(grid_scale * xi), the analytic prolate
// self-transform of include/finufft_common/kernel.h:114-123 and plan.hpp:255-265) and
// include/finufft/execute.hpp:432-558 (per vector: multiply, spread onto the nf grid with no
// deconvolution, run the inner type-2 plan whose "modes" are that grid, multiply).
// The GPU reference does the same with thrust calls (src/cuda/setpts.cu:77-314,
// src/cuda/execute.cu:208-258); here the element-wise steps are three small kernels.
#include <algorithm>
#include <cmath>
#include <limits>

#include "engine.hpp"
#include "planmath.hpp"

namespace b200 {

namespace {

void cu(cudaError_t e) {
  if (e == cudaSuccess) return;
  cudaGetLastError();
  throw Failure{e == cudaErrorMemoryAllocation ? ERR_ALLOC : ERR_CUDA_FAILURE};
}

constexpr int kMmBlocks = 296, kMmThreads = 256;

// per-block min / max of an array; the <= 296 partial pairs are finished on the host
template<class T>
__global__ void k_minmax(const T *__restrict__ v, int64_t n, T *__restrict__ part) {
  __shared__ T slo[kMmThreads / 32], shi[kMmThreads / 32];
  T lo = INFINITY, hi = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const T a = v[i];
    lo        = a < lo ? a : lo;
    hi        = a > hi ? a : hi;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const T l2 = __shfl_xor_sync(0xffffffffu, lo, d), h2 = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) slo[threadIdx.x >> 5] = lo, shi[threadIdx.x >> 5] = hi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kMmThreads / 32; ++w) {
      lo = slo[w] < lo ? slo[w] : lo;
      hi = shi[w] > hi ? shi[w] : hi;
    }
    part[2 * blockIdx.x]     = lo;
    part[2 * blockIdx.x + 1] = hi;
  }
}

// half-width and centre of an array, include/finufft_common/utils.h:77-90 (arraywidcen)
template<class T>
void width_centre(const T *d_v, int64_t n, T *d_part, std::vector<T> &h_part, cudaStream_t st,
                  T &w, T &c) {
  T lo = INFINITY, hi = -INFINITY;
  if (n > 0) {
    k_minmax<T><<<kMmBlocks, kMmThreads, 0, st>>>(d_v, n, d_part);
    cu(cudaGetLastError());
    cu(cudaMemcpyAsync(h_part.data(), d_part, sizeof(T) * 2 * kMmBlocks, cudaMemcpyDeviceToHost,
                       st));
    cu(cudaStreamSynchronize(st));
    for (int b = 0; b < kMmBlocks; ++b) {
      lo = h_part[2 * b] < lo ? h_part[2 * b] : lo;
      hi = h_part[2 * b + 1] > hi ? h_part[2 * b + 1] : hi;
    }
  }
  w = (hi - lo) / 2;
  c = (hi + lo) / 2;
  if (std::abs(c) < (T)0.1 * w) {
    w += std::abs(c);
    c = 0;
  }
}

template<class T> struct T3Geom {
  int dim;
  T C[3], D[3], ig[3], hg[3];  // centres, 1/gam, h*gam
  T grid_scale, prefac, isign;
  int ns, nc, any_D, do_phase;
};

__device__ __forceinline__ void polar_unit(float ph, float &re, float &im) { sincosf(ph, &im, &re); }
__device__ __forceinline__ void polar_unit(double ph, double &re, double &im) { sincos(ph, &im, &re); }

// x' = (x - C) / gam ; prephase = e^{i isign D.x}
template<class T>
__global__ void k_t3_sources(T3Geom<T> g, int64_t M, const T *__restrict__ x,
                             const T *__restrict__ y, const T *__restrict__ z,
                             T *__restrict__ xp, T *__restrict__ yp, T *__restrict__ zp,
                             typename CxOf<T>::type *__restrict__ prephase) {
  using C = typename CxOf<T>::type;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M;
       j += (int64_t)gridDim.x * blockDim.x) {
    const T xj = x[j], yj = g.dim > 1 ? y[j] : (T)0, zj = g.dim > 2 ? z[j] : (T)0;
    xp[j] = mul_rn(sub_rn(xj, g.C[0]), g.ig[0]);
    if (g.dim > 1) yp[j] = mul_rn(sub_rn(yj, g.C[1]), g.ig[1]);
    if (g.dim > 2) zp[j] = mul_rn(sub_rn(zj, g.C[2]), g.ig[2]);
    C p{(T)1, (T)0};
    if (g.any_D) {
      T ph = mul_rn(g.D[0], xj);
      if (g.dim > 1) ph += mul_rn(g.D[1], yj);
      if (g.dim > 2) ph += mul_rn(g.D[2], zj);
      polar_unit(g.isign * ph, p.x, p.y);
    }
    prephase[j] = p;
  }
}

// the tabulated window at grid-unit argument x in (-ns/2, ns/2], zero outside
// (include/finufft/spreadinterp.hpp:57-94 applied to one argument)
template<class T>
__device__ __forceinline__ T eval_table_dev(T x, int ns, int nc, const T *__restrict__ coef) {
  const T ns2 = (T)ns / (T)2;
  T res       = (T)0;
  for (int i = 0; i < ns; ++i) {
    if (x > -ns2 + (T)i && x <= -ns2 + (T)(i + 1)) {
      const T zz = fma_rn((T)2, sub_rn(x, (T)i), (T)(ns - 1));
      for (int j = 0; j < nc; ++j) res = fma_rn(res, zz, coef[j * ns + i]);
      break;
    }
  }
  return res;
}

// s' = h gam (s - D) ; deconv = e^{i isign (s-D).C} / prod_d prefac*phi(grid_scale*s'_d)
template<class T>
__global__ void k_t3_targets(T3Geom<T> g, int64_t N, const T *__restrict__ s,
                             const T *__restrict__ t, const T *__restrict__ u,
                             const T *__restrict__ coef, T *__restrict__ sp, T *__restrict__ tp,
                             T *__restrict__ up, typename CxOf<T>::type *__restrict__ deconv) {
  using C = typename CxOf<T>::type;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < N;
       k += (int64_t)gridDim.x * blockDim.x) {
    const T in[3] = {s[k], g.dim > 1 ? t[k] : (T)0, g.dim > 2 ? u[k] : (T)0};
    T *out[3]     = {sp, tp, up};
    T ph = 0, phi = 1;
    for (int d = 0; d < g.dim; ++d) {
      const T rel = sub_rn(in[d], g.D[d]);
      const T v   = mul_rn(g.hg[d], rel);
      phi *= g.prefac * eval_table_dev<T>(mul_rn(g.grid_scale, v), g.ns, g.nc, coef);
      if (g.do_phase) ph += mul_rn(rel, g.C[d]);
      out[d][k] = v;
    }
    const T amp = (T)1 / phi;
    C r{amp, (T)0};
    if (g.do_phase) {
      T re, im;
      polar_unit(g.isign * ph, re, im);
      r = C{amp * re, amp * im};
    }
    deconv[k] = r;
  }
}

int grid1d(int64_t n) {
  const int64_t want = (n + 255) / 256;
  return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}

}  // namespace

template<class T>
void Engine<T>::setpts_type3(int64_t M_, const T *x, const T *y, const T *z, int64_t N, const T *s,
                             const T *t, const T *u) {
  if (N < 0) throw Failure{ERR_NUM_NU_PTS_INVALID};
  if (N > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
  const T *src[3] = {x, y, z}, *tgt[3] = {s, t, u};
  for (int d = 0; d < dim; ++d)
    if ((M_ > 0 && !src[d]) || (N > 0 && !tgt[d])) throw Failure{ERR_INVALID_ARGUMENT};
  cudaStream_t st = opts.stream;
  M  = M_;
  nk = N;
  mark(4);

  // centres / half-widths of sources and targets, then the grid (setpts.hpp:176-214)
  DevBuf<T> part;
  part.alloc(2 * kMmBlocks);
  std::vector<T> h_part(2 * kMmBlocks);
  T X[3] = {0, 0, 0}, S[3] = {0, 0, 0};
  for (int d = 0; d < 3; ++d) t3C_[d] = t3D_[d] = 0, t3h_[d] = 0, t3gam_[d] = 1;
  for (int d = 0; d < dim; ++d) {
    width_centre<T>(src[d], M, part.p, h_part, st, X[d], t3C_[d]);
    width_centre<T>(tgt[d], N, part.p, h_part, st, S[d], t3D_[d]);
    if (M == 0) X[d] = 0, t3C_[d] = 0;
    if (N == 0) S[d] = 0, t3D_[d] = 0;
  }
  if (opts.auto_sigma) {
    // upsampfac = 0 on the host API: sigma3 from the half-widths and the point counts, as the
    // reference CPU library does here (include/finufft/setpts.hpp:186-200); the inner type 2
    // below then picks its own sigma from its grid and the number of targets (:302-304)
    const double Xd[3] = {(double)X[0], (double)X[1], (double)X[2]};
    const double Sd[3] = {(double)S[0], (double)S[1], (double)S[2]};
    const double s_new = choose_sigma_type3(tol_req_, dim, std::is_same<T, float>::value,
                                            (double)M, (double)N, Xd, Sd);
    if (std::abs(s_new - sigma) > 1e-12) {
      cu(cudaStreamSynchronize(st));
      sigma = s_new;
      tol   = tol_req_;
      plan_kernel();
      coef_dev_.release();
    }
  }
  for (int d = 0; d < 3; ++d) nf[d] = 1;
  for (int d = 0; d < dim; ++d) {
    int64_t n1;
    double h, gam;
    type3_grid(sigma, (double)X[d], (double)S[d], ns, n1, h, gam);
    if (n1 > std::numeric_limits<int32_t>::max()) throw Failure{ERR_MAXNALLOC};
    nf[d]     = n1;
    t3h_[d]   = (T)h;
    t3gam_[d] = (T)gam;
  }
  plan_grid();  // geometry + spread grid (no FFT / Fourier series: the inner plan has those)

  T3Geom<T> g;
  g.dim = dim;
  for (int d = 0; d < 3; ++d) {
    g.C[d]  = t3C_[d];
    g.D[d]  = t3D_[d];
    g.ig[d] = (T)1 / t3gam_[d];
    g.hg[d] = t3h_[d] * t3gam_[d];
  }
  double gs, pf;
  selfft_params<T>(ns, beta, nc, coef.data(), gs, pf);
  g.grid_scale = (T)gs;
  g.prefac     = (T)pf;
  g.isign      = sign >= 0 ? (T)1 : (T)-1;
  g.ns         = ns;
  g.nc         = nc;
  g.any_D      = (t3D_[0] != 0 || t3D_[1] != 0 || t3D_[2] != 0) ? 1 : 0;
  const bool finiteC = std::isfinite((double)t3C_[0]) && std::isfinite((double)t3C_[1]) &&
                       std::isfinite((double)t3C_[2]);
  g.do_phase = (finiteC && (t3C_[0] != 0 || t3C_[1] != 0 || t3C_[2] != 0)) ? 1 : 0;

  for (int d = 0; d < dim; ++d) {
    xp_[d].alloc(std::max<int64_t>(M, 1));
    sp_[d].alloc(std::max<int64_t>(N, 1));
  }
  prephase_.alloc(std::max<int64_t>(M, 1));
  deconv_.alloc(std::max<int64_t>(N, 1));
  cp_.alloc(std::max<int64_t>(M, 1));
  DevBuf<T> dcoef;
  dcoef.alloc(coef.size());
  cu(cudaMemcpyAsync(dcoef.p, coef.data(), sizeof(T) * coef.size(), cudaMemcpyHostToDevice, st));
  if (M > 0) {
    k_t3_sources<T><<<grid1d(M), 256, 0, st>>>(g, M, x, y, z, xp_[0].p, xp_[1].p, xp_[2].p,
                                               prephase_.p);
    ++launches;
  }
  if (N > 0) {
    k_t3_targets<T><<<grid1d(N), 256, 0, st>>>(g, N, s, t, u, dcoef.p, sp_[0].p, sp_[1].p, sp_[2].p,
                                               deconv_.p);
    ++launches;
  }
  cu(cudaGetLastError());

  for (int d = 0; d < dim; ++d)
    if (nf[d] < 2 * ns) throw Failure{ERR_SPREAD_BOX_SMALL};
  sort_points(xp_[0].p, xp_[1].p, xp_[2].p);  // syncs the stream before returning

  // inner type-2 plan on the nf grid (setpts.hpp:286-319); one vector at a time
  EngineOpts io        = opts;
  io.upsampfac         = opts.auto_sigma ? 0.0 : sigma;
  io.spreadinterponly  = 0;
  io.modeord           = 0;
  io.maxbatch          = 1;
  io.check_sigma       = 0;
  io.allow_eps_too_small = 1;
  inner_.reset();
  inner_.reset(new Engine<T>(2, dim, nf, sign, 1, tol, io));
  inner_->setpts(N, sp_[0].p, sp_[1].p, sp_[2].p, 0, nullptr, nullptr, nullptr);
  cu(cudaStreamSynchronize(st));  // dcoef / part are freed on return
  mark(5);
}

template<class T> void Engine<T>::exec_type3(C *c, C *fk, bool adjoint) {
  if (!inner_) throw Failure{ERR_PLAN_NOTVALID};
  cudaStream_t st = opts.stream;
  const int64_t G = grid_cells();
  if (adjoint) ck_.alloc((size_t)std::max<int64_t>(nk, 1));
  for (int b = 0; b < ntr; ++b) {
    C *cb  = c + (int64_t)b * M;
    C *fkb = fk + (int64_t)b * nk;
    if (!adjoint) {
      order_[0] = 0, order_[1] = 1, order_[2] = 2;
      mark(0);
      launch_cmul<T>(1, cb, prephase_.p, cp_.p, M, 0, st);
      ++launches;
      cu(cudaMemsetAsync(fw_.p, 0, sizeof(C) * (size_t)G, st));
      run_spread(cp_.p, fw_.p);
      mark(1);
      const uint64_t before = inner_->launches;
      inner_->execute(fkb, fw_.p, false);
      launches += inner_->launches - before;
      mark(2);
      launch_cmul<T>(1, fkb, deconv_.p, fkb, nk, 0, st);
      ++launches;
      mark(3);
    } else {
      // adjoint (include/finufft/execute.hpp:515-545): conj deconvolve, adjoint of the inner
      // type 2 (a type 1 from the targets onto the fw grid), interpolate, conj post-phase
      order_[0] = 2, order_[1] = 1, order_[2] = 0;
      mark(0);
      launch_cmul<T>(1, fkb, deconv_.p, ck_.p, nk, 1, st);
      ++launches;
      mark(1);
      const uint64_t before = inner_->launches;
      inner_->execute(ck_.p, fw_.p, true);
      launches += inner_->launches - before;
      mark(2);
      run_interp(cb, fw_.p);
      launch_cmul<T>(1, cb, prephase_.p, cb, M, 1, st);
      ++launches;
      mark(3);
    }
  }
  cu(cudaGetLastError());
}

template void Engine<float>::setpts_type3(int64_t, const float *, const float *, const float *,
                                          int64_t, const float *, const float *, const float *);
template void Engine<double>::setpts_type3(int64_t, const double *, const double *,
                                           const double *, int64_t, const double *,
                                           const double *, const double *);
template void Engine<float>::exec_type3(float2 *, float2 *, bool);
template void Engine<double>::exec_type3(double2 *, double2 *, bool);
}  // namespace b200
