// Point-driven (unsorted) spread / interp: see direct.cuh.
#include "direct.cuh"

#include "spreadinterp.cuh"  // atomic_add_cx

namespace b200 {

constexpr int kMaxNs = 16;

template<class T>
__device__ __forceinline__ void window_runtime(const T *__restrict__ coef, int ns, int nc, T X,
                                               bool clamp, int &i0, T *kv) {
  // stencil start and window argument of the leftmost cell (devmath.cuh stencil_start)
  const T c = ceil_t(sub_rn(X, (T)0.5 * (T)ns));
  i0        = (int)c;
  T x1      = sub_rn(c, X);
  if (clamp) {  // 1D: include/finufft/spread.hpp:116-121
    const T lo = (T)(-0.5) * (T)ns;
    x1 = x1 < lo ? lo : x1;
    x1 = x1 > lo + (T)1 ? lo + (T)1 : x1;
  }
  const T zz = fma_rn((T)2.0, x1, (T)(ns - 1));
  for (int j = 0; j < ns; ++j) {
    T r = coef[j];
    for (int k = 1; k < nc; ++k) r = fma_rn(r, zz, coef[k * ns + j]);
    kv[j] = r;
  }
}

template<class T, int DIM, bool SPREAD>
__global__ void __launch_bounds__(128)
k_direct(const T *__restrict__ coef, int ns, int nc, const T *__restrict__ x,
         const T *__restrict__ y, const T *__restrict__ z, uint32_t M, GridGeom<T> g,
         const typename CxOf<T>::type *__restrict__ c_in,
         typename CxOf<T>::type *__restrict__ c_out, typename CxOf<T>::type *fw) {
  using C = typename CxOf<T>::type;
  extern __shared__ __align__(16) unsigned char sm[];
  T *tab = reinterpret_cast<T *>(sm);
  for (int i = threadIdx.x; i < ns * nc; i += blockDim.x) tab[i] = coef[i];
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    T k1[kMaxNs], k2[kMaxNs], k3[kMaxNs];
    int i1 = 0, i2 = 0, i3 = 0;
    window_runtime<T>(tab, ns, nc, fold_rescale<T>(x[i], g.nf_t[0]), DIM == 1, i1, k1);
    if (DIM > 1) window_runtime<T>(tab, ns, nc, fold_rescale<T>(y[i], g.nf_t[1]), false, i2, k2);
    if (DIM > 2) window_runtime<T>(tab, ns, nc, fold_rescale<T>(z[i], g.nf_t[2]), false, i3, k3);
    C cj = C{(T)0, (T)0};
    if (SPREAD) cj = c_in[i];
    T ar = 0, ai = 0;
    for (int dz = 0; dz < (DIM > 2 ? ns : 1); ++dz) {
      const int gz = DIM > 2 ? grid_plane(g, wrap_index(i3 + dz, g.nf[2])) : 0;
      if (gz < 0) continue;
      const T wz = DIM > 2 ? k3[dz] : (T)1;
      for (int dy = 0; dy < (DIM > 1 ? ns : 1); ++dy) {
        const int gy = DIM > 1 ? wrap_index(i2 + dy, g.nf[1]) : 0;
        const T wyz  = DIM > 1 ? mul_rn(wz, k2[dy]) : wz;
        C *row       = fw + ((size_t)gz * g.nf[1] + gy) * (size_t)g.nf[0];
        T sr = 0, si = 0;
        for (int dx = 0; dx < ns; ++dx) {
          const int gx = wrap_index(i1 + dx, g.nf[0]);
          if (SPREAD) {
            const T w = mul_rn(wyz, k1[dx]);
            atomic_add_cx(row + gx, C{mul_rn(cj.x, w), mul_rn(cj.y, w)});
          } else {
            const C v = row[gx];
            sr        = fma_rn(v.x, k1[dx], sr);
            si        = fma_rn(v.y, k1[dx], si);
          }
        }
        if (!SPREAD) {
          ar = fma_rn(sr, wyz, ar);
          ai = fma_rn(si, wyz, ai);
        }
      }
    }
    if (!SPREAD) c_out[i] = C{ar, ai};
  }
}

template<class T>
cudaError_t launch_direct(bool spread, int dim, int ns, int nc, const T *coef_dev, const T *x,
                          const T *y, const T *z, uint32_t M, const GridGeom<T> &g,
                          const typename CxOf<T>::type *c_in, typename CxOf<T>::type *c_out,
                          typename CxOf<T>::type *fw, cudaStream_t st) {
  if (M == 0) return cudaSuccess;
  if (ns > kMaxNs) return cudaErrorInvalidValue;
  const int threads  = 128;
  const int64_t want = ((int64_t)M + threads - 1) / threads;
  const int blocks   = (int)(want > 148 * 16 ? 148 * 16 : want);
  const size_t shm   = sizeof(T) * (size_t)ns * nc;
#define B200_DIRECT(D, S) \
  k_direct<T, D, S><<<blocks, threads, shm, st>>>(coef_dev, ns, nc, x, y, z, M, g, c_in, c_out, fw)
  if (dim == 1) {
    if (spread) B200_DIRECT(1, true);
    else B200_DIRECT(1, false);
  } else if (dim == 2) {
    if (spread) B200_DIRECT(2, true);
    else B200_DIRECT(2, false);
  } else {
    if (spread) B200_DIRECT(3, true);
    else B200_DIRECT(3, false);
  }
#undef B200_DIRECT
  return cudaGetLastError();
}
template cudaError_t launch_direct<float>(bool, int, int, int, const float *, const float *,
                                          const float *, const float *, uint32_t,
                                          const GridGeom<float> &, const float2 *, float2 *,
                                          float2 *, cudaStream_t);
template cudaError_t launch_direct<double>(bool, int, int, int, const double *, const double *,
                                           const double *, const double *, uint32_t,
                                           const GridGeom<double> &, const double2 *, double2 *,
                                           double2 *, cudaStream_t);

}  // namespace b200
