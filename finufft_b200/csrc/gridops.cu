// Uniform-grid side of the transform: deconvolve + mode-order copy (type 1), amplify +
// zero-pad (type 2), and the element-wise multiplies type 3 needs.
//
// Spec: include/finufft/execute.hpp:69-237 (deconvolveshuffle1d/2d/3d): per dimension
// kmin = -(ms/2), kmax = (ms-1)/2; k >= 0 lives at fine index k, k < 0 at nf+k; modeord 0
// stores k ascending, modeord 1 stores k >= 0 first.  The divisor is applied as nested real
// divisions 1/phihat3 -> /phihat2 -> (x * p)/phihat1, the order the reference uses, so given
// equal tables the result is bit-identical.  Type 2 writes every fine-grid cell exactly once
// (zero where no mode maps), which replaces the reference GPU code's separate memset
// (src/cuda/execute.cu:82-84).
#include "gridops.cuh"

namespace b200 {

template<class T, int DIM>
__global__ void k_grid_to_modes(const typename CxOf<T>::type *__restrict__ fw,
                                typename CxOf<T>::type *__restrict__ fk, ModeGeom<T> g) {
  using C = typename CxOf<T>::type;
  const int64_t nm = (int64_t)g.ms[0] * g.ms[1] * g.ms[2];
  const int64_t ng = (int64_t)g.nf[0] * g.nf[1] * g.nf[2];
  const C *fwb     = fw + (int64_t)blockIdx.y * ng;
  C *fkb           = fk + (int64_t)blockIdx.y * nm;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < nm; m += stride) {
    int pos[3] = {(int)(m % g.ms[0]), (int)((m / g.ms[0]) % g.ms[1]),
                  (int)(m / ((int64_t)g.ms[0] * g.ms[1]))};
    int64_t src = 0, pitch = 1;
    int ak[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int kmin = -(g.ms[d] / 2), kmax = (g.ms[d] - 1) / 2;
      const int k = g.modeord == 0 ? pos[d] + kmin : (pos[d] <= kmax ? pos[d] : pos[d] - g.ms[d]);
      src += pitch * (k >= 0 ? k : g.nf[d] + k);
      pitch *= g.nf[d];
      ak[d] = k >= 0 ? k : -k;
    }
    T p = (T)1;
    if (DIM > 2) p = p / g.ph[2][ak[2]];
    if (DIM > 1) p = p / g.ph[1][ak[1]];
    const T div = g.ph[0][ak[0]];
    const C v   = fwb[src];
    fkb[m]      = C{mul_rn(p, v.x) / div, mul_rn(p, v.y) / div};
  }
}

template<class T, int DIM>
__global__ void k_modes_to_grid(const typename CxOf<T>::type *__restrict__ fk,
                                typename CxOf<T>::type *__restrict__ fw, ModeGeom<T> g) {
  using C = typename CxOf<T>::type;
  const int64_t nm = (int64_t)g.ms[0] * g.ms[1] * g.ms[2];
  const int64_t ng = (int64_t)g.nf[0] * g.nf[1] * g.nf[2];
  C *fwb           = fw + (int64_t)blockIdx.y * ng;
  const C *fkb     = fk + (int64_t)blockIdx.y * nm;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ng; c += stride) {
    int cell[3] = {(int)(c % g.nf[0]), (int)((c / g.nf[0]) % g.nf[1]),
                   (int)(c / ((int64_t)g.nf[0] * g.nf[1]))};
    int64_t src = 0, pitch = 1;
    int ak[3] = {0, 0, 0};
    bool inside = true;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const int kmin = -(g.ms[d] / 2), kmax = (g.ms[d] - 1) / 2;
      int k;
      if (cell[d] <= kmax) k = cell[d];
      else if (cell[d] >= g.nf[d] + kmin) k = cell[d] - g.nf[d];
      else {
        inside = false;
        k      = 0;
      }
      src += pitch * (g.modeord == 0 ? k - kmin : (k >= 0 ? k : g.ms[d] + k));
      pitch *= g.ms[d];
      ak[d] = k >= 0 ? k : -k;
    }
    C out = C{(T)0, (T)0};
    if (inside) {
      T p = (T)1;
      if (DIM > 2) p = p / g.ph[2][ak[2]];
      if (DIM > 1) p = p / g.ph[1][ak[1]];
      const T div = g.ph[0][ak[0]];
      const C v   = fkb[src];
      out         = C{mul_rn(p, v.x) / div, mul_rn(p, v.y) / div};
    }
    fwb[c] = out;
  }
}

static inline int blocks_for(int64_t n, int threads) {
  int64_t want = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template<class T>
void launch_grid_to_modes(int dim, int batch, const typename CxOf<T>::type *fw,
                          typename CxOf<T>::type *fk, const ModeGeom<T> &g, cudaStream_t st) {
  const int64_t nm = (int64_t)g.ms[0] * g.ms[1] * g.ms[2];
  dim3 grid(blocks_for(nm, 256), batch);
  if (dim == 1) k_grid_to_modes<T, 1><<<grid, 256, 0, st>>>(fw, fk, g);
  else if (dim == 2) k_grid_to_modes<T, 2><<<grid, 256, 0, st>>>(fw, fk, g);
  else k_grid_to_modes<T, 3><<<grid, 256, 0, st>>>(fw, fk, g);
}
template<class T>
void launch_modes_to_grid(int dim, int batch, const typename CxOf<T>::type *fk,
                          typename CxOf<T>::type *fw, const ModeGeom<T> &g, cudaStream_t st) {
  const int64_t ng = (int64_t)g.nf[0] * g.nf[1] * g.nf[2];
  dim3 grid(blocks_for(ng, 256), batch);
  if (dim == 1) k_modes_to_grid<T, 1><<<grid, 256, 0, st>>>(fk, fw, g);
  else if (dim == 2) k_modes_to_grid<T, 2><<<grid, 256, 0, st>>>(fk, fw, g);
  else k_modes_to_grid<T, 3><<<grid, 256, 0, st>>>(fk, fw, g);
}

// ---- type 3 element-wise helpers -----------------------------------------------------------
template<class T>
__global__ void k_cmul(const typename CxOf<T>::type *__restrict__ a,
                       const typename CxOf<T>::type *__restrict__ b,
                       typename CxOf<T>::type *__restrict__ out, int64_t n, int conj_b) {
  using C = typename CxOf<T>::type;
  const int64_t nn     = n;
  const C *ab          = a + (int64_t)blockIdx.y * nn;
  C *ob                = out + (int64_t)blockIdx.y * nn;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const C u = ab[i];
    C w       = b[i];
    if (conj_b) w.y = -w.y;
    ob[i] = C{u.x * w.x - u.y * w.y, u.x * w.y + u.y * w.x};
  }
}
template<class T>
void launch_cmul(int batch, const typename CxOf<T>::type *a, const typename CxOf<T>::type *b,
                 typename CxOf<T>::type *out, int64_t n, int conj_b, cudaStream_t st) {
  if (n == 0) return;
  dim3 grid(blocks_for(n, 256), batch);
  k_cmul<T><<<grid, 256, 0, st>>>(a, b, out, n, conj_b);
}

#define B200_INST(T)                                                                           \
  template void launch_grid_to_modes<T>(int, int, const CxOf<T>::type *, CxOf<T>::type *,      \
                                        const ModeGeom<T> &, cudaStream_t);                    \
  template void launch_modes_to_grid<T>(int, int, const CxOf<T>::type *, CxOf<T>::type *,      \
                                        const ModeGeom<T> &, cudaStream_t);                    \
  template void launch_cmul<T>(int, const CxOf<T>::type *, const CxOf<T>::type *,              \
                               CxOf<T>::type *, int64_t, int, cudaStream_t);
B200_INST(float)
B200_INST(double)

}  // namespace b200
