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

#include <algorithm>

namespace b200 {

// Both kernels walk rows (fixed y, z index) so that the index arithmetic and the y/z factors
// are paid once per row and the x loop is a coalesced stream: block = kRowsPerBlock rows of one
// x tile of kTileX cells; blockIdx.z = transform in the batch.
constexpr int kRowsPerBlock = 8, kTileX = 2048, kGridThreads = 256;

template<class T, int DIM>
__global__ void __launch_bounds__(kGridThreads)
k_grid_to_modes(const typename CxOf<T>::type *__restrict__ fw,
                typename CxOf<T>::type *__restrict__ fk, ModeGeom<T> g) {
  using C = typename CxOf<T>::type;
  const int64_t nm = (int64_t)g.ms[0] * g.ms[1] * g.ms[2];
  const int64_t ng = (int64_t)g.nf[0] * g.nf[1] * g.nf[2];
  const C *fwb     = fw + (int64_t)blockIdx.z * ng;
  C *fkb           = fk + (int64_t)blockIdx.z * nm;
  const int nrows  = g.ms[1] * g.ms[2];
  for (int x0 = blockIdx.y * kTileX; x0 < g.ms[0]; x0 += gridDim.y * kTileX) {
    const int x1 = min(g.ms[0], x0 + kTileX);
    for (int i = 0; i < kRowsPerBlock; ++i) {
      const int row = blockIdx.x * kRowsPerBlock + i;
      if (row >= nrows) break;
      const int py = row % g.ms[1], pz = row / g.ms[1];
      int64_t src = 0;
      T p         = (T)1;
      if (DIM > 2) {
        const int k = mode_freq(pz, g.ms[2], g.modeord);
        src += (int64_t)(k >= 0 ? k : g.nf[2] + k) * g.nf[1] * g.nf[0];
        p = p / g.ph[2][k >= 0 ? k : -k];
      }
      if (DIM > 1) {
        const int k = mode_freq(py, g.ms[1], g.modeord);
        src += (int64_t)(k >= 0 ? k : g.nf[1] + k) * g.nf[0];
        p = p / g.ph[1][k >= 0 ? k : -k];
      }
      const C *srow = fwb + src;
      C *drow       = fkb + (int64_t)row * g.ms[0];
      for (int px = x0 + threadIdx.x; px < x1; px += kGridThreads) {
        const int k  = mode_freq(px, g.ms[0], g.modeord);
        const T div  = g.ph[0][k >= 0 ? k : -k];
        const C v    = srow[k >= 0 ? k : g.nf[0] + k];
        drow[px]     = C{mul_rn(p, v.x) / div, mul_rn(p, v.y) / div};
      }
    }
  }
}

template<class T, int DIM>
__global__ void __launch_bounds__(kGridThreads)
k_modes_to_grid(const typename CxOf<T>::type *__restrict__ fk,
                typename CxOf<T>::type *__restrict__ fw, ModeGeom<T> g) {
  using C = typename CxOf<T>::type;
  const int64_t nm = (int64_t)g.ms[0] * g.ms[1] * g.ms[2];
  const int64_t ng = (int64_t)g.nf[0] * g.nf[1] * g.nf[2];
  C *fwb           = fw + (int64_t)blockIdx.z * ng;
  const C *fkb     = fk + (int64_t)blockIdx.z * nm;
  const int nrows  = g.nf[1] * g.nf[2];
  for (int x0 = blockIdx.y * kTileX; x0 < g.nf[0]; x0 += gridDim.y * kTileX) {
    const int x1 = min(g.nf[0], x0 + kTileX);
    for (int i = 0; i < kRowsPerBlock; ++i) {
      const int row = blockIdx.x * kRowsPerBlock + i;
      if (row >= nrows) break;
      const int cy = row % g.nf[1], cz = row / g.nf[1];
      bool inside = true;
      int64_t src = 0;
      T p         = (T)1;
      if (DIM > 2) {
        int k = 0;
        inside = cell_freq(cz, g.ms[2], g.nf[2], k) && inside;
        src += (int64_t)mode_pos(k, g.ms[2], g.modeord) * g.ms[1] * g.ms[0];
        p = p / g.ph[2][k >= 0 ? k : -k];
      }
      if (DIM > 1) {
        int k = 0;
        inside = cell_freq(cy, g.ms[1], g.nf[1], k) && inside;
        src += (int64_t)mode_pos(k, g.ms[1], g.modeord) * g.ms[0];
        p = p / g.ph[1][k >= 0 ? k : -k];
      }
      const C *srow = fkb + src;
      C *drow       = fwb + (int64_t)row * g.nf[0];
      if (!inside) {  // 1 - (ms/nf)^(dim-1) of the rows: plain zero fill, 16 bytes per store
        if ((reinterpret_cast<uintptr_t>(drow + x0) & 15) == 0 && ((x1 - x0) * sizeof(C)) % 16 == 0) {
          uint4 *d4   = reinterpret_cast<uint4 *>(drow + x0);
          const int n = (int)((x1 - x0) * sizeof(C) / 16);
          for (int q = threadIdx.x; q < n; q += kGridThreads) d4[q] = make_uint4(0u, 0u, 0u, 0u);
        } else {
          for (int cx = x0 + threadIdx.x; cx < x1; cx += kGridThreads) drow[cx] = C{(T)0, (T)0};
        }
        continue;
      }
      for (int cx = x0 + threadIdx.x; cx < x1; cx += kGridThreads) {
        C out = C{(T)0, (T)0};
        int k = 0;
        if (cell_freq(cx, g.ms[0], g.nf[0], k)) {
          const T div = g.ph[0][k >= 0 ? k : -k];
          const C v   = srow[mode_pos(k, g.ms[0], g.modeord)];
          out         = C{mul_rn(p, v.x) / div, mul_rn(p, v.y) / div};
        }
        drow[cx] = out;
      }
    }
  }
}

template<class T>
void launch_grid_to_modes(int dim, int batch, const typename CxOf<T>::type *fw,
                          typename CxOf<T>::type *fk, const ModeGeom<T> &g, cudaStream_t st) {
  const int nrows = g.ms[1] * g.ms[2];
  dim3 grid((nrows + kRowsPerBlock - 1) / kRowsPerBlock,
            std::min((g.ms[0] + kTileX - 1) / kTileX, 65535), batch);
  if (dim == 1) k_grid_to_modes<T, 1><<<grid, kGridThreads, 0, st>>>(fw, fk, g);
  else if (dim == 2) k_grid_to_modes<T, 2><<<grid, kGridThreads, 0, st>>>(fw, fk, g);
  else k_grid_to_modes<T, 3><<<grid, kGridThreads, 0, st>>>(fw, fk, g);
}
template<class T>
void launch_modes_to_grid(int dim, int batch, const typename CxOf<T>::type *fk,
                          typename CxOf<T>::type *fw, const ModeGeom<T> &g, cudaStream_t st) {
  const int nrows = g.nf[1] * g.nf[2];
  dim3 grid((nrows + kRowsPerBlock - 1) / kRowsPerBlock,
            std::min((g.nf[0] + kTileX - 1) / kTileX, 65535), batch);
  if (dim == 1) k_modes_to_grid<T, 1><<<grid, kGridThreads, 0, st>>>(fk, fw, g);
  else if (dim == 2) k_modes_to_grid<T, 2><<<grid, kGridThreads, 0, st>>>(fk, fw, g);
  else k_modes_to_grid<T, 3><<<grid, kGridThreads, 0, st>>>(fk, fw, g);
}

static inline int blocks_for(int64_t n, int threads) {
  int64_t want = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
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
