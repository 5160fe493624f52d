// Declarations for the uniform-grid kernels (gridops.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "devmath.cuh"

namespace b200 {

template<class T> struct ModeGeom {
  int ms[3];        // modes per dim (1 for unused dims)
  int nf[3];        // fine grid per dim
  int modeord;      // 0: ascending k, 1: FFT order
  const T *ph[3];   // window Fourier series, indices 0..nf/2, device pointers
};

// mode index (0..ms-1 in storage order) -> signed frequency k
__device__ __forceinline__ int mode_freq(int pos, int ms, int modeord) {
  const int kmin = -(ms / 2), kmax = (ms - 1) / 2;
  return modeord == 0 ? pos + kmin : (pos <= kmax ? pos : pos - ms);
}
// fine-grid cell -> signed frequency, false if the cell is outside the kept band
__device__ __forceinline__ bool cell_freq(int cell, int ms, int nf, int &k) {
  const int kmin = -(ms / 2), kmax = (ms - 1) / 2;
  if (cell <= kmax) k = cell;
  else if (cell >= nf + kmin) k = cell - nf;
  else return false;
  return true;
}
__device__ __forceinline__ int mode_pos(int k, int ms, int modeord) {
  return modeord == 0 ? k + ms / 2 : (k >= 0 ? k : ms + k);
}

template<class T>
void launch_grid_to_modes(int dim, int batch, const typename CxOf<T>::type *fw,
                          typename CxOf<T>::type *fk, const ModeGeom<T> &g, cudaStream_t st);
template<class T>
void launch_modes_to_grid(int dim, int batch, const typename CxOf<T>::type *fk,
                          typename CxOf<T>::type *fw, const ModeGeom<T> &g, cudaStream_t st);
template<class T>
void launch_cmul(int batch, const typename CxOf<T>::type *a, const typename CxOf<T>::type *b,
                 typename CxOf<T>::type *out, int64_t n, int conj_b, cudaStream_t st);

}  // namespace b200
