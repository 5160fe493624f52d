// Device-side scalar helpers shared by every kernel: explicitly rounded arithmetic (so nvcc
// cannot contract or reassociate the steps whose bits must match the CPU library), the
// fold-rescale of a nonuniform coordinate, and the piecewise-polynomial window evaluator.
//
// Spec: fold_rescale   include/finufft/simd.hpp:318-325 (reference CPU, round-to-nearest),
//       stencil start  include/finufft/spread.hpp:328-333,
//       Horner         include/finufft/spreadinterp.hpp:85-92.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

template<class T> struct CxOf;
template<> struct CxOf<float> { using type = float2; };
template<> struct CxOf<double> { using type = double2; };

__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float floor_t(float a) { return floorf(a); }
__device__ __forceinline__ double floor_t(double a) { return floor(a); }
__device__ __forceinline__ float ceil_t(float a) { return ceilf(a); }
__device__ __forceinline__ double ceil_t(double a) { return ceil(a); }

// x (any real) -> [0, N]:  r = fma(x, 1/(2 pi), 0.5);  (r - floor r) * N, all round-to-nearest in T.
template<class T> __device__ __forceinline__ T fold_rescale(T x, T n_as_t) {
  const T r = fma_rn(x, (T)0.159154943091895335768883763372514362, (T)0.5);
  return mul_rn(sub_rn(r, floor_t(r)), n_as_t);
}

// Stencil placement for a rescaled coordinate X: leftmost cell i = ceil(X - ns/2) and the
// window argument of that cell x1 = i - X in [-ns/2, -ns/2+1].
template<class T, int NS> __device__ __forceinline__ void stencil_start(T X, int &i0, T &x1) {
  const T c = ceil_t(sub_rn(X, (T)(0.5 * NS)));
  i0        = (int)c;
  x1        = sub_rn(c, X);
}

// Rows of the polynomial table carried as a kernel parameter: NCP = min(19, NS+3) rows of NS
// coefficients, highest degree first, zero rows on top when the plan needs fewer.
template<int NS> struct TableRows {
  static constexpr int value = (NS + 3 < 19) ? NS + 3 : 19;
};
template<class T, int NS> struct WindowTable {
  T c[TableRows<NS>::value * NS];
};

// All NS window values for one coordinate; `out` may be registers or shared memory.
template<class T, int NS>
__device__ __forceinline__ void eval_window(const WindowTable<T, NS> &tab, T x1, T *out,
                                            int stride = 1) {
  const T z = fma_rn((T)2.0, x1, (T)(NS - 1));
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    T r = tab.c[j];
#pragma unroll
    for (int k = 1; k < TableRows<NS>::value; ++k) r = fma_rn(r, z, tab.c[k * NS + j]);
    out[j * stride] = r;
  }
}

__device__ __forceinline__ int wrap_index(int i, int n) {
  if (i < 0) i += n;
  while (i >= n) i -= n;
  return i;
}

}  // namespace b200
