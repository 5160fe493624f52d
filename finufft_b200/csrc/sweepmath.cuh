// Arithmetic helpers shared by the sweep kernels (sweep2d.cu, sweep3d.cu): packed FP32 pairs
// (Blackwell FFMA2 / FMUL2 with a broadcast scalar), complex accumulate for float and double,
// compile-time loops, and the packed Horner evaluation of the window polynomials.
//
// Spec of the window evaluation: include/finufft/spreadinterp.hpp:85-92 (reference CPU, scalar
// Horner, highest degree first, fused multiply-add per step); the packed form evaluates two
// panels per instruction with exactly the same roundings.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "devmath.cuh"

namespace b200 {

// (re,im) = s * (wr,wi) + (ar,ai): one packed FFMA2 (the scalar operand is broadcast in hardware)
__device__ __forceinline__ float2 ffma2_s(float s, float2 w, float2 acc) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n"
      " mov.b64 ra, {%2,%2};\n mov.b64 rb, {%3,%4};\n mov.b64 rc, {%5,%6};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(s), "f"(w.x), "f"(w.y), "f"(acc.x), "f"(acc.y));
  return d;
}
__device__ __forceinline__ float2 fmul2_s(float s, float2 w) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n"
      " mov.b64 ra, {%2,%2};\n mov.b64 rb, {%3,%4};\n"
      " mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(s), "f"(w.x), "f"(w.y));
  return d;
}

// real-times-complex accumulate / scale in the plan's precision
__device__ __forceinline__ float2 cx_fma(float s, float2 w, float2 acc) { return ffma2_s(s, w, acc); }
__device__ __forceinline__ float2 cx_mul(float s, float2 w) { return fmul2_s(s, w); }
__device__ __forceinline__ double2 cx_fma(double s, double2 w, double2 acc) {
  return double2{__fma_rn(s, w.x, acc.x), __fma_rn(s, w.y, acc.y)};
}
__device__ __forceinline__ double2 cx_mul(double s, double2 w) {
  return double2{__dmul_rn(s, w.x), __dmul_rn(s, w.y)};
}

// statically unrolled loop with the index available as a compile-time constant
template<int I, int N, class F> __device__ __forceinline__ void static_for(F &&f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// Horner table of a float plan with the panels padded to an even count, so that two
// neighbouring panels form one aligned pair: c[k*COLS + j], k = 0 highest degree, columns
// j >= NS are zero.
template<int NS> struct alignas(16) PairTab {
  static constexpr int COLS = 2 * ((NS + 1) / 2);
  static constexpr int ROWS = TableRows<NS>::value;
  float c[ROWS * COLS];
};
// the table a kernel carries: packed pairs for float, the plain layout for double
template<class T, int NS> struct SweepTab { using type = WindowTable<T, NS>; };
template<int NS> struct SweepTab<float, NS> { using type = PairTab<NS>; };

template<int NS>
__host__ inline void fill_table(PairTab<NS> &tab, int nc, const float *coef) {
  for (int k = 0; k < PairTab<NS>::ROWS; ++k)
    for (int j = 0; j < PairTab<NS>::COLS; ++j) {
      const int src                  = k - (PairTab<NS>::ROWS - nc);
      tab.c[k * PairTab<NS>::COLS + j] = (src >= 0 && j < NS) ? coef[src * NS + j] : 0.f;
    }
}
template<class T, int NS>
__host__ inline void fill_table(WindowTable<T, NS> &tab, int nc, const T *coef) {
  constexpr int rows = TableRows<NS>::value;
  for (int k = 0; k < rows; ++k)
    for (int j = 0; j < NS; ++j) {
      const int src     = k - (rows - nc);
      tab.c[k * NS + j] = src >= 0 ? coef[src * NS + j] : (T)0;
    }
}

// All NS window values at offset x1 into out[0..NS) (out has room for an even count).
template<int NS>
__device__ __forceinline__ void eval_window_t(const PairTab<NS> &tab, float x1, float *out) {
  constexpr int COLS = PairTab<NS>::COLS;
  const float z      = fma_rn(2.0f, x1, (float)(NS - 1));
#pragma unroll
  for (int p = 0; p < COLS / 2; ++p) {
    float2 r = *reinterpret_cast<const float2 *>(&tab.c[2 * p]);
#pragma unroll
    for (int k = 1; k < PairTab<NS>::ROWS; ++k)
      r = ffma2_s(z, r, *reinterpret_cast<const float2 *>(&tab.c[k * COLS + 2 * p]));
    out[2 * p]     = r.x;
    out[2 * p + 1] = r.y;
  }
}
template<class T, int NS>
__device__ __forceinline__ void eval_window_t(const WindowTable<T, NS> &tab, T x1, T *out) {
  eval_window<T, NS>(tab, x1, out);
}

}  // namespace b200
