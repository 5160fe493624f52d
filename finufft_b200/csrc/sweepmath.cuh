// Arithmetic helpers shared by the sweep kernels (sweep2d.cu, sweep3d.cu): packed FP32 pairs
// (Blackwell FFMA2 / FMUL2 with a broadcast scalar), complex accumulate for float and double,
// compile-time loops, and the packed Horner evaluation of the window polynomials.
//
// Spec of the window evaluation: include/finufft/spreadinterp.hpp:85-92 (reference CPU, scalar
// Horner, highest degree first, fused multiply-add per step).  PairTab evaluates two panels per
// instruction with exactly the same roundings; EOTab (the default for float) uses the even/odd
// symmetry of the window and agrees to ~1 ulp of the window peak.
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
// Even/odd form of the same polynomials (float).  The window is even, so panel NS-1-j is panel j
// mirrored: p_{NS-1-j}(z) = p_j(-z), i.e. p_j = E(z^2) + z O(z^2) and its mirror E - z O: one
// packed Horner in z^2 on the pair (E, O) gives two window values (the reference CPU library
// evaluates this way when its SIMD width is below NS, include/finufft/simd.hpp:373-422).
// c[i][p] = (coefficient of z^(2i), coefficient of z^(2i+1)) of panel p, p < (NS+1)/2.  The fitted
// table is symmetric to rounding only, so results differ from the plain Horner form by ~1 ulp of
// the window peak (checked: <= 2.4e-7 absolute in float).
template<int NS> struct alignas(16) EOTab {
  static constexpr int ROWS = TableRows<NS>::value;
  static constexpr int NE   = (ROWS + 1) / 2;  // terms of E (O is padded to the same count)
  static constexpr int NP   = (NS + 1) / 2;    // panel pairs (the middle panel pairs with itself)
  float2 c[NE * NP];
};
#ifndef B200_EVENODD
#define B200_EVENODD 1
#endif
// the table a kernel carries: packed for float, the plain layout for double
template<class T, int NS> struct SweepTab { using type = WindowTable<T, NS>; };
#if B200_EVENODD
template<int NS> struct SweepTab<float, NS> { using type = EOTab<NS>; };
#else
template<int NS> struct SweepTab<float, NS> { using type = PairTab<NS>; };
#endif

template<int NS> __host__ inline void fill_table(EOTab<NS> &tab, int nc, const float *coef) {
  constexpr int ROWS = EOTab<NS>::ROWS;
  // a(d, j) = coefficient of z^d of panel j; table rows are highest degree first, padded on top
  auto a = [&](int d, int j) -> float {
    if (d > ROWS - 1) return 0.f;
    const int src = (ROWS - 1 - d) - (ROWS - nc);
    return src >= 0 ? coef[src * NS + j] : 0.f;
  };
  for (int i = 0; i < EOTab<NS>::NE; ++i)
    for (int p = 0; p < EOTab<NS>::NP; ++p)
      tab.c[i * EOTab<NS>::NP + p] = float2{a(2 * i, p), a(2 * i + 1, p)};
}

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
template<int NS>
__device__ __forceinline__ void eval_window_t(const EOTab<NS> &tab, float x1, float *out) {
  constexpr int NE = EOTab<NS>::NE, NP = EOTab<NS>::NP;
  const float z  = fma_rn(2.0f, x1, (float)(NS - 1));
  const float z2 = mul_rn(z, z);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float2 r = tab.c[(NE - 1) * NP + p];
#pragma unroll
    for (int i = NE - 2; i >= 0; --i) r = ffma2_s(z2, r, tab.c[i * NP + p]);
    out[p] = fma_rn(z, r.y, r.x);
    if (NS - 1 - p != p) out[NS - 1 - p] = fma_rn(-z, r.y, r.x);
  }
}
template<class T, int NS>
__device__ __forceinline__ void eval_window_t(const WindowTable<T, NS> &tab, T x1, T *out) {
  eval_window<T, NS>(tab, x1, out);
}

}  // namespace b200
