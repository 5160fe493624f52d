// Row-sweep spread / interp kernels for 2D (see sweep2d.cuh for the design).  Included by the
// per-precision translation units sweep2d_f32.cu / sweep2d_f64.cu.
#pragma once
#include <limits.h>

#include <type_traits>

#include "spreadinterp.cuh"
#include "sweep2d.cuh"
#include "sweepmath.cuh"

namespace b200 {

template<class T, int NS, bool SPREAD> struct Sweep2Cfg {
  using WN = Sweep2Win<NS>;
  using C  = typename CxOf<T>::type;
  using XE = typename std::conditional<SPREAD, C, T>::type;  // x-window record element
  static constexpr int W = WN::W, S = WN::S, G = WN::G, XB = WN::XB;
  static constexpr int HL  = NS / 2;              // cells below a bin a stencil can reach
  static constexpr int YR  = kBinY + NS;          // y rows of the register window
  static constexpr int CH  = 32;                  // points per chunk: one per lane
  static constexpr int WPB = 4;                   // warps (= work items) per block
  static constexpr int KV  = 16 / (int)sizeof(T);  // values per 16-byte shared-memory access
  static constexpr int KP  = (NS + KV - 1) / KV * KV;  // y-window record pitch
  // x-window record [slot][point]: pitch chosen so that the lanes of one shared-memory phase
  // (128 bytes) read distinct banks: lane (g, a) reads slot a of point q+g
  static constexpr int LPP = 128 / (int)sizeof(XE);
  static constexpr int XP  = CH + (LPP / W > 1 ? LPP / W : 1);
  static constexpr size_t KY_BYTES  = (size_t)CH * KP * sizeof(T);
  static constexpr size_t XW_BYTES  = ((size_t)W * XP * sizeof(XE) + 15) / 16 * 16;
  static constexpr size_t OUT_BYTES = SPREAD ? 0 : (size_t)CH * sizeof(C);
  static constexpr size_t WARP_BYTES = KY_BYTES + XW_BYTES + OUT_BYTES;
};

template<class T, int NS> struct Sweep2Args {
  using C = typename CxOf<T>::type;
  Sweep2Points<T> pts;
  GridGeom<T> g;
  typename SweepTab<T, NS>::type tab;
  const C *c_in;
  C *c_out;
  C *fw;
};

template<class T> __device__ __forceinline__ T shfl_xor_t(T v, int d) {
  return __shfl_xor_sync(0xffffffffu, v, d);
}

// NS window values of one point from its 16-byte aligned record
template<class T, int NS, int KP>
__device__ __forceinline__ void load_ky(const T *rec, T (&ky)[KP]) {
  constexpr int KV = 16 / (int)sizeof(T);
#pragma unroll
  for (int v = 0; v < KP / KV; ++v) {
    if constexpr (sizeof(T) == 4) {
      const float4 q = *reinterpret_cast<const float4 *>(rec + 4 * v);
      ky[4 * v] = q.x, ky[4 * v + 1] = q.y, ky[4 * v + 2] = q.z, ky[4 * v + 3] = q.w;
    } else {
      const double2 q = *reinterpret_cast<const double2 *>(rec + 2 * v);
      ky[2 * v] = q.x, ky[2 * v + 1] = q.y;
    }
  }
}

// points [p, e) of the current chunk, all with y stencil start J: every lane group takes one
template<class T, int NS, int J, class CF>
__device__ __forceinline__ void spread_run2(typename CF::C (&acc)[CF::YR], const T *sky,
                                            const typename CF::C *sxw, int p, int e, int g,
                                            int la) {
  for (int q = p; q < e; q += CF::G) {
    const int idx = q + g;
    if (idx < e) {
      T ky[CF::KP];
      load_ky<T, NS, CF::KP>(sky + idx * CF::KP, ky);
      const typename CF::C cw = sxw[la * CF::XP + idx];
#pragma unroll
      for (int t = 0; t < NS; ++t) acc[J + t] = cx_fma(ky[t], cw, acc[J + t]);
    }
  }
}
template<class T, int NS, int J, class CF>
__device__ __forceinline__ void interp_run2(const typename CF::C (&acc)[CF::YR], const T *sky,
                                            const T *sxw, typename CF::C *sout, int p, int e,
                                            int g, int la) {
  for (int q = p; q < e; q += CF::G) {
    const int idx    = q + g;
    const bool valid = idx < e;
    const int idc    = valid ? idx : p;
    T ky[CF::KP];
    load_ky<T, NS, CF::KP>(sky + idc * CF::KP, ky);
    const T wx          = sxw[la * CF::XP + idc];
    typename CF::C v = cx_mul(ky[0], acc[J]);
#pragma unroll
    for (int t = 1; t < NS; ++t) v = cx_fma(ky[t], acc[J + t], v);
    v = cx_mul(wx, v);
#pragma unroll
    for (int d = 1; d < CF::W; d <<= 1) {
      v.x += shfl_xor_t(v.x, d);
      v.y += shfl_xor_t(v.y, d);
    }
    if (valid && la == 0) sout[idx] = v;
  }
}

// One warp per work item: a run of consecutive points of one row of bins (i2), which the
// refined bin order keeps sorted by x window position.
template<class T, int NS, bool SPREAD>
__global__ void __launch_bounds__(Sweep2Cfg<T, NS, SPREAD>::WPB * 32)
k_sweep2(const Sweep2Args<T, NS> a) {
  using CF = Sweep2Cfg<T, NS, SPREAD>;
  using C  = typename CF::C;
  using XE = typename CF::XE;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t it = blockIdx.x * CF::WPB + warp;
  if (it >= a.pts.nitems) return;  // warps never synchronise with each other
  unsigned char *base = smem + (size_t)warp * CF::WARP_BYTES;
  T *sky   = reinterpret_cast<T *>(base);
  XE *sxw  = reinterpret_cast<XE *>(base + CF::KY_BYTES);
  C *sout  = reinterpret_cast<C *>(base + CF::KY_BYTES + CF::XW_BYTES);

  const int g = lane / CF::W, la = lane % CF::W;
  const SweepItem item = a.pts.items[it];
  const int i2  = (int)item.row;
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1];
  const int y0  = wrap_index(kBinY * i2 - CF::HL, nf2);  // fine-grid row of window row 0

  C acc[CF::YR];
#pragma unroll
  for (int r = 0; r < CF::YR; ++r) acc[r] = C{0, 0};
  constexpr int NONE = INT_MIN;
  int jw = NONE;  // first column of the window; NONE = window empty

  // spread: add this lane's column to the fine grid and clear it
  auto flush_col = [&](int col) {
    const uint32_t gx = (uint32_t)wrap_index(col, nf1);
#pragma unroll
    for (int r = 0; r < CF::YR; ++r) {
      const int gy = wrap_index(y0 + r, nf2);
      const C v = acc[r];
      if (v.x != (T)0 || v.y != (T)0)
        atomic_add_cx(a.fw + ((uint32_t)gy * (uint32_t)nf1 + gx), v);
      acc[r] = C{0, 0};
    }
  };
  // interp: load this lane's column from the fine grid
  auto load_col = [&](int col) {
    const uint32_t gx = (uint32_t)wrap_index(col, nf1);
#pragma unroll
    for (int r = 0; r < CF::YR; ++r) {
      const int gy = wrap_index(y0 + r, nf2);
      acc[r] = __ldg(a.fw + ((uint32_t)gy * (uint32_t)nf1 + gx));
    }
  };
  // the window moves by S columns: [jw, jw+W) -> [jw+S, jw+S+W)
  auto slide = [&]() {
    const int rel = (la - jw) & (CF::W - 1);  // this lane's column is jw + rel
    if (rel < CF::S) {
      if (SPREAD) flush_col(jw + rel);
      else load_col(jw + CF::W + rel);
    }
    jw += CF::S;
  };
  auto empty_window = [&]() {
    if (jw == NONE) return;
    if (SPREAD) flush_col(jw + ((la - jw) & (CF::W - 1)));
    jw = NONE;
  };
  auto advance_to = [&](int x) {
    if (x == jw) return;
    if (jw != NONE && (x < jw || x - jw >= CF::W)) empty_window();
    if (jw == NONE) jw = SPREAD ? x : x - CF::W;  // interp: slide the real columns in
    while (jw < x) slide();
  };

  // ---- software pipeline over chunks of 32 points: raw data two chunks ahead, strength one
  struct Raw {
    T x, y;
    uint32_t j;
  };
  auto load_raw = [&](uint32_t q) {
    Raw r{(T)0, (T)0, 0u};
    if (q < item.qb) {
      r.x = __ldcs(a.pts.xs + q), r.y = __ldcs(a.pts.ys + q);
      r.j = __ldcs(a.pts.sidx + q);
    }
    return r;
  };
  auto load_c = [&](uint32_t q, const Raw &r) {
    C c{0, 0};
    if (SPREAD && q < item.qb) c = __ldcs(a.c_in + r.j);
    return c;
  };
  Raw r1 = load_raw(item.qa + lane);
  Raw r2 = load_raw(item.qa + 32 + lane);
  C c1   = load_c(item.qa + lane, r1);

  for (uint32_t q0 = item.qa; q0 < item.qb; q0 += CF::CH) {
    const int nc  = (int)min((uint32_t)CF::CH, item.qb - q0);
    const Raw cur = r1;
    const C ccur  = c1;
    r1            = r2;
    c1            = load_c(q0 + 32 + lane, r1);
    r2            = load_raw(q0 + 64 + lane);

    // ---- thread-per-point preparation
    int key = INT_MIN;
    if (lane < nc) {
      int i0;
      T x1;
      T kv[CF::KP + 2];
      stencil_start<T, NS>(fold_rescale<T>(cur.x, a.g.nf_t[0]), i0, x1);
      eval_window_t(a.tab, x1, kv);
      const int gpos = (i0 + CF::XB) >> (CF::S - 1);
      // rotate: the weight of stencil cell t belongs to the window column x = i0 + t (mod W)
#pragma unroll
      for (int t = 0; t < CF::W; ++t) {
        const T w = t < NS ? kv[t < NS ? t : 0] : (T)0;
        XE *dst   = sxw + ((i0 + t) & (CF::W - 1)) * CF::XP + lane;
        if constexpr (SPREAD) *dst = cx_mul(w, ccur);
        else *dst = w;
      }
      int j0;
      stencil_start<T, NS>(fold_rescale<T>(cur.y, a.g.nf_t[1]), j0, x1);
      eval_window_t(a.tab, x1, kv);
#pragma unroll
      for (int t = NS; t < CF::KP; ++t) kv[t] = (T)0;
      T *rk = sky + lane * CF::KP;
#pragma unroll
      for (int v = 0; v < CF::KP / CF::KV; ++v) {
        if constexpr (sizeof(T) == 4)
          *reinterpret_cast<float4 *>(rk + 4 * v) =
              make_float4(kv[4 * v], kv[4 * v + 1], kv[4 * v + 2], kv[4 * v + 3]);
        else
          *reinterpret_cast<double2 *>(rk + 2 * v) = make_double2(kv[2 * v], kv[2 * v + 1]);
      }
      const int jb = min(max(j0 - (kBinY * i2 - CF::HL), 0), kBinY);
      key          = gpos * 8 + jb;
    }
    // runs of equal key: bit l of heads = point l starts a run
    const int prev       = __shfl_up_sync(0xffffffffu, key, 1);
    const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    __syncwarp();

    int p = 0;
    while (p < nc) {
      const int kp        = __shfl_sync(0xffffffffu, key, p);
      const uint32_t rest = p < 31 ? heads >> (p + 1) : 0u;
      int e               = rest ? p + __ffs(rest) : 32;
      e                   = min(e, nc);
      advance_to(CF::S * (kp >> 3) - CF::XB);
#define B200_RUN2(J)                                                              \
  case J:                                                                         \
    if constexpr (SPREAD) spread_run2<T, NS, J, CF>(acc, sky, sxw, p, e, g, la);  \
    else interp_run2<T, NS, J, CF>(acc, sky, sxw, sout, p, e, g, la);             \
    break;
      switch (kp & 7) {
        B200_RUN2(0) B200_RUN2(1) B200_RUN2(2) B200_RUN2(3) B200_RUN2(4)
      default: break;
      }
#undef B200_RUN2
      p = e;
    }
    __syncwarp();
    if (!SPREAD && lane < nc) a.c_out[cur.j] = sout[lane];
    __syncwarp();
  }
  empty_window();
}

template<class T, int NS, bool SPREAD>
static cudaError_t launch_sweep2_ns(const Sweep2Points<T> &pts, const GridGeom<T> &g, int nc,
                                    const T *coef, const typename CxOf<T>::type *c_in,
                                    typename CxOf<T>::type *c_out, typename CxOf<T>::type *fw,
                                    cudaStream_t st) {
  using CF = Sweep2Cfg<T, NS, SPREAD>;
  if (pts.nitems == 0) return cudaSuccess;
  Sweep2Args<T, NS> a;
  a.pts = pts;
  a.g   = g;
  fill_table(a.tab, nc, coef);
  a.c_in  = c_in;
  a.c_out = c_out;
  a.fw    = fw;
  const size_t shbytes = CF::WPB * CF::WARP_BYTES;
  auto kern            = k_sweep2<T, NS, SPREAD>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)shbytes);
  if (e != cudaSuccess) return e;
  kern<<<(pts.nitems + CF::WPB - 1) / CF::WPB, CF::WPB * 32, shbytes, st>>>(a);
  return cudaGetLastError();
}

#define B200_SWEEP2_CASE(NSV)                                                                  \
  case NSV:                                                                                    \
    return spread ? launch_sweep2_ns<T, NSV, true>(pts, g, nc, coef, c_in, c_out, fw, st)      \
                  : launch_sweep2_ns<T, NSV, false>(pts, g, nc, coef, c_in, c_out, fw, st);

}  // namespace b200
