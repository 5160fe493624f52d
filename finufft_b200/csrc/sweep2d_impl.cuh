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
  // y-window record of a point: its NS values placed at the rows of the register window they
  // multiply (offset jb = stencil start inside the bin row), zeros elsewhere
  static constexpr int KP  = (YR + KV - 1) / KV * KV;
  // x-window record [slot][point]: pitch chosen so that the lanes of one shared-memory phase
  // (128 bytes) read distinct banks: lane (g, a) reads slot a of point q+g
  static constexpr int LPP = 128 / (int)sizeof(XE);
  static constexpr int XP  = CH + (LPP / W > 1 ? LPP / W : 1);
  static constexpr size_t KY_BYTES  = (size_t)CH * KP * sizeof(T);
  static constexpr size_t XW_BYTES  = ((size_t)W * XP * sizeof(XE) + 15) / 16 * 16;
  // interp: the W per-column shares of every point, summed thread-per-point after the chunk
  static constexpr size_t OUT_BYTES = SPREAD ? 0 : (size_t)CH * W * sizeof(C);
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

// the padded y record of one point (16-byte aligned)
template<class T, int KP> __device__ __forceinline__ void load_ky(const T *rec, T (&ky)[KP]) {
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

// One step: lane group g handles point idx of the chunk.  The two-point forms issue the
// shared-memory loads of both points before any arithmetic.
template<class T, class CF>
__device__ __forceinline__ void spread_math2(typename CF::C (&acc)[CF::YR], const T (&ky)[CF::KP],
                                             typename CF::C cw) {
#pragma unroll
  for (int r = 0; r < CF::YR; ++r) acc[r] = cx_fma(ky[r], cw, acc[r]);
}
template<class T, class CF>
__device__ __forceinline__ void spread_step2(typename CF::C (&acc)[CF::YR], const T *sky,
                                             const typename CF::C *sxw, int idx, int la) {
  T ky[CF::KP];
  load_ky<T, CF::KP>(sky + idx * CF::KP, ky);
  spread_math2<T, CF>(acc, ky, sxw[la * CF::XP + idx]);
}
template<class T, class CF>
__device__ __forceinline__ void spread_step2x2(typename CF::C (&acc)[CF::YR], const T *sky,
                                               const typename CF::C *sxw, int idx, int la) {
  T ka[CF::KP], kb[CF::KP];
  load_ky<T, CF::KP>(sky + idx * CF::KP, ka);
  load_ky<T, CF::KP>(sky + (idx + CF::G) * CF::KP, kb);
  const typename CF::C ca = sxw[la * CF::XP + idx], cb = sxw[la * CF::XP + idx + CF::G];
  spread_math2<T, CF>(acc, ka, ca);
  spread_math2<T, CF>(acc, kb, cb);
}
template<class T, class CF>
__device__ __forceinline__ typename CF::C interp_math2(const typename CF::C (&acc)[CF::YR],
                                                       const T (&ky)[CF::KP], T wx) {
  // two independent chains halve the dependent-FMA latency
  typename CF::C v0 = cx_mul(ky[0], acc[0]), v1 = cx_mul(ky[1], acc[1]);
#pragma unroll
  for (int r = 2; r + 1 < CF::YR; r += 2) {
    v0 = cx_fma(ky[r], acc[r], v0);
    v1 = cx_fma(ky[r + 1], acc[r + 1], v1);
  }
  if constexpr (CF::YR % 2 == 1) v0 = cx_fma(ky[CF::YR - 1], acc[CF::YR - 1], v0);
  v0.x += v1.x, v0.y += v1.y;
  return cx_mul(wx, v0);
}
template<class T, class CF>
__device__ __forceinline__ void interp_step2(const typename CF::C (&acc)[CF::YR], const T *sky,
                                             const T *sxw, typename CF::C *spart, int idx,
                                             int la) {
  T ky[CF::KP];
  load_ky<T, CF::KP>(sky + idx * CF::KP, ky);
  spart[idx * CF::W + la] = interp_math2<T, CF>(acc, ky, sxw[la * CF::XP + idx]);
}
template<class T, class CF>
__device__ __forceinline__ void interp_step2x2(const typename CF::C (&acc)[CF::YR], const T *sky,
                                               const T *sxw, typename CF::C *spart, int idx,
                                               int la) {
  T ka[CF::KP], kb[CF::KP];
  load_ky<T, CF::KP>(sky + idx * CF::KP, ka);
  load_ky<T, CF::KP>(sky + (idx + CF::G) * CF::KP, kb);
  const T wa = sxw[la * CF::XP + idx], wb = sxw[la * CF::XP + idx + CF::G];
  const typename CF::C va = interp_math2<T, CF>(acc, ka, wa);
  const typename CF::C vb = interp_math2<T, CF>(acc, kb, wb);
  spart[idx * CF::W + la]           = va;
  spart[(idx + CF::G) * CF::W + la] = vb;
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
  C *spart = reinterpret_cast<C *>(base + CF::KY_BYTES + CF::XW_BYTES);

  const int g = lane / CF::W, la = lane % CF::W;
  const SweepItem item = a.pts.items[it];
  const int i2  = (int)(item.row % (uint32_t)a.g.nb[1]);  // rows are group-major (sort.cuh)
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1];
  const int y0  = wrap_index(kBinY * i2 - CF::HL, nf2);  // fine-grid row of window row 0

  C acc[CF::YR];
#pragma unroll
  for (int r = 0; r < CF::YR; ++r) acc[r] = C{0, 0};
  constexpr int NONE = INT_MIN;
  int jw = NONE;  // first column of the window; NONE = window empty

  // spread: add this lane's column to the fine grid and clear it.  Every lane group flushes its
  // own copy (summing the copies with shuffles first was measured slower: the reductions hit
  // the L2-resident grid and are cheaper than 4*YR shuffles per window move), zeros included:
  // testing for them costs a branch per row.
  auto flush_col = [&](int col) {
    const uint32_t gx = (uint32_t)wrap_index(col, nf1);
#pragma unroll
    for (int r = 0; r < CF::YR; ++r) {
      const int gy = wrap_index(y0 + r, nf2);
      atomic_add_cx(a.fw + ((uint32_t)gy * (uint32_t)nf1 + gx), acc[r]);
      acc[r] = C{0, 0};
    }
  };
  // interp: columns enter from the fine grid.  The owners of the columns that enter at the NEXT
  // window move load them one move ahead into nxt[] (tag nxt_col), so the latency of the
  // fine-grid reads is covered by a whole window position of work.
  // (single precision only: in double the extra window column costs too many registers)
  constexpr bool PREFETCH = !SPREAD && sizeof(T) == 4;
  C nxt[PREFETCH ? CF::YR : 1];
  int nxt_col = NONE;
  auto fetch_col = [&](int col, auto &dst) {
    if constexpr (!SPREAD) {
      const uint32_t gx = (uint32_t)wrap_index(col, nf1);
#pragma unroll
      for (int r = 0; r < CF::YR; ++r) {
        const int gy = wrap_index(y0 + r, nf2);
        dst[r]       = __ldg(a.fw + ((uint32_t)gy * (uint32_t)nf1 + gx));
      }
    }
  };
  // the window moves by S columns: [jw, jw+W) -> [jw+S, jw+S+W)
  auto slide = [&]() {
    const int rel = (la - jw) & (CF::W - 1);  // this lane's column is jw + rel
    if constexpr (SPREAD) {
      if (rel < CF::S) flush_col(jw + rel);
      jw += CF::S;
    } else {
      if (rel < CF::S) {
        const int col = jw + CF::W + rel;
        bool have = false;
        if constexpr (PREFETCH) {
          if (nxt_col == col) {
            have = true;
#pragma unroll
            for (int r = 0; r < CF::YR; ++r) acc[r] = nxt[r];
          }
        }
        if (!have) {
          C tmp[CF::YR];
          fetch_col(col, tmp);
#pragma unroll
          for (int r = 0; r < CF::YR; ++r) acc[r] = tmp[r];
        }
      }
      jw += CF::S;
      if constexpr (PREFETCH) {
        const int rn = (la - jw) & (CF::W - 1);
        if (rn < CF::S) {
          nxt_col = jw + CF::W + rn;
          fetch_col(nxt_col, nxt);
        }
      }
    }
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
    if (SPREAD && q < item.qb) c = __ldg(a.c_in + r.j);
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
      T kv[NS + 2];
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
      const int jb = min(max(j0 - (kBinY * i2 - CF::HL), 0), kBinY);
      T *rk        = sky + lane * CF::KP;
#pragma unroll
      for (int v = 0; v < CF::KP / CF::KV; ++v) {
        if constexpr (sizeof(T) == 4)
          *reinterpret_cast<float4 *>(rk + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
        else
          *reinterpret_cast<double2 *>(rk + 2 * v) = make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int t = 0; t < NS; ++t) rk[jb + t] = kv[t];
      key = gpos;
    }
    // runs of equal window position: bit l of heads = point l starts a run
    const int prev       = __shfl_up_sync(0xffffffffu, key, 1);
    const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    __syncwarp();

    int p = 0;
    while (p < nc) {
      const int kp        = __shfl_sync(0xffffffffu, key, p);
      const uint32_t rest = p < 31 ? heads >> (p + 1) : 0u;
      int e               = rest ? p + __ffs(rest) : 32;
      e                   = min(e, nc);
      advance_to(CF::S * kp - CF::XB);
      int q = p;
      for (; q + 2 * CF::G <= e; q += 2 * CF::G) {  // two full steps: loads of both first
        if constexpr (SPREAD) spread_step2x2<T, CF>(acc, sky, sxw, q + g, la);
        else interp_step2x2<T, CF>(acc, sky, sxw, spart, q + g, la);
      }
      for (; q < e; q += CF::G) {
        if (q + g < e) {
          if constexpr (SPREAD) spread_step2<T, CF>(acc, sky, sxw, q + g, la);
          else interp_step2<T, CF>(acc, sky, sxw, spart, q + g, la);
        }
      }
      p = e;
    }
    __syncwarp();
    if (!SPREAD && lane < nc) {  // lane = point: add the W column shares, scatter
      const C *pp = spart + lane * CF::W;
      C s0 = pp[0];
#pragma unroll
      for (int k = 1; k < CF::W; ++k) s0.x += pp[k].x, s0.y += pp[k].y;
      a.c_out[cur.j] = s0;
    }
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
