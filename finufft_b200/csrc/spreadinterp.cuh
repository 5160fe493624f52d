// Type-1 spreading and type-2 interpolation kernels ("warp per bin" design).
//
// What they compute (reference CPU semantics):
//   spread  fw[n] += sum_j c_j phi1(X_j-n1) phi2(Y_j-n2) phi3(Z_j-n3)   on the periodic fine grid,
//           include/finufft/spreadinterp.hpp:310-485, include/finufft/spread.hpp:53-451
//   interp  c_j = sum_n fw[n] phi1 phi2 phi3,  include/finufft/interp.hpp:11-556
//   stencil start ceil(X - ns/2), ns^d cells, periodic wrap (spread.hpp:328-336).
//
// How (B200): points arrive bin-sorted (16x4x4-cell bins, sort.cuh).  One warp owns one
// subproblem = up to `maxsub` consecutive points of one bin, and a private padded tile of the
// fine grid in shared memory, (16+ns) x (4+ns) x (4+ns) cells, which contains every stencil
// of the bin.  Because the warp is the only writer of its tile, accumulation is plain
// load-fma-store on shared memory: no shared-memory atomics.  Lanes split each point's
// stencil so that no two lanes touch the same cell, and the tile pitches are chosen so the
// lanes of one access hit distinct banks:
//   3D: lane <-> (dy,dz) row, ns cells along x each;   2D/1D: lane <-> one cell.
// The window values of 32 points are evaluated thread-per-point (full unrolled Horner with
// the table in the constant bank) and parked in shared memory, then the warp walks the 32
// points.  The tile is flushed once with vector atomics (type 1) or filled once with
// coalesced row loads (type 2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "devmath.cuh"
#include "sort.cuh"

namespace b200 {

template<class T> struct PointSet {
  const T *xs, *ys, *zs;        // coordinates in sorted order
  const uint32_t *sidx;         // sorted position -> user index
  const uint32_t *binstart;     // nbins+1
  const uint32_t *sub_bin, *sub_off;
  uint32_t nsub, maxsub;
};

template<class T, int DIM, int NS> struct Tile {
  static constexpr int PX = kBinX + NS;
  static constexpr int PY = DIM > 1 ? kBinY + NS : 1;
  static constexpr int PZ = DIM > 2 ? kBinZ + NS : 1;
  // 3D: lanes enumerate (dy,dz) rows -> want row pitch L odd and plane pitch P = L*NS (mod 16)
  //     so that lane l lands in bank group (L*l) mod 16.  2D/1D: lanes enumerate cells, pitch
  //     PX = 16+NS = NS (mod 16) already makes consecutive cells consecutive mod 16.
  static constexpr int L = DIM == 3 ? (PX | 1) : PX;
  static constexpr int plane_min = L * PY;
  static constexpr int want = (L * NS) % 16;
  static constexpr int P = DIM == 3 ? plane_min + ((want - plane_min % 16) + 16) % 16 : plane_min;
  static constexpr int CELLS = P * PZ;
  static constexpr int KV = (DIM * NS) | 1;  // odd pitch of the per-point window values
  static constexpr size_t BYTES =
      ((CELLS * 2 * sizeof(T) + 32 * KV * sizeof(T)) + 15) / 16 * 16;
  static constexpr int ITEMS = DIM == 3 ? NS * NS : (DIM == 2 ? NS * NS : NS);
  static constexpr int NIT   = (ITEMS + 31) / 32;
};

template<class T, int NS> struct SpreadArgs {
  PointSet<T> pts;
  GridGeom<T> g;
  WindowTable<T, NS> tab;
  const typename CxOf<T>::type *c_in;   // spread: strengths (user order)
  typename CxOf<T>::type *c_out;        // interp: outputs (user order)
  typename CxOf<T>::type *fw;           // fine grid
};

__device__ __forceinline__ void atomic_add_cx(float2 *p, float2 v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_cx(double2 *p, double2 v) {
  atomicAdd(&p->x, v.x);
  atomicAdd(&p->y, v.y);
}

// Phase A for one lane's point: fold, stencil start, window values -> shared memory.
template<class T, int DIM, int NS>
__device__ __forceinline__ int prepare_point(const SpreadArgs<T, NS> &a, uint32_t q, T *kv,
                                             int org1, int org2, int org3) {
  using TL = Tile<T, DIM, NS>;
  int i0;
  T x1;
  stencil_start<T, NS>(fold_rescale<T>(a.pts.xs[q], a.g.nf_t[0]), i0, x1);
  if (DIM == 1) {  // spread.hpp:116-121
    x1 = x1 < (T)(-0.5 * NS) ? (T)(-0.5 * NS) : x1;
    x1 = x1 > (T)(-0.5 * NS + 1) ? (T)(-0.5 * NS + 1) : x1;
  }
  eval_window<T, NS>(a.tab, x1, kv);
  int off = min(max(i0 - org1, 0), kBinX);
  if (DIM > 1) {
    stencil_start<T, NS>(fold_rescale<T>(a.pts.ys[q], a.g.nf_t[1]), i0, x1);
    eval_window<T, NS>(a.tab, x1, kv + NS);
    off += TL::L * min(max(i0 - org2, 0), kBinY);
  }
  if (DIM > 2) {
    stencil_start<T, NS>(fold_rescale<T>(a.pts.zs[q], a.g.nf_t[2]), i0, x1);
    eval_window<T, NS>(a.tab, x1, kv + 2 * NS);
    off += TL::P * min(max(i0 - org3, 0), kBinZ);
  }
  return off;
}

template<class T, int DIM, int NS>
__global__ void __launch_bounds__(256) k_spread(const SpreadArgs<T, NS> a) {
  using TL = Tile<T, DIM, NS>;
  using C  = typename CxOf<T>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t s = blockIdx.x * (blockDim.x >> 5) + warp;
  if (s >= a.pts.nsub) return;
  C *tile = reinterpret_cast<C *>(smem + warp * TL::BYTES);
  T *kvs  = reinterpret_cast<T *>(tile + TL::CELLS);

  const uint32_t bin = a.pts.sub_bin[s], p0 = a.pts.sub_off[s];
  const uint32_t pend = min(a.pts.binstart[bin + 1], p0 + a.pts.maxsub);
  const int b1 = bin % a.g.nb[0], b2 = (bin / a.g.nb[0]) % a.g.nb[1],
            b3 = (bin / (a.g.nb[0] * a.g.nb[1])) % a.g.nb[2];  // bins are group-major
  const int org1 = kBinX * b1 - NS / 2, org2 = kBinY * b2 - NS / 2, org3 = kBinZ * b3 - NS / 2;

  for (int i = lane; i < TL::CELLS; i += 32) tile[i] = C{0, 0};

  // per-lane stencil items
  int item_off[TL::NIT], item_ka[TL::NIT], item_kb[TL::NIT];
#pragma unroll
  for (int it = 0; it < TL::NIT; ++it) {
    const int item = lane + 32 * it;
    if (DIM == 3) {
      item_off[it] = (item % NS) * TL::L + (item / NS) * TL::P;
      item_ka[it]  = NS + item % NS;
      item_kb[it]  = 2 * NS + item / NS;
    } else if (DIM == 2) {
      item_off[it] = (item % NS) + (item / NS) * TL::L;
      item_ka[it]  = item % NS;
      item_kb[it]  = NS + item / NS;
    } else {
      item_off[it] = item;
      item_ka[it]  = item;
      item_kb[it]  = 0;
    }
  }
  __syncwarp();

  for (uint32_t q0 = p0; q0 < pend; q0 += 32) {
    const int cnt = (int)min(32u, pend - q0);
    int myoff = 0;
    T cre = 0, cim = 0;
    if (lane < cnt) {
      myoff       = prepare_point<T, DIM, NS>(a, q0 + lane, kvs + lane * TL::KV, org1, org2, org3);
      const C cc  = a.c_in[a.pts.sidx[q0 + lane]];
      cre         = cc.x;
      cim         = cc.y;
    }
    __syncwarp();
    for (int j = 0; j < cnt; ++j) {
      const int off = __shfl_sync(0xffffffffu, myoff, j);
      const T re = __shfl_sync(0xffffffffu, cre, j), im = __shfl_sync(0xffffffffu, cim, j);
      const T *kj = kvs + j * TL::KV;
      if (DIM == 3) {
        T k1[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) k1[t] = kj[t];
#pragma unroll
        for (int it = 0; it < TL::NIT; ++it) {
          if (lane + 32 * it < TL::ITEMS) {
            const T w  = kj[item_ka[it]] * kj[item_kb[it]];
            const T wr = w * re, wi = w * im;
            C *cell = tile + off + item_off[it];
#pragma unroll
            for (int t = 0; t < NS; ++t) {
              C v     = cell[t];
              v.x     = fma_rn(k1[t], wr, v.x);
              v.y     = fma_rn(k1[t], wi, v.y);
              cell[t] = v;
            }
          }
        }
      } else {
#pragma unroll
        for (int it = 0; it < TL::NIT; ++it) {
          if (lane + 32 * it < TL::ITEMS) {
            const T w = DIM == 2 ? kj[item_ka[it]] * kj[item_kb[it]] : kj[item_ka[it]];
            C *cell   = tile + off + item_off[it];
            C v       = *cell;
            v.x       = fma_rn(w, re, v.x);
            v.y       = fma_rn(w, im, v.y);
            *cell     = v;
          }
        }
      }
      __syncwarp();
    }
  }

  // flush: every tile cell to its periodic image with one vector atomic
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1], nf3 = a.g.nf[2];
  for (int r = 0; r < TL::PY * TL::PZ; ++r) {
    const int oy = r % TL::PY, oz = r / TL::PY;
    const int gy = DIM > 1 ? wrap_index(org2 + oy, nf2) : 0;
    const int gz = DIM > 2 ? grid_plane(a.g, wrap_index(org3 + oz, nf3)) : 0;
    if (gz < 0) continue;  // outside the z window: no point of this plan reaches it
    C *row        = a.fw + ((size_t)gz * nf2 + gy) * (size_t)nf1;
    const C *trow = tile + oy * TL::L + oz * TL::P;
    for (int ox = lane; ox < TL::PX; ox += 32) {
      const C v = trow[ox];
      if (v.x != (T)0 || v.y != (T)0) atomic_add_cx(row + wrap_index(org1 + ox, nf1), v);
    }
  }
}

template<class T, int DIM, int NS>
__global__ void __launch_bounds__(256) k_interp(const SpreadArgs<T, NS> a) {
  using TL = Tile<T, DIM, NS>;
  using C  = typename CxOf<T>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t s = blockIdx.x * (blockDim.x >> 5) + warp;
  if (s >= a.pts.nsub) return;
  C *tile = reinterpret_cast<C *>(smem + warp * TL::BYTES);
  T *kvs  = reinterpret_cast<T *>(tile + TL::CELLS);

  const uint32_t bin = a.pts.sub_bin[s], p0 = a.pts.sub_off[s];
  const uint32_t pend = min(a.pts.binstart[bin + 1], p0 + a.pts.maxsub);
  const int b1 = bin % a.g.nb[0], b2 = (bin / a.g.nb[0]) % a.g.nb[1],
            b3 = (bin / (a.g.nb[0] * a.g.nb[1])) % a.g.nb[2];  // bins are group-major
  const int org1 = kBinX * b1 - NS / 2, org2 = kBinY * b2 - NS / 2, org3 = kBinZ * b3 - NS / 2;

  // fill the tile from the fine grid (periodic)
  const int nf1 = a.g.nf[0], nf2 = a.g.nf[1], nf3 = a.g.nf[2];
  for (int r = 0; r < TL::PY * TL::PZ; ++r) {
    const int oy = r % TL::PY, oz = r / TL::PY;
    const int gy = DIM > 1 ? wrap_index(org2 + oy, nf2) : 0;
    const int gz = DIM > 2 ? grid_plane(a.g, wrap_index(org3 + oz, nf3)) : 0;
    C *trow      = tile + oy * TL::L + oz * TL::P;
    if (gz < 0) {  // outside the z window
      for (int ox = lane; ox < TL::PX; ox += 32) trow[ox] = C{0, 0};
      continue;
    }
    const C *row = a.fw + ((size_t)gz * nf2 + gy) * (size_t)nf1;
    for (int ox = lane; ox < TL::PX; ox += 32) trow[ox] = row[wrap_index(org1 + ox, nf1)];
  }

  int item_off[TL::NIT], item_ka[TL::NIT], item_kb[TL::NIT];
#pragma unroll
  for (int it = 0; it < TL::NIT; ++it) {
    const int item = lane + 32 * it;
    if (DIM == 3) {
      item_off[it] = (item % NS) * TL::L + (item / NS) * TL::P;
      item_ka[it]  = NS + item % NS;
      item_kb[it]  = 2 * NS + item / NS;
    } else if (DIM == 2) {
      item_off[it] = (item % NS) + (item / NS) * TL::L;
      item_ka[it]  = item % NS;
      item_kb[it]  = NS + item / NS;
    } else {
      item_off[it] = item;
      item_ka[it]  = item;
      item_kb[it]  = 0;
    }
  }
  __syncwarp();

  for (uint32_t q0 = p0; q0 < pend; q0 += 32) {
    const int cnt = (int)min(32u, pend - q0);
    int myoff = 0;
    if (lane < cnt)
      myoff = prepare_point<T, DIM, NS>(a, q0 + lane, kvs + lane * TL::KV, org1, org2, org3);
    __syncwarp();
    T out_re = 0, out_im = 0;
    for (int j = 0; j < cnt; ++j) {
      const int off = __shfl_sync(0xffffffffu, myoff, j);
      const T *kj   = kvs + j * TL::KV;
      T ar = 0, ai = 0;
      if (DIM == 3) {
        T k1[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) k1[t] = kj[t];
#pragma unroll
        for (int it = 0; it < TL::NIT; ++it) {
          if (lane + 32 * it < TL::ITEMS) {
            const T w     = kj[item_ka[it]] * kj[item_kb[it]];
            const C *cell = tile + off + item_off[it];
            T sr = 0, si = 0;
#pragma unroll
            for (int t = 0; t < NS; ++t) {
              const C v = cell[t];
              sr        = fma_rn(v.x, k1[t], sr);
              si        = fma_rn(v.y, k1[t], si);
            }
            ar = fma_rn(sr, w, ar);
            ai = fma_rn(si, w, ai);
          }
        }
      } else {
#pragma unroll
        for (int it = 0; it < TL::NIT; ++it) {
          if (lane + 32 * it < TL::ITEMS) {
            const T w = DIM == 2 ? kj[item_ka[it]] * kj[item_kb[it]] : kj[item_ka[it]];
            const C v = tile[off + item_off[it]];
            ar        = fma_rn(v.x, w, ar);
            ai        = fma_rn(v.y, w, ai);
          }
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, d);
        ai += __shfl_xor_sync(0xffffffffu, ai, d);
      }
      if (lane == j) {
        out_re = ar;
        out_im = ai;
      }
    }
    if (lane < cnt) a.c_out[a.pts.sidx[q0 + lane]] = C{out_re, out_im};
    __syncwarp();
  }
}

// --------------------------------------------------------------------------- host launch
// Choose warps per block to maximise resident warps per SM within the shared-memory budget.
inline int pick_warps_per_block(size_t warp_bytes) {
  const size_t sm_budget = 227 * 1024, per_block_cap = 227 * 1024;
  int best = 1, best_res = 0;
  for (int w = 1; w <= 8; ++w) {
    const size_t blk = w * warp_bytes;
    if (blk > per_block_cap) break;
    int blocks = (int)(sm_budget / (blk + 1024));
    if (blocks > 32) blocks = 32;
    int res = blocks * w;
    if (res > 64) res = 64;
    if (res > best_res) {
      best_res = res;
      best     = w;
    }
  }
  return best;
}

template<class T, int DIM, int NS>
cudaError_t launch_spreadinterp_ns(bool spread, const PointSet<T> &pts, const GridGeom<T> &g,
                                   int nc, const T *coef, const typename CxOf<T>::type *c_in,
                                   typename CxOf<T>::type *c_out, typename CxOf<T>::type *fw,
                                   cudaStream_t st) {
  using TL = Tile<T, DIM, NS>;
  SpreadArgs<T, NS> a;
  a.pts = pts;
  a.g   = g;
  constexpr int rows = TableRows<NS>::value;
  for (int k = 0; k < rows; ++k)
    for (int j = 0; j < NS; ++j) {
      const int src   = k - (rows - nc);
      a.tab.c[k * NS + j] = src >= 0 ? coef[src * NS + j] : (T)0;
    }
  a.c_in  = c_in;
  a.c_out = c_out;
  a.fw    = fw;
  if (TL::BYTES > 227 * 1024) return cudaErrorInvalidConfiguration;
  const int wpb       = pick_warps_per_block(TL::BYTES);
  const size_t shbytes = wpb * TL::BYTES;
  auto kern = spread ? k_spread<T, DIM, NS> : k_interp<T, DIM, NS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)shbytes);
  if (e != cudaSuccess) return e;
  const uint32_t nblk = (pts.nsub + wpb - 1) / wpb;
  kern<<<nblk, wpb * 32, shbytes, st>>>(a);
  return cudaGetLastError();
}

// runtime (dim handled by the per-dim translation unit, ns by this switch)
template<class T, int DIM>
cudaError_t launch_spreadinterp(bool spread, int ns, const PointSet<T> &pts,
                                const GridGeom<T> &g, int nc, const T *coef,
                                const typename CxOf<T>::type *c_in,
                                typename CxOf<T>::type *c_out, typename CxOf<T>::type *fw,
                                cudaStream_t st);
#define B200_DECLARE_LAUNCH(T_, DIM_)                                                         \
  template<>                                                                                  \
  cudaError_t launch_spreadinterp<T_, DIM_>(                                                  \
      bool spread, int ns, const PointSet<T_> &pts, const GridGeom<T_> &g, int nc,            \
      const T_ *coef, const CxOf<T_>::type *c_in, CxOf<T_>::type *c_out, CxOf<T_>::type *fw,  \
      cudaStream_t st);
B200_DECLARE_LAUNCH(float, 1)
B200_DECLARE_LAUNCH(float, 2)
B200_DECLARE_LAUNCH(float, 3)
B200_DECLARE_LAUNCH(double, 1)
B200_DECLARE_LAUNCH(double, 2)
B200_DECLARE_LAUNCH(double, 3)

#define B200_NS_CASE(NSV)                                                                     \
  case NSV:                                                                                   \
    return launch_spreadinterp_ns<T, DIM, NSV>(spread, pts, g, nc, coef, c_in, c_out, fw, st);

#define B200_DEFINE_LAUNCH(T_, DIM_)                                                          \
  template<>                                                                                  \
  cudaError_t launch_spreadinterp<T_, DIM_>(                                                  \
      bool spread, int ns, const PointSet<T_> &pts, const GridGeom<T_> &g, int nc,            \
      const T_ *coef, const CxOf<T_>::type *c_in, CxOf<T_>::type *c_out, CxOf<T_>::type *fw,  \
      cudaStream_t st) {                                                                      \
    using T = T_;                                                                             \
    constexpr int DIM = DIM_;                                                                 \
    switch (ns) {                                                                             \
      B200_NS_LIST                                                                            \
    default:                                                                                  \
      return cudaErrorInvalidValue;                                                           \
    }                                                                                         \
  }

}  // namespace b200
