// setpts as a multi-level partition with sequential streams: see partition.cuh.
#include "partition.cuh"

#include <algorithm>

#include "devmath.cuh"
#include "scratch.hpp"
#include "sweep2d.cuh"

namespace b200 {

// ------------------------------------------------------------------------------ geometry
int part_class_count(int cls, int ns) {
  if (cls == kClassSweep3) return (kBinX / 2 + 1) * (kBinY + 1);
  if (cls == kClassSweep2) {
    const int W = ns <= 8 ? 8 : 16, S = ns + 1 <= W ? 2 : 1;
    return (kBinX / S + 2) * (kBinY + 1);
  }
  return 1;
}

PartPlan plan_partition(uint64_t M, uint32_t nbins, int cls, int ns, bool is_double, bool force) {
  PartPlan pp;
  pp.cls  = cls;
  pp.ns   = ns;
  pp.ncls = part_class_count(cls, ns);
  if (M == 0 || nbins == 0) return pp;
  const uint32_t cap = is_double ? SegCap<double>::value : SegCap<float>::value;
  const double target = 0.72 * cap;  // Poisson fluctuations of a uniform point set stay below cap
  const double rho    = (double)M / (double)nbins;
  int ssmax = 0;
  while (ssmax < 10 && (2u << ssmax) * (uint32_t)pp.ncls <= (uint32_t)kPartFanout) ++ssmax;
  int ss = 0;
  while (ss < ssmax && rho * (double)(2u << ss) <= target) ++ss;
  pp.ss   = ss;
  pp.nseg = (uint32_t)(((uint64_t)nbins + (1u << ss) - 1) >> ss);
  pp.sb   = 0;
  while (((uint64_t)(pp.nseg - 1) >> pp.sb) + 1 > (uint64_t)kPartFanout) ++pp.sb;
  pp.levels = pp.sb > 0 ? 2 : 1;
  pp.nA     = (uint32_t)(((uint64_t)(pp.nseg - 1) >> pp.sb) + 1);
  if (pp.sb > 10) return pp;  // more than 2^20 segments: a third level would be needed
  if (!force) {
    if (rho > 0.8 * cap) return pp;            // a single bin already overflows a segment
    if (M < (1u << 19)) return pp;             // launch-bound regime: nothing to gain
    if ((uint64_t)pp.nseg * 64 > M) return pp;  // mostly empty segments (huge grid, few points)
  }
  pp.ok = true;
  return pp;
}

// ------------------------------------------------------------------------------ bin key
template<class T, int DIM>
__device__ __forceinline__ uint32_t bin_key_of(T x, T y, T z, uint32_t i, const GridGeom<T> &g) {
  // 1/binsize are powers of two, so the product is exact; conversion truncates (X >= 0)
  uint32_t key = (uint32_t)(int)mul_rn(fold_rescale<T>(x, g.nf_t[0]), (T)(1.0 / kBinX));
  if (DIM > 1)
    key += (uint32_t)g.nb[0] * (uint32_t)(int)mul_rn(fold_rescale<T>(y, g.nf_t[1]), (T)(1.0 / kBinY));
  if (DIM > 2)
    key += (uint32_t)g.nb[0] * (uint32_t)g.nb[1] *
           (uint32_t)(int)mul_rn(fold_rescale<T>(z, g.nf_t[2]), (T)(1.0 / kBinZ));
  key = key < g.nbins1 ? key : g.nbins1 - 1;  // only non-finite input can trip this
  if (g.nchunks > 1) key += point_group(g, i) * g.nbins1;  // group-major (sort.cuh)
  return key;
}

__device__ __forceinline__ float idx_as(float, uint32_t i) { return __uint_as_float(i); }
__device__ __forceinline__ double idx_as(double, uint32_t i) {
  return __longlong_as_double((long long)i);
}
__device__ __forceinline__ uint32_t idx_of(float w) { return __float_as_uint(w); }
__device__ __forceinline__ uint32_t idx_of(double w) { return (uint32_t)__double_as_longlong(w); }

static inline int blocks_for(uint64_t n, int threads, int per_sm) {
  const uint64_t want = (n + threads - 1) / threads, cap = 148ull * per_sm;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// ------------------------------------------------------------------------------ 1. histogram
template<class T, int DIM>
__global__ void __launch_bounds__(256)
k_bin_hist(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, uint32_t M,
           GridGeom<T> g, uint32_t *__restrict__ cnt) {
  const int lane        = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < M; i0 += stride) {
    const uint32_t i = i0 + lane;
    const bool valid = i < M;
    T px = 0, py = 0, pz = 0;
    if (valid) {
      px = __ldcs(x + i);
      if (DIM > 1) py = __ldcs(y + i);
      if (DIM > 2) pz = __ldcs(z + i);
    }
    const uint32_t key = valid ? bin_key_of<T, DIM>(px, py, pz, i, g) : 0xffffffffu;
    // Clustered input: one RED per distinct bin in the warp instead of up to 32 serialised on one
    // address (match_any).  Spread-out input never has two lanes in one bin, and match_any is
    // the most expensive instruction of the loop (short-scoreboard stalls, profiles/
    // r2g_setpts_ncu.txt), so it runs only when a cheap neighbour test sees a repeated bin.
#ifndef B200_HIST_TWIN
#define B200_HIST_TWIN 1
#endif
    // (both shuffles unconditionally: a short-circuited || would leave lanes out of the second)
    const uint32_t kn1 = __shfl_xor_sync(0xffffffffu, key, 1);
    const uint32_t kn2 = __shfl_xor_sync(0xffffffffu, key, 2);
    const bool twin = !B200_HIST_TWIN || key == kn1 || key == kn2;
    if (!__any_sync(0xffffffffu, twin && valid)) {
      if (valid) atomicAdd(&cnt[key], 1u);
    } else {
      const uint32_t peers = __match_any_sync(0xffffffffu, key);
      if (valid && lane == __ffs(peers) - 1) atomicAdd(&cnt[key], (uint32_t)__popc(peers));
    }
  }
}

// Every `step`-th point only: a cheap look at the distribution before the full passes.
template<class T, int DIM>
__global__ void __launch_bounds__(256)
k_bin_hist_sample(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                  uint32_t M, uint32_t step, GridGeom<T> g, uint32_t *__restrict__ cnt) {
  const uint32_t n      = M / step;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t i = k * step;
    atomicAdd(&cnt[bin_key_of<T, DIM>(x[i], DIM > 1 ? y[i] : (T)0, DIM > 2 ? z[i] : (T)0, i, g)], 1u);
  }
}
__global__ void __launch_bounds__(256)
k_max_u32(const uint32_t *__restrict__ v, uint32_t n, uint32_t *__restrict__ out) {
  uint32_t mx           = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) mx = max(mx, v[i]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

// ------------------------------------------------------------------------------ 2. segments
__global__ void k_seg_prep(const uint32_t *__restrict__ binstart, uint32_t nbins, int ss, int sb,
                           uint32_t nseg, uint32_t nA, uint32_t *__restrict__ cursorB,
                           uint32_t *__restrict__ cursorA, uint32_t *__restrict__ maxcnt) {
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t mx = 0;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += stride) {
    const uint64_t b0 = (uint64_t)s << ss, b1 = (uint64_t)(s + 1) << ss;
    const uint32_t lo = binstart[b0 < nbins ? b0 : nbins], hi = binstart[b1 < nbins ? b1 : nbins];
    cursorB[s] = lo;
    mx = max(mx, hi - lo);
    if ((s & ((1u << sb) - 1u)) == 0) cursorA[s >> sb] = lo;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(maxcnt, mx);
  (void)nA;
}

// ------------------------------------------------------------------------------ 3/4. partition
// One pass: a block takes tiles of TILE points, ranks them inside the tile by
// digit = (segment >> shift) - (smallest such value still pending in the tile), reserves one run
// per non-empty digit with a single global atomic on that bucket's cursor and writes the tile
// out run by run (consecutive threads -> consecutive records).  Tiles of the raw pass see at
// most kPartFanout digits by construction; tiles of a record pass lie inside one or two
// level-A buckets, and the rare tile that spans more is finished in further rounds.
template<class T> struct PartCfg {
  static constexpr int THREADS = 512;
  static constexpr int TILE    = sizeof(T) == 4 ? 4096 : 2048;
  static constexpr int PPT     = TILE / THREADS;
  static constexpr size_t SMEM = (size_t)TILE * sizeof(Packed4<T>) +
                                 (3 * kPartFanout + 1) * sizeof(uint32_t) +
                                 (size_t)TILE * sizeof(uint16_t);
};

// two resident blocks per SM (<= 64 registers): the pass is latency-bound between its barriers,
// measured 25 % warp occupancy with one 100-register block (profiles/r2g_setpts_ncu.txt)
#ifndef B200_PART_MINBLOCKS
#define B200_PART_MINBLOCKS 2
#endif
template<class T, int DIM, bool RAW>
__global__ void __launch_bounds__(PartCfg<T>::THREADS, B200_PART_MINBLOCKS)
k_part(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
       const Packed4<T> *__restrict__ rin, uint32_t M, GridGeom<T> g, int ss, int shift,
       uint32_t *__restrict__ cursor, Packed4<T> *__restrict__ rout) {
  using CF = PartCfg<T>;
  constexpr int TILE = CF::TILE, NT = CF::THREADS, PPT = CF::PPT, PF = kPartFanout;
  extern __shared__ __align__(16) unsigned char sm[];
  Packed4<T> *rec = reinterpret_cast<Packed4<T> *>(sm);
  uint32_t *cnt   = reinterpret_cast<uint32_t *>(sm + (size_t)TILE * sizeof(Packed4<T>));
  uint32_t *off = cnt + PF, *gb = off + PF + 1;
  uint16_t *dsm = reinterpret_cast<uint16_t *>(gb + PF);
  __shared__ uint32_t wsum[NT / 32];
  __shared__ uint32_t base_s, left_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ntiles = (M + TILE - 1) / TILE;
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    Packed4<T> r[PPT];
    uint32_t dg[PPT], rk[PPT];
    bool todo[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const uint32_t i = t * TILE + k * NT + tid;
      todo[k]          = i < M;
      if (todo[k]) {
        if (RAW) {
          r[k].x = __ldcs(x + i);
          r[k].y = DIM > 1 ? __ldcs(y + i) : (T)0;
          r[k].z = DIM > 2 ? __ldcs(z + i) : (T)0;
          r[k].w = idx_as((T)0, i);
        } else {
          r[k] = rin[i];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k)
      dg[k] = todo[k] ? (bin_key_of<T, DIM>(r[k].x, r[k].y, r[k].z, idx_of(r[k].w), g) >> ss) >> shift
                      : 0xffffffffu;
    for (;;) {  // rounds: one unless the tile spans more than PF buckets
      if (tid == 0) base_s = 0xffffffffu, left_s = 0;
      for (int d = tid; d < PF; d += NT) cnt[d] = 0;
      __syncthreads();
      uint32_t mn = 0xffffffffu;
#pragma unroll
      for (int k = 0; k < PPT; ++k)
        if (todo[k]) mn = min(mn, dg[k]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      if (lane == 0 && mn != 0xffffffffu) atomicMin(&base_s, mn);
      __syncthreads();
      const uint32_t base = base_s;
      bool in[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        in[k] = todo[k] && dg[k] - base < (uint32_t)PF;
        if (in[k]) rk[k] = atomicAdd(&cnt[dg[k] - base], 1u);
      }
      __syncthreads();
      {  // exclusive scan of the PF counters (two per thread), one reservation per non-empty digit
        const uint32_t a = cnt[2 * tid], b = cnt[2 * tid + 1];
        uint32_t incl = a + b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += up;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += wsum[w];
        off[2 * tid]     = wbase + incl - a - b;
        off[2 * tid + 1] = wbase + incl - b;
        if (tid == NT - 1) off[PF] = wbase + incl;
        if (a) gb[2 * tid] = atomicAdd(&cursor[base + 2 * tid], a);
        if (b) gb[2 * tid + 1] = atomicAdd(&cursor[base + 2 * tid + 1], b);
      }
      __syncthreads();
      bool left = false;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        if (in[k]) {
          const uint32_t d = dg[k] - base, p = off[d] + rk[k];
          rec[p]  = r[k];
          dsm[p]  = (uint16_t)d;
          todo[k] = false;
        }
        left = left || todo[k];
      }
      if (left) left_s = 1;  // benign race: every writer stores 1
      __syncthreads();
      const uint32_t n = off[PF];
      for (uint32_t p = tid; p < n; p += NT) {
        const uint32_t d = dsm[p];
        rout[gb[d] + (p - off[d])] = rec[p];
      }
      const bool more = left_s != 0;
      __syncthreads();
      if (!more) break;
    }
  }
}

// ------------------------------------------------------------------------------ 5. segment sort
// window class of a point inside its bin: what the sweep kernels key their runs on
template<class T>
__device__ __forceinline__ int stencil_first(T X, int ns) {  // ceil(X - ns/2), devmath.cuh
  return (int)ceil_t(sub_rn(X, (T)0.5 * (T)ns));
}
template<class T, int CLS>
__device__ __forceinline__ int window_class(T x, T y, uint32_t bin, const GridGeom<T> &g, int ns) {
  if (CLS == kClassNone) return 0;
  const int i1 = bin % g.nb[0], i2 = (bin / g.nb[0]) % g.nb[1];
  const int i0 = stencil_first<T>(fold_rescale<T>(x, g.nf_t[0]), ns);
  const int j0 = stencil_first<T>(fold_rescale<T>(y, g.nf_t[1]), ns);
  const int HL = ns / 2, NJB = kBinY + 1;
  const int jb = min(max(j0 - (kBinY * i2 - HL), 0), kBinY);
  if (CLS == kClassSweep3) {  // sweep3d.cu: window positions step two cells, XB = 4
    const int gg = min(max((i0 - (kBinX * i1 - 4)) >> 1, 0), kBinX / 2);
    return gg * NJB + jb;
  }
  // sweep2d.cuh Sweep2Win<NS>: W = 8 or 16 columns, step S = 2 if ns+1 <= W else 1, XB = 16
  const int W = ns <= 8 ? 8 : 16, S = ns + 1 <= W ? 2 : 1, SH = S - 1, XB = 16;
  const int NG    = kBinX / S + 2;
  const int gbase = (kBinX * i1 - HL + XB) >> SH;
  const int gg    = min(max(((i0 + XB) >> SH) - gbase, 0), NG - 1);
  return gg * NJB + jb;
}

template<class T> struct SegCfg {
  static constexpr int THREADS = 512;
  static constexpr int CAP     = (int)SegCap<T>::value;
  static constexpr int NK      = kPartFanout;
  static constexpr size_t SMEM = (size_t)CAP * sizeof(Packed4<T>) + 3 * (size_t)CAP * sizeof(uint16_t) +
                                 (2 * NK + 1) * sizeof(uint32_t);
};
constexpr int kIdxGroup = 48;  // groups up to this size are put in index order

template<class T, int DIM, int CLS>
__global__ void __launch_bounds__(SegCfg<T>::THREADS)
k_seg_sort(const Packed4<T> *__restrict__ rin, const uint32_t *__restrict__ binstart,
           uint32_t nbins, GridGeom<T> g, int ss, int ncls, int ns, T *__restrict__ xs,
           T *__restrict__ ys, T *__restrict__ zs, uint32_t *__restrict__ sidx) {
  using CF = SegCfg<T>;
  constexpr int NT = CF::THREADS, CAP = CF::CAP, NK = CF::NK;
  extern __shared__ __align__(16) unsigned char sm[];
  Packed4<T> *rec = reinterpret_cast<Packed4<T> *>(sm);
  uint32_t *cnt   = reinterpret_cast<uint32_t *>(sm + (size_t)CAP * sizeof(Packed4<T>));
  uint32_t *fill  = cnt + NK + 1;
  uint16_t *skey = reinterpret_cast<uint16_t *>(fill + NK), *sord = skey + CAP, *sfin = sord + CAP;
  __shared__ uint32_t wsum[NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t s  = blockIdx.x;
  const uint64_t b0 = (uint64_t)s << ss, b1 = (uint64_t)(s + 1) << ss;
  const uint32_t q0 = binstart[b0 < nbins ? b0 : nbins], q1 = binstart[b1 < nbins ? b1 : nbins];
  const int n = (int)min(q1 - q0, (uint32_t)CAP);
  if (n == 0) return;
  for (int d = tid; d < NK; d += NT) cnt[d] = 0, fill[d] = 0;
  __syncthreads();
  for (int k = tid; k < n; k += NT) {
    const Packed4<T> p = rin[q0 + k];
    rec[k]             = p;
    const uint32_t bin = bin_key_of<T, DIM>(p.x, p.y, p.z, idx_of(p.w), g);
    const int key = (int)(bin - (uint32_t)b0) * ncls + window_class<T, CLS>(p.x, p.y, bin, g, ns);
    skey[k]       = (uint16_t)key;
    atomicAdd(&cnt[key], 1u);
  }
  __syncthreads();
  {  // exclusive scan of the NK counters (two per thread); cnt[NK] = n
    const uint32_t a = cnt[2 * tid], b = cnt[2 * tid + 1];
    uint32_t incl = a + b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += up;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wsum[w];
    cnt[2 * tid]     = wbase + incl - a - b;
    cnt[2 * tid + 1] = wbase + incl - b;
    if (tid == NT - 1) cnt[NK] = wbase + incl;
  }
  __syncthreads();
  for (int k = tid; k < n; k += NT) {
    const int key = skey[k];
    sord[cnt[key] + atomicAdd(&fill[key], 1u)] = (uint16_t)k;
  }
  __syncthreads();
  // index order inside every (bin, class) group: rank by comparison (groups hold a handful of
  // points; the rare large group keeps its arrival order)
  for (int p = tid; p < n; p += NT) {
    const int k = sord[p], key = skey[k];
    const int g0 = (int)cnt[key], g1 = (int)cnt[key + 1];
    int dst = p;
    if (g1 - g0 <= kIdxGroup) {
      const uint32_t mine = idx_of(rec[k].w);
      int r = 0;
      for (int q = g0; q < g1; ++q) r += idx_of(rec[sord[q]].w) < mine ? 1 : 0;
      dst = g0 + r;
    }
    sfin[dst] = (uint16_t)k;
  }
  __syncthreads();
  for (int p = tid; p < n; p += NT) {
    const Packed4<T> v = rec[sfin[p]];
    xs[q0 + p] = v.x;
    if (DIM > 1) ys[q0 + p] = v.y;
    if (DIM > 2) zs[q0 + p] = v.z;
    sidx[q0 + p] = idx_of(v.w);
  }
}

// ------------------------------------------------------------------------------ driver
static void cu(cudaError_t e) {
  if (e == cudaSuccess) return;
  cudaGetLastError();
  throw Failure{e == cudaErrorMemoryAllocation ? ERR_ALLOC : ERR_CUDA_FAILURE};
}

template<class T, int DIM>
static bool partition_sort_dim(const T *x, const T *y, const T *z, uint32_t M,
                               const GridGeom<T> &g, const PartPlan &pp, uint32_t *binstart, T *xs,
                               T *ys, T *zs, uint32_t *sidx, uint32_t *scan_tmp, int device,
                               cudaStream_t st) {
  const uint32_t nbins = g.nbins;
  Scratch<uint32_t> cnt(nbins, st, device), cursorB(pp.nseg, st, device),
      cursorA(pp.nA, st, device), dmax(1, st, device);
  cu(cudaMemsetAsync(cnt.p, 0, sizeof(uint32_t) * nbins, st));
  cu(cudaMemsetAsync(dmax.p, 0, sizeof(uint32_t), st));
  if (M >= (1u << 22)) {
    // Clustered input cannot take this path (a segment must fit shared memory) and its full
    // histogram is expensive (same-address atomics): look at 65536 evenly spaced points first
    // and leave at once when a single bin alone would overflow a segment.
    const uint32_t step = M >> 16;
    k_bin_hist_sample<T, DIM><<<64, 256, 0, st>>>(x, y, z, M, step, g, cnt.p);
    k_max_u32<<<blocks_for(nbins, 256, 4), 256, 0, st>>>(cnt.p, nbins, dmax.p);
    uint32_t seen = 0;
    cu(cudaMemcpyAsync(&seen, dmax.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    cu(cudaStreamSynchronize(st));
    if ((uint64_t)seen * step > 2ull * SegCap<T>::value && seen >= 8) return false;
    cu(cudaMemsetAsync(cnt.p, 0, sizeof(uint32_t) * nbins, st));
    cu(cudaMemsetAsync(dmax.p, 0, sizeof(uint32_t), st));
  }
  k_bin_hist<T, DIM><<<blocks_for(M, 256, 8), 256, 0, st>>>(x, y, z, M, g, cnt.p);
  exclusive_scan_u32(cnt.p, binstart, nbins, scan_tmp, st);
  k_seg_prep<<<blocks_for(pp.nseg, 256, 8), 256, 0, st>>>(binstart, nbins, pp.ss, pp.sb, pp.nseg,
                                                         pp.nA, cursorB.p, cursorA.p, dmax.p);
  uint32_t largest = 0;
  cu(cudaMemcpyAsync(&largest, dmax.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  cu(cudaStreamSynchronize(st));
  if (largest > SegCap<T>::value) return false;

  using PC = PartCfg<T>;
  Scratch<Packed4<T>> recB(M, st, device);
  const int nblk = (int)std::min<uint64_t>(((uint64_t)M + PC::TILE - 1) / PC::TILE,
                                          148ull * 2 * B200_PART_MINBLOCKS);
  {  // per device, so on every call (cheap)
    cu(cudaFuncSetAttribute(k_part<T, DIM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)PC::SMEM));
    cu(cudaFuncSetAttribute(k_part<T, DIM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)PC::SMEM));
    cu(cudaFuncSetAttribute(k_part<T, DIM, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    cu(cudaFuncSetAttribute(k_part<T, DIM, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
    cu(cudaFuncSetAttribute(k_seg_sort<T, DIM, kClassNone>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SegCfg<T>::SMEM));
    cu(cudaFuncSetAttribute(k_seg_sort<T, DIM, kClassSweep3>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SegCfg<T>::SMEM));
    cu(cudaFuncSetAttribute(k_seg_sort<T, DIM, kClassSweep2>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SegCfg<T>::SMEM));
  }
  if (pp.levels == 1) {
    k_part<T, DIM, true><<<nblk, PC::THREADS, PC::SMEM, st>>>(x, y, z, nullptr, M, g, pp.ss, 0,
                                                             cursorB.p, recB.p);
  } else {
    Scratch<Packed4<T>> recA(M, st, device);
    k_part<T, DIM, true><<<nblk, PC::THREADS, PC::SMEM, st>>>(x, y, z, nullptr, M, g, pp.ss, pp.sb,
                                                             cursorA.p, recA.p);
    k_part<T, DIM, false><<<nblk, PC::THREADS, PC::SMEM, st>>>(nullptr, nullptr, nullptr, recA.p,
                                                              M, g, pp.ss, 0, cursorB.p, recB.p);
  }
  using SC = SegCfg<T>;
  if (pp.cls == kClassSweep3)
    k_seg_sort<T, DIM, kClassSweep3><<<pp.nseg, SC::THREADS, SC::SMEM, st>>>(
        recB.p, binstart, nbins, g, pp.ss, pp.ncls, pp.ns, xs, ys, zs, sidx);
  else if (pp.cls == kClassSweep2)
    k_seg_sort<T, DIM, kClassSweep2><<<pp.nseg, SC::THREADS, SC::SMEM, st>>>(
        recB.p, binstart, nbins, g, pp.ss, pp.ncls, pp.ns, xs, ys, zs, sidx);
  else
    k_seg_sort<T, DIM, kClassNone><<<pp.nseg, SC::THREADS, SC::SMEM, st>>>(
        recB.p, binstart, nbins, g, pp.ss, pp.ncls, pp.ns, xs, ys, zs, sidx);
  cu(cudaGetLastError());
  return true;
}

template<class T>
bool partition_sort(int dim, const T *x, const T *y, const T *z, uint32_t M,
                    const GridGeom<T> &g, const PartPlan &pp, uint32_t *binstart, T *xs, T *ys,
                    T *zs, uint32_t *sidx, uint32_t *scan_tmp, int device, cudaStream_t st) {
  if (dim == 1)
    return partition_sort_dim<T, 1>(x, y, z, M, g, pp, binstart, xs, ys, zs, sidx, scan_tmp,
                                    device, st);
  if (dim == 2)
    return partition_sort_dim<T, 2>(x, y, z, M, g, pp, binstart, xs, ys, zs, sidx, scan_tmp,
                                    device, st);
  return partition_sort_dim<T, 3>(x, y, z, M, g, pp, binstart, xs, ys, zs, sidx, scan_tmp, device,
                                  st);
}
template bool partition_sort<float>(int, const float *, const float *, const float *, uint32_t,
                                    const GridGeom<float> &, const PartPlan &, uint32_t *, float *,
                                    float *, float *, uint32_t *, uint32_t *, int, cudaStream_t);
template bool partition_sort<double>(int, const double *, const double *, const double *, uint32_t,
                                     const GridGeom<double> &, const PartPlan &, uint32_t *,
                                     double *, double *, double *, uint32_t *, uint32_t *, int,
                                     cudaStream_t);

}  // namespace b200
