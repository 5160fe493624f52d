// setpts kernels: bin keys, stable LSD radix sort, bin bounds, coordinate gather, subproblems.
// See sort.cuh for the contract and the reference lines the permutation must reproduce.
#include "devmath.cuh"
#include "sort.cuh"
#include "sweep2d.cuh"

namespace b200 {

// ------------------------------------------------------------------------------ bin keys
template<class T, int DIM>
__global__ void k_bin_keys(const T *__restrict__ x, const T *__restrict__ y,
                           const T *__restrict__ z, uint32_t M, GridGeom<T> g,
                           uint32_t *__restrict__ keys) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    // 1/binsize are powers of two, so the product is exact; conversion truncates (X >= 0).
    uint32_t key = (uint32_t)(int)mul_rn(fold_rescale<T>(x[i], g.nf_t[0]), (T)(1.0 / kBinX));
    if (DIM > 1)
      key += (uint32_t)g.nb[0] *
             (uint32_t)(int)mul_rn(fold_rescale<T>(y[i], g.nf_t[1]), (T)(1.0 / kBinY));
    if (DIM > 2)
      key += (uint32_t)g.nb[0] * (uint32_t)g.nb[1] *
             (uint32_t)(int)mul_rn(fold_rescale<T>(z[i], g.nf_t[2]), (T)(1.0 / kBinZ));
    keys[i] = key < g.nbins1 ? key : g.nbins1 - 1;  // only non-finite input can trip this
  }
}

static inline int grid_for(uint32_t n, int threads, int per_sm = 8) {
  const long long want = ((long long)n + threads - 1) / threads;
  const long long cap  = 148LL * per_sm;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template<class T>
void launch_bin_keys(int dim, const T *x, const T *y, const T *z, uint32_t M,
                     const GridGeom<T> &g, uint32_t *keys, cudaStream_t st) {
  if (M == 0) return;
  const int nb = grid_for(M, 256);
  if (dim == 1) k_bin_keys<T, 1><<<nb, 256, 0, st>>>(x, y, z, M, g, keys);
  else if (dim == 2) k_bin_keys<T, 2><<<nb, 256, 0, st>>>(x, y, z, M, g, keys);
  else k_bin_keys<T, 3><<<nb, 256, 0, st>>>(x, y, z, M, g, keys);
}
template void launch_bin_keys<float>(int, const float *, const float *, const float *, uint32_t,
                                     const GridGeom<float> &, uint32_t *, cudaStream_t);
template void launch_bin_keys<double>(int, const double *, const double *, const double *,
                                      uint32_t, const GridGeom<double> &, uint32_t *,
                                      cudaStream_t);

// ------------------------------------------------------------------------------ block scan
template<int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums,
                                                         uint32_t &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NT / 32 ? warp_sums[lane] : 0u, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += up;
    }
    if (lane < NT / 32) warp_sums[lane] = wi - w;
    if (lane == 31) warp_sums[32] = wi;
  }
  __syncthreads();
  total                = warp_sums[32];
  const uint32_t result = incl - v + warp_sums[warp];
  __syncthreads();
  return result;
}

// ------------------------------------------------------------------------------ device scan
constexpr int kScanThreads = 512, kScanItems = 8, kScanChunk = kScanThreads * kScanItems;

__global__ void k_scan_partial(const uint32_t *__restrict__ in, uint32_t n,
                               uint32_t *__restrict__ partial) {
  __shared__ uint32_t ws[33];
  const uint32_t base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) s += in[base + k];
  uint32_t tot;
  block_exclusive_scan<kScanThreads>(s, ws, tot);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
__global__ void k_scan_partials_inplace(uint32_t *partial, uint32_t nblk) {
  __shared__ uint32_t ws[33];
  uint32_t carry = 0;
  for (uint32_t c0 = 0; c0 < nblk; c0 += 1024) {
    const uint32_t i = c0 + threadIdx.x;
    const uint32_t v = i < nblk ? partial[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan<1024>(v, ws, tot);
    if (i < nblk) partial[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) partial[nblk] = carry;
}
__global__ void k_scan_final(const uint32_t *__restrict__ in, uint32_t n,
                             const uint32_t *__restrict__ partial, uint32_t nblk,
                             uint32_t *__restrict__ out) {
  __shared__ uint32_t ws[33];
  const uint32_t base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  uint32_t v[kScanItems], s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  uint32_t tot;
  uint32_t run = block_exclusive_scan<kScanThreads>(s, ws, tot) + partial[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = partial[nblk];
}
void exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tmp,
                        cudaStream_t st) {
  const uint32_t nblk = (n + kScanChunk - 1) / kScanChunk;
  if (nblk == 0) {
    cudaMemsetAsync(out, 0, sizeof(uint32_t), st);
    return;
  }
  k_scan_partial<<<nblk, kScanThreads, 0, st>>>(in, n, tmp);
  k_scan_partials_inplace<<<1, 1024, 0, st>>>(tmp, nblk);
  k_scan_final<<<nblk, kScanThreads, 0, st>>>(in, n, tmp, nblk, out);
}

// ------------------------------------------------------------------------------ radix sort
// One pass = histogram per block (digit-major), scan, stable scatter.  Blocks own contiguous
// ranges of the input, walked tile by tile, so order inside a digit is input order.
constexpr int kRsThreads = 256, kRsItems = 8, kRsTile = kRsThreads * kRsItems, kRsWarps = 8;

__global__ void k_radix_hist(const uint32_t *__restrict__ keys, uint32_t M, int shift,
                             uint32_t mask, uint32_t per_block, uint32_t *__restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lo = blockIdx.x * per_block;
  const uint32_t hi = min(M, lo + per_block);
  for (uint32_t i = lo + threadIdx.x; i < hi; i += kRsThreads)
    atomicAdd(&h[(keys[i] >> shift) & mask], 1u);
  __syncthreads();
  if (threadIdx.x <= mask) hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kRsThreads)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t M,
                int shift, uint32_t mask, uint32_t per_block,
                const uint32_t *__restrict__ hist_scanned) {
  __shared__ uint32_t warp_cnt[kRsWarps][256];
  __shared__ uint32_t run_base[256];
  __shared__ uint32_t tile_base[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (threadIdx.x <= mask) run_base[threadIdx.x] = hist_scanned[threadIdx.x * gridDim.x + blockIdx.x];
  const uint32_t lo = blockIdx.x * per_block;
  const uint32_t hi = min(M, lo + per_block);
  for (uint32_t t0 = lo; t0 < hi; t0 += kRsTile) {
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    uint32_t key[kRsItems], val[kRsItems], rank[kRsItems];
    const uint32_t wbase = t0 + warp * (32 * kRsItems);
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
      const uint32_t e  = wbase + k * 32 + lane;
      const bool valid  = e < hi;
      key[k]            = valid ? keys_in[e] : 0xffffffffu;
      val[k]            = valid ? (vals_in ? vals_in[e] : e) : 0u;
      const uint32_t d  = valid ? ((key[k] >> shift) & mask) : 256u;  // 256 groups the invalid lanes
      const uint32_t peers  = __match_any_sync(0xffffffffu, d);
      const int leader      = __ffs(peers) - 1;
      uint32_t old          = 0;
      if (valid && lane == leader) old = warp_cnt[warp][d];
      old     = __shfl_sync(0xffffffffu, old, leader);
      rank[k] = old + __popc(peers & lt_mask);
      if (valid && lane == leader) warp_cnt[warp][d] = old + __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x <= mask) {  // per digit: exclusive scan over warps, then advance the run
      uint32_t s = 0;
#pragma unroll
      for (int w = 0; w < kRsWarps; ++w) {
        const uint32_t c = warp_cnt[w][threadIdx.x];
        warp_cnt[w][threadIdx.x] = s;
        s += c;
      }
      tile_base[threadIdx.x] = run_base[threadIdx.x];
      run_base[threadIdx.x] += s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
      const uint32_t e = wbase + k * 32 + lane;
      if (e < hi) {
        const uint32_t d   = (key[k] >> shift) & mask;
        const uint32_t pos = tile_base[d] + warp_cnt[warp][d] + rank[k];
        keys_out[pos]      = key[k];
        vals_out[pos]      = val[k];
      }
    }
    __syncthreads();
  }
}

int radix_sort_pairs(uint32_t *keys_a, uint32_t *keys_b, uint32_t *vals_a, uint32_t *vals_b,
                     uint32_t M, int nbits, uint32_t *hist, uint32_t *scan_tmp, cudaStream_t st) {
  if (M == 0) return 0;
  int which = 0;
  const int npass = (nbits + 7) / 8;
  if (npass == 0) {
    launch_iota(vals_a, M, st);
    return 0;
  }
  const int dbits = (nbits + npass - 1) / npass;  // <= 8
  // block ranges: multiple of the tile, at most kRadixMaxBlocks blocks
  uint32_t tiles     = (M + kRsTile - 1) / kRsTile;
  uint32_t nblk      = tiles < kRadixMaxBlocks ? tiles : kRadixMaxBlocks;
  uint32_t per_block = ((tiles + nblk - 1) / nblk) * kRsTile;
  nblk               = (M + per_block - 1) / per_block;
  for (int p = 0; p < npass; ++p) {
    const int shift     = p * dbits;
    const int bits      = (shift + dbits <= nbits) ? dbits : nbits - shift;
    const uint32_t mask = (1u << bits) - 1u;
    uint32_t *kin = which ? keys_b : keys_a, *kout = which ? keys_a : keys_b;
    uint32_t *vin = which ? vals_b : vals_a, *vout = which ? vals_a : vals_b;
    k_radix_hist<<<nblk, kRsThreads, 0, st>>>(kin, M, shift, mask, per_block, hist);
    exclusive_scan_u32(hist, hist, (mask + 1) * nblk, scan_tmp, st);
    k_radix_scatter<<<nblk, kRsThreads, 0, st>>>(kin, p == 0 ? nullptr : vin, kout, vout, M,
                                                  shift, mask, per_block, hist);
    which ^= 1;
  }
  return which;
}

// ------------------------------------------------------------------------------ bin bounds
__global__ void k_bin_bounds(const uint32_t *__restrict__ keys, uint32_t M, uint32_t nbins,
                             uint32_t *__restrict__ binstart) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const uint32_t k = keys[i];
    if (i == 0) {
      for (uint32_t b = 0; b <= k; ++b) binstart[b] = 0;
    } else {
      const uint32_t prev = keys[i - 1];
      for (uint32_t b = prev + 1; b <= k; ++b) binstart[b] = i;
    }
    if (i == M - 1)
      for (uint32_t b = k + 1; b <= nbins; ++b) binstart[b] = M;
  }
}
void launch_bin_bounds(const uint32_t *sorted_keys, uint32_t M, uint32_t nbins,
                       uint32_t *binstart, cudaStream_t st) {
  if (M == 0) {
    cudaMemsetAsync(binstart, 0, sizeof(uint32_t) * ((size_t)nbins + 1), st);
    return;
  }
  k_bin_bounds<<<grid_for(M, 256), 256, 0, st>>>(sorted_keys, M, nbins, binstart);
}

// ------------------------------------------------------------------------------ gather
template<class T, int DIM>
__global__ void k_gather_coords(const T *__restrict__ x, const T *__restrict__ y,
                                const T *__restrict__ z, const uint32_t *__restrict__ sidx,
                                uint32_t M, T *__restrict__ xs, T *__restrict__ ys,
                                T *__restrict__ zs) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const uint32_t j = sidx[i];
    xs[i] = x[j];
    if (DIM > 1) ys[i] = y[j];
    if (DIM > 2) zs[i] = z[j];
  }
}
template<class T>
void launch_gather_coords(int dim, const T *x, const T *y, const T *z, const uint32_t *sidx,
                          uint32_t M, T *xs, T *ys, T *zs, cudaStream_t st) {
  if (M == 0) return;
  const int nb = grid_for(M, 256, 16);
  if (dim == 1) k_gather_coords<T, 1><<<nb, 256, 0, st>>>(x, y, z, sidx, M, xs, ys, zs);
  else if (dim == 2) k_gather_coords<T, 2><<<nb, 256, 0, st>>>(x, y, z, sidx, M, xs, ys, zs);
  else k_gather_coords<T, 3><<<nb, 256, 0, st>>>(x, y, z, sidx, M, xs, ys, zs);
}
template void launch_gather_coords<float>(int, const float *, const float *, const float *,
                                          const uint32_t *, uint32_t, float *, float *, float *,
                                          cudaStream_t);
template void launch_gather_coords<double>(int, const double *, const double *, const double *,
                                           const uint32_t *, uint32_t, double *, double *,
                                           double *, cudaStream_t);

// ------------------------------------------------------------------------------ subproblems
__global__ void k_sub_count(const uint32_t *__restrict__ binstart, uint32_t nbins,
                            uint32_t maxsub, uint32_t *__restrict__ nsub) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nbins; b += stride) {
    const uint32_t n = binstart[b + 1] - binstart[b];
    nsub[b]          = (n + maxsub - 1) / maxsub;
  }
}
__global__ void k_sub_fill(const uint32_t *__restrict__ binstart,
                           const uint32_t *__restrict__ substart, uint32_t nbins,
                           uint32_t maxsub, uint32_t *__restrict__ sub_bin,
                           uint32_t *__restrict__ sub_off) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nbins; b += stride) {
    const uint32_t s0 = substart[b], s1 = substart[b + 1];
    uint32_t off = binstart[b];
    for (uint32_t s = s0; s < s1; ++s, off += maxsub) {
      sub_bin[s] = b;
      sub_off[s] = off;
    }
  }
}
void launch_sub_count(const uint32_t *binstart, uint32_t nbins, uint32_t maxsub, uint32_t *nsub,
                      cudaStream_t st) {
  k_sub_count<<<grid_for(nbins, 256), 256, 0, st>>>(binstart, nbins, maxsub, nsub);
}
void launch_sub_fill(const uint32_t *binstart, const uint32_t *substart, uint32_t nbins,
                     uint32_t maxsub, uint32_t *sub_bin, uint32_t *sub_off, cudaStream_t st) {
  k_sub_fill<<<grid_for(nbins, 256), 256, 0, st>>>(binstart, substart, nbins, maxsub, sub_bin,
                                                   sub_off);
}

// ------------------------------------------------------------------------------ counting sort
template<class T, int DIM>
__device__ __forceinline__ uint32_t bin_key(T x, T y, T z, const GridGeom<T> &g) {
  // 1/binsize are powers of two, so the product is exact; conversion truncates (X >= 0).
  uint32_t key = (uint32_t)(int)mul_rn(fold_rescale<T>(x, g.nf_t[0]), (T)(1.0 / kBinX));
  if (DIM > 1)
    key += (uint32_t)g.nb[0] * (uint32_t)(int)mul_rn(fold_rescale<T>(y, g.nf_t[1]), (T)(1.0 / kBinY));
  if (DIM > 2)
    key += (uint32_t)g.nb[0] * (uint32_t)g.nb[1] *
           (uint32_t)(int)mul_rn(fold_rescale<T>(z, g.nf_t[2]), (T)(1.0 / kBinZ));
  return key < g.nbins1 ? key : g.nbins1 - 1;  // only non-finite input can trip this
}

template<class T, int DIM>
__global__ void __launch_bounds__(256)
k_bin_count(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, uint32_t M,
            GridGeom<T> g, uint32_t *__restrict__ keys, uint32_t *__restrict__ ranks,
            uint32_t *__restrict__ cnt, Packed4<T> *__restrict__ packed) {
  const int lane         = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t stride  = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < M; i0 += stride) {
    const uint32_t i = i0 + lane;
    const bool valid = i < M;
    T px = 0, py = 0, pz = 0;
    if (valid) {
      px = x[i];
      if (DIM > 1) py = y[i];
      if (DIM > 2) pz = z[i];
    }
    uint32_t key = valid ? bin_key<T, DIM>(px, py, pz, g) : 0xffffffffu;
    if (valid && g.nchunks > 1) key += point_group(g, i) * g.nbins1;  // group-major (sort.cuh)
    // one atomic per distinct bin in the warp
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const int leader     = __ffs(peers) - 1;
    uint32_t base        = 0;
    if (valid && lane == leader) base = atomicAdd(&cnt[key], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (valid) {
      keys[i]   = key;
      ranks[i]  = base + __popc(peers & lt_mask);
      packed[i] = Packed4<T>{px, py, pz, (T)0};
    }
  }
}
template<class T>
void launch_bin_count(int dim, const T *x, const T *y, const T *z, uint32_t M,
                      const GridGeom<T> &g, uint32_t *keys, uint32_t *ranks, uint32_t *cnt,
                      Packed4<T> *packed, cudaStream_t st) {
  if (M == 0) return;
  const int nb = grid_for(M, 256, 16);
  if (dim == 1) k_bin_count<T, 1><<<nb, 256, 0, st>>>(x, y, z, M, g, keys, ranks, cnt, packed);
  else if (dim == 2) k_bin_count<T, 2><<<nb, 256, 0, st>>>(x, y, z, M, g, keys, ranks, cnt, packed);
  else k_bin_count<T, 3><<<nb, 256, 0, st>>>(x, y, z, M, g, keys, ranks, cnt, packed);
}
template void launch_bin_count<float>(int, const float *, const float *, const float *, uint32_t,
                                      const GridGeom<float> &, uint32_t *, uint32_t *, uint32_t *,
                                      Packed4<float> *, cudaStream_t);
template void launch_bin_count<double>(int, const double *, const double *, const double *,
                                       uint32_t, const GridGeom<double> &, uint32_t *, uint32_t *,
                                       uint32_t *, Packed4<double> *, cudaStream_t);

__global__ void k_bin_place(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ ranks,
                            const uint32_t *__restrict__ binstart, uint32_t M,
                            uint32_t *__restrict__ sidx) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride)
    sidx[binstart[keys[i]] + ranks[i]] = i;
}
void launch_bin_place(const uint32_t *keys, const uint32_t *ranks, const uint32_t *binstart,
                      uint32_t M, uint32_t *sidx, cudaStream_t st) {
  if (M) k_bin_place<<<grid_for(M, 256, 16), 256, 0, st>>>(keys, ranks, binstart, M, sidx);
}

template<class T, int DIM>
__global__ void k_gather_packed(const Packed4<T> *__restrict__ packed,
                                const uint32_t *__restrict__ sidx, uint32_t M, T *__restrict__ xs,
                                T *__restrict__ ys, T *__restrict__ zs) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const Packed4<T> p = packed[sidx[i]];
    xs[i] = p.x;
    if (DIM > 1) ys[i] = p.y;
    if (DIM > 2) zs[i] = p.z;
  }
}
template<class T>
void launch_gather_packed(int dim, const Packed4<T> *packed, const uint32_t *sidx, uint32_t M,
                          T *xs, T *ys, T *zs, cudaStream_t st) {
  if (M == 0) return;
  const int nb = grid_for(M, 256, 16);
  if (dim == 1) k_gather_packed<T, 1><<<nb, 256, 0, st>>>(packed, sidx, M, xs, ys, zs);
  else if (dim == 2) k_gather_packed<T, 2><<<nb, 256, 0, st>>>(packed, sidx, M, xs, ys, zs);
  else k_gather_packed<T, 3><<<nb, 256, 0, st>>>(packed, sidx, M, xs, ys, zs);
}
template void launch_gather_packed<float>(int, const Packed4<float> *, const uint32_t *, uint32_t,
                                          float *, float *, float *, cudaStream_t);
template void launch_gather_packed<double>(int, const Packed4<double> *, const uint32_t *,
                                           uint32_t, double *, double *, double *, cudaStream_t);

// ------------------------------------------------------------------------------ sweep support
// One warp per bin.  The bin's points (index + packed coordinates) are pulled into shared
// memory in batches, bucketed by key = g*5 + jb (g: x window position inside the bin in steps
// of two cells, jb: y stencil start inside the bin), ordered by index inside each bucket and
// written out: coordinates to the sorted arrays, indices back to sidx.
constexpr int kRefCap = 512, kRefWarps = 4, kRefKeys = 64;
constexpr int kRefBatch = 8;  // gathers a lane issues back to back
static_assert(kRefCap == (int)kRefineChunk, "chunk size");

template<int NS>
__global__ void __launch_bounds__(kRefWarps * 32)
k_refine_bins3(const Packed4<float> *__restrict__ packed, float *__restrict__ xs,
               float *__restrict__ ys, float *__restrict__ zs, uint32_t *__restrict__ sidx,
               const uint32_t *__restrict__ binstart, const uint32_t *__restrict__ chunk_bin,
               const uint32_t *__restrict__ chunk_off, const uint32_t *__restrict__ nchunks,
               GridGeom<float> g) {
  __shared__ float sx[kRefWarps][kRefCap], sy[kRefWarps][kRefCap], sz[kRefWarps][kRefCap];
  __shared__ uint32_t si[kRefWarps][kRefCap];
  __shared__ uint16_t skey[kRefWarps][kRefCap], sord[kRefWarps][kRefCap];
  __shared__ int cnt[kRefWarps][kRefKeys + 1], fill[kRefWarps][kRefKeys];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * kRefWarps;
  constexpr int HL = NS / 2, XB = 4, NG = kBinX / 2 + 1, NJB = kBinY + 1;
  const uint32_t total = *nchunks;
  for (uint32_t ch = blockIdx.x * kRefWarps + warp; ch < total; ch += nwarps) {
    const uint32_t bin = chunk_bin[ch], q0 = chunk_off[ch];
    const int i1 = bin % g.nb[0], i2 = (bin / g.nb[0]) % g.nb[1];
    {
      const int n = (int)min((uint32_t)kRefCap, binstart[bin + 1] - q0);
      cnt[warp][lane] = 0, cnt[warp][lane + 32] = 0;
      fill[warp][lane] = 0, fill[warp][lane + 32] = 0;
      __syncwarp();
      // the gathers are random 16-byte reads (a 128-byte line of HBM traffic each): issue
      // kRefBatch of them per lane before touching the data, or the kernel is latency-bound
      for (int kb = 0; kb < n; kb += 32 * kRefBatch) {
        uint32_t jj[kRefBatch];
        Packed4<float> pp[kRefBatch];
#pragma unroll
        for (int u = 0; u < kRefBatch; ++u) {
          const int k = kb + 32 * u + lane;
          jj[u]       = k < n ? sidx[q0 + k] : 0u;
        }
#pragma unroll
        for (int u = 0; u < kRefBatch; ++u)
          if (kb + 32 * u + lane < n) pp[u] = packed[jj[u]];
#pragma unroll
        for (int u = 0; u < kRefBatch; ++u) {
          const int k = kb + 32 * u + lane;
          if (k >= n) continue;
          const Packed4<float> pt = pp[u];
          sx[warp][k] = pt.x, sy[warp][k] = pt.y, sz[warp][k] = pt.z, si[warp][k] = jj[u];
          int i0, j0;
          float t;
          stencil_start<float, NS>(fold_rescale<float>(pt.x, g.nf_t[0]), i0, t);
          stencil_start<float, NS>(fold_rescale<float>(pt.y, g.nf_t[1]), j0, t);
          const int gg  = min(max((i0 - (kBinX * i1 - XB)) >> 1, 0), NG - 1);
          const int jb  = min(max(j0 - (kBinY * i2 - HL), 0), kBinY);
          const int key = gg * NJB + jb;
          skey[warp][k] = (uint16_t)key;
          atomicAdd(&cnt[warp][key], 1);
        }
      }
      __syncwarp();
      {  // exclusive scan of the 64 counters, two per lane; cnt[64] = n
        const int v0 = cnt[warp][2 * lane], v1 = cnt[warp][2 * lane + 1];
        int incl = v0 + v1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += up;
        }
        __syncwarp();
        cnt[warp][2 * lane]     = incl - v0 - v1;
        cnt[warp][2 * lane + 1] = incl - v1;
        if (lane == 31) cnt[warp][kRefKeys] = incl;
      }
      __syncwarp();
      // bucket the elements (any order inside a bucket) ...
      for (int k = lane; k < n; k += 32) {
        const int key = skey[warp][k];
        sord[warp][cnt[warp][key] + atomicAdd(&fill[warp][key], 1)] = (uint16_t)k;
      }
      __syncwarp();
      // ... then rank every element inside its bucket by index and write it out
      for (int p = lane; p < n; p += 32) {
        const int k = sord[warp][p], key = skey[warp][k];
        const uint32_t mine = si[warp][k];
        const int b0 = cnt[warp][key], b1 = cnt[warp][key + 1];
        int r = 0;
        for (int q = b0; q < b1; ++q) r += si[warp][sord[warp][q]] < mine ? 1 : 0;
        const uint32_t dst = q0 + b0 + r;
        xs[dst] = sx[warp][k], ys[dst] = sy[warp][k], zs[dst] = sz[warp][k], sidx[dst] = mine;
      }
      __syncwarp();
    }
  }
}

void launch_refine_bins3(int ns, const Packed4<float> *packed, float *xs, float *ys, float *zs,
                         uint32_t *sidx, const uint32_t *binstart, const uint32_t *chunk_bin,
                         const uint32_t *chunk_off, const uint32_t *nchunks, uint32_t max_chunks,
                         const GridGeom<float> &g, cudaStream_t st) {
  if (max_chunks == 0) return;
  const int nb = grid_for((max_chunks + kRefWarps - 1) / kRefWarps * 32 * kRefWarps, kRefWarps * 32, 16);
  switch (ns) {
#define B200_REF(NSV)                                                                          \
  case NSV:                                                                                    \
    k_refine_bins3<NSV><<<nb, kRefWarps * 32, 0, st>>>(packed, xs, ys, zs, sidx, binstart,     \
                                                       chunk_bin, chunk_off, nchunks, g);      \
    break;
    B200_REF(2) B200_REF(3) B200_REF(4) B200_REF(5) B200_REF(6) B200_REF(7)
#undef B200_REF
  default: break;
  }
}

// ---- 2D (sweep2d.cuh): same refinement, any precision and kernel width.  key = gg*5 + jb with
// gg the x window position relative to the bin's first one (window step S = 1 or 2 cells).
template<class T> struct Ref2Cfg {
  static constexpr int WARPS = sizeof(T) == 4 ? 4 : 2;
  static constexpr int KEYS  = 128;
};

template<class T, int NS>
__global__ void __launch_bounds__(Ref2Cfg<T>::WARPS * 32)
k_refine_bins2(const Packed4<T> *__restrict__ packed, T *__restrict__ xs, T *__restrict__ ys,
               uint32_t *__restrict__ sidx, const uint32_t *__restrict__ binstart,
               const uint32_t *__restrict__ chunk_bin, const uint32_t *__restrict__ chunk_off,
               const uint32_t *__restrict__ nchunks, GridGeom<T> g) {
  constexpr int WARPS = Ref2Cfg<T>::WARPS, KEYS = Ref2Cfg<T>::KEYS;
  using WN = Sweep2Win<NS>;
  constexpr int HL = NS / 2, SH = WN::S - 1, XB = WN::XB, NG = kBinX / WN::S + 2,
                NJB = kBinY + 1;
  static_assert(NG * NJB <= KEYS, "key space");
  __shared__ T sx[WARPS][kRefCap], sy[WARPS][kRefCap];
  __shared__ uint32_t si[WARPS][kRefCap];
  __shared__ uint16_t skey[WARPS][kRefCap], sord[WARPS][kRefCap];
  __shared__ int cnt[WARPS][KEYS + 1], fill[WARPS][KEYS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * WARPS;
  const uint32_t total  = *nchunks;
  for (uint32_t ch = blockIdx.x * WARPS + warp; ch < total; ch += nwarps) {
    const uint32_t bin = chunk_bin[ch], q0 = chunk_off[ch];
    const int i1 = bin % g.nb[0], i2 = (bin / g.nb[0]) % g.nb[1];
    const int n  = (int)min((uint32_t)kRefCap, binstart[bin + 1] - q0);
    const int gbase = (kBinX * i1 - HL + XB) >> SH;
#pragma unroll
    for (int k = 0; k < KEYS / 32; ++k) cnt[warp][lane + 32 * k] = 0, fill[warp][lane + 32 * k] = 0;
    __syncwarp();
    constexpr int UB = sizeof(T) == 4 ? kRefBatch : kRefBatch / 2;  // gathers in flight per lane
    for (int kb = 0; kb < n; kb += 32 * UB) {
      uint32_t jj[UB];
      Packed4<T> pp[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int k = kb + 32 * u + lane;
        jj[u]       = k < n ? sidx[q0 + k] : 0u;
      }
#pragma unroll
      for (int u = 0; u < UB; ++u)
        if (kb + 32 * u + lane < n) pp[u] = packed[jj[u]];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int k = kb + 32 * u + lane;
        if (k >= n) continue;
        const Packed4<T> pt = pp[u];
        sx[warp][k] = pt.x, sy[warp][k] = pt.y, si[warp][k] = jj[u];
        int i0, j0;
        T t;
        stencil_start<T, NS>(fold_rescale<T>(pt.x, g.nf_t[0]), i0, t);
        stencil_start<T, NS>(fold_rescale<T>(pt.y, g.nf_t[1]), j0, t);
        const int gg  = min(max(((i0 + XB) >> SH) - gbase, 0), NG - 1);
        const int jb  = min(max(j0 - (kBinY * i2 - HL), 0), kBinY);
        const int key = gg * NJB + jb;
        skey[warp][k] = (uint16_t)key;
        atomicAdd(&cnt[warp][key], 1);
      }
    }
    __syncwarp();
    {  // exclusive scan of the KEYS counters, KEYS/32 per lane; cnt[KEYS] = n
      constexpr int PL = KEYS / 32;
      int v[PL], sum = 0;
#pragma unroll
      for (int k = 0; k < PL; ++k) v[k] = cnt[warp][PL * lane + k], sum += v[k];
      int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      __syncwarp();
      int run = incl - sum;
#pragma unroll
      for (int k = 0; k < PL; ++k) cnt[warp][PL * lane + k] = run, run += v[k];
      if (lane == 31) cnt[warp][KEYS] = incl;
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const int key = skey[warp][k];
      sord[warp][cnt[warp][key] + atomicAdd(&fill[warp][key], 1)] = (uint16_t)k;
    }
    __syncwarp();
    for (int p = lane; p < n; p += 32) {
      const int k = sord[warp][p], key = skey[warp][k];
      const uint32_t mine = si[warp][k];
      const int b0 = cnt[warp][key], b1 = cnt[warp][key + 1];
      int r = 0;
      for (int q = b0; q < b1; ++q) r += si[warp][sord[warp][q]] < mine ? 1 : 0;
      const uint32_t dst = q0 + b0 + r;
      xs[dst] = sx[warp][k], ys[dst] = sy[warp][k], sidx[dst] = mine;
    }
    __syncwarp();
  }
}

template<class T, int NS>
static void refine2_ns(const Packed4<T> *packed, T *xs, T *ys, uint32_t *sidx,
                       const uint32_t *binstart, const uint32_t *chunk_bin,
                       const uint32_t *chunk_off, const uint32_t *nchunks, uint32_t max_chunks,
                       const GridGeom<T> &g, cudaStream_t st) {
  constexpr int WARPS = Ref2Cfg<T>::WARPS;
  const int nb = grid_for((max_chunks + WARPS - 1) / WARPS * 32 * WARPS, WARPS * 32, 16);
  k_refine_bins2<T, NS><<<nb, WARPS * 32, 0, st>>>(packed, xs, ys, sidx, binstart, chunk_bin,
                                                   chunk_off, nchunks, g);
}
template<class T>
void launch_refine_bins2(int ns, const Packed4<T> *packed, T *xs, T *ys, uint32_t *sidx,
                         const uint32_t *binstart, const uint32_t *chunk_bin,
                         const uint32_t *chunk_off, const uint32_t *nchunks, uint32_t max_chunks,
                         const GridGeom<T> &g, cudaStream_t st) {
  if (max_chunks == 0) return;
  switch (ns) {
#define B200_REF2(NSV)                                                                         \
  case NSV:                                                                                    \
    refine2_ns<T, NSV>(packed, xs, ys, sidx, binstart, chunk_bin, chunk_off, nchunks,          \
                       max_chunks, g, st);                                                     \
    break;
    B200_REF2(2) B200_REF2(3) B200_REF2(4) B200_REF2(5) B200_REF2(6) B200_REF2(7) B200_REF2(8)
    B200_REF2(9) B200_REF2(10) B200_REF2(11) B200_REF2(12) B200_REF2(13) B200_REF2(14)
    B200_REF2(15) B200_REF2(16)
#undef B200_REF2
  default: break;
  }
}
template void launch_refine_bins2<float>(int, const Packed4<float> *, float *, float *, uint32_t *,
                                         const uint32_t *, const uint32_t *, const uint32_t *,
                                         const uint32_t *, uint32_t, const GridGeom<float> &,
                                         cudaStream_t);
template void launch_refine_bins2<double>(int, const Packed4<double> *, double *, double *,
                                          uint32_t *, const uint32_t *, const uint32_t *,
                                          const uint32_t *, const uint32_t *, uint32_t,
                                          const GridGeom<double> &, cudaStream_t);

__global__ void k_row_item_count(const uint32_t *__restrict__ binstart, uint32_t nrows,
                                 uint32_t nb1, uint32_t maxpts, uint32_t *__restrict__ nitems) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const uint32_t n = binstart[(size_t)(r + 1) * nb1] - binstart[(size_t)r * nb1];
    nitems[r]        = (n + maxpts - 1) / maxpts;
  }
}
__global__ void k_row_item_fill(const uint32_t *__restrict__ binstart,
                                const uint32_t *__restrict__ itemstart, uint32_t nrows,
                                uint32_t nb1, uint32_t maxpts, SweepItem *__restrict__ items) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const uint32_t qs = binstart[(size_t)r * nb1], qe = binstart[(size_t)(r + 1) * nb1];
    uint32_t s = itemstart[r];
    for (uint32_t q = qs; q < qe; q += maxpts, ++s) items[s] = SweepItem{r, q, min(qe, q + maxpts)};
  }
}
void launch_row_item_count(const uint32_t *binstart, uint32_t nrows, uint32_t nb1,
                           uint32_t maxpts, uint32_t *nitems, cudaStream_t st) {
  k_row_item_count<<<grid_for(nrows, 256), 256, 0, st>>>(binstart, nrows, nb1, maxpts, nitems);
}
void launch_row_item_fill(const uint32_t *binstart, const uint32_t *itemstart, uint32_t nrows,
                          uint32_t nb1, uint32_t maxpts, SweepItem *items, cudaStream_t st) {
  k_row_item_fill<<<grid_for(nrows, 256), 256, 0, st>>>(binstart, itemstart, nrows, nb1, maxpts,
                                                       items);
}

// Reference order inside every bin (ascending user index, include/finufft/spread.hpp:559-581)
// from any order: one warp per bin, indices ranked by comparison in shared memory.  Bins with
// more than kCanonCap points are left as they are and counted in *nbig.
constexpr int kCanonCap = 1024, kCanonWarps = 8;
__global__ void __launch_bounds__(kCanonWarps * 32)
k_canon_bins(uint32_t *__restrict__ perm, const uint32_t *__restrict__ binstart, uint32_t nbins,
             uint32_t *__restrict__ nbig) {
  __shared__ uint32_t s[kCanonWarps][kCanonCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * kCanonWarps;
  for (uint32_t b = blockIdx.x * kCanonWarps + warp; b < nbins; b += nwarps) {
    const uint32_t q0 = binstart[b], n = binstart[b + 1] - q0;
    if (n < 2) continue;
    if (n > (uint32_t)kCanonCap) {
      if (lane == 0) atomicAdd(nbig, 1u);
      continue;
    }
    for (uint32_t k = lane; k < n; k += 32) s[warp][k] = perm[q0 + k];
    __syncwarp();
    for (uint32_t k = lane; k < n; k += 32) {
      const uint32_t mine = s[warp][k];
      uint32_t r = 0;
      for (uint32_t q = 0; q < n; ++q) r += s[warp][q] < mine ? 1u : 0u;
      perm[q0 + r] = mine;
    }
    __syncwarp();
  }
}
void launch_canon_bins(uint32_t *perm, const uint32_t *binstart, uint32_t nbins, uint32_t *nbig,
                       cudaStream_t st) {
  if (nbins == 0) return;
  const uint32_t want = (nbins + kCanonWarps - 1) / kCanonWarps;
  k_canon_bins<<<want < 148u * 8 ? want : 148u * 8, kCanonWarps * 32, 0, st>>>(perm, binstart, nbins,
                                                                              nbig);
}

__global__ void k_iota(uint32_t *v, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = i;
}
void launch_iota(uint32_t *v, uint32_t n, cudaStream_t st) {
  if (n) k_iota<<<grid_for(n, 256), 256, 0, st>>>(v, n);
}

}  // namespace b200
