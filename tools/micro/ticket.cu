// Micro-benchmark for the setpts partition: 1e8 records of 16 bytes appended to F buckets through
// global cursors ("ticketing").  The question: with F write frontiers (one partially filled
// 128-byte line each) does the L2 merge the 16-byte appends into full-line DRAM writes?
//   nvcc -arch=sm_100a -O3 ticket.cu -o ticket && ./ticket
// Run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
// for the traffic.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16, x *= 0x7feb352dU, x ^= x >> 15, x *= 0x846ca68bU, x ^= x >> 16;
  return x;
}

// mode 0: one atomic per point; mode 1: warp-aggregated (match_any)
template<int MODE>
__global__ void __launch_bounds__(256) k_ticket(const float *__restrict__ x, const float *__restrict__ y,
                                                 const float *__restrict__ z, uint32_t M, uint32_t F,
                                                 uint32_t *__restrict__ cursor, float4 *__restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (uint32_t i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < M; i0 += stride) {
    const uint32_t i = i0 + lane;
    const bool valid = i < M;
    float px = 0, py = 0, pz = 0;
    if (valid) px = __ldcs(x + i), py = __ldcs(y + i), pz = __ldcs(z + i);
    // bucket from the coordinates (all three in [0,1)): uniform over F
    uint32_t key = valid ? (uint32_t)(((uint64_t)hash32(__float_as_uint(px) ^ hash32(__float_as_uint(py) ^ hash32(__float_as_uint(pz)))) * F) >> 32) : 0xffffffffu;
    uint32_t pos;
    if (MODE == 0) {
      pos = valid ? atomicAdd(&cursor[key], 1u) : 0u;
    } else {
      const uint32_t peers = __match_any_sync(0xffffffffu, key);
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (valid && lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      pos = base + __popc(peers & ((1u << lane) - 1u));
    }
    if (valid) out[pos] = make_float4(px, py, pz, __uint_as_float(i));
  }
}

// variant A: the atomics alone (ticket written back coalesced)
__global__ void __launch_bounds__(256) k_atom_only(const float *__restrict__ x, const float *__restrict__ y,
                                                    const float *__restrict__ z, uint32_t M, uint32_t F,
                                                    uint32_t *__restrict__ cursor, uint32_t *__restrict__ pos_out, int red) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const float px = __ldcs(x + i), py = __ldcs(y + i), pz = __ldcs(z + i);
    const uint32_t key = (uint32_t)(((uint64_t)hash32(__float_as_uint(px) ^ hash32(__float_as_uint(py) ^ hash32(__float_as_uint(pz)))) * F) >> 32);
    if (red) { atomicAdd(&cursor[key], 1u); pos_out[i] = key; }
    else pos_out[i] = atomicAdd(&cursor[key], 1u);
  }
}
// variant B: the scattered 16-byte stores alone, positions from a previous ticket run
__global__ void __launch_bounds__(256) k_scatter_only(const float *__restrict__ x, const float *__restrict__ y,
                                                       const float *__restrict__ z, uint32_t M,
                                                       const uint32_t *__restrict__ pos_in, float4 *__restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const float px = __ldcs(x + i), py = __ldcs(y + i), pz = __ldcs(z + i);
    out[__ldcs(pos_in + i)] = make_float4(px, py, pz, __uint_as_float(i));
  }
}

// variant P: tile-sorted partition.  A block takes tiles of T points, ranks them inside the tile
// by bucket with shared-memory atomics, reserves one run per non-empty bucket with ONE global
// atomic, and writes the tile out bucket by bucket (consecutive threads -> consecutive records).
constexpr int PT = 4096, PTH = 512, PPT = PT / PTH, PF = 1024;
__global__ void __launch_bounds__(PTH) k_part(const float *__restrict__ x, const float *__restrict__ y,
                                               const float *__restrict__ z, uint32_t M, uint32_t F,
                                               uint32_t *__restrict__ cursor, float4 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char sm[];
  float4 *rec = reinterpret_cast<float4 *>(sm);
  uint32_t *cnt = reinterpret_cast<uint32_t *>(sm + PT * 16), *off = cnt + PF, *gb = off + PF + 1;
  uint16_t *dsm = reinterpret_cast<uint16_t *>(gb + PF);
  __shared__ uint32_t wsum[PTH / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ntiles = (M + PT - 1) / PT;
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    for (int d = tid; d < PF; d += PTH) cnt[d] = 0;
    __syncthreads();
    float px[PPT], py[PPT], pz[PPT];
    uint32_t dg[PPT], rk[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const uint32_t i = t * PT + k * PTH + tid;
      if (i < M) px[k] = __ldcs(x + i), py[k] = __ldcs(y + i), pz[k] = __ldcs(z + i);
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const uint32_t i = t * PT + k * PTH + tid;
      if (i < M) {
        dg[k] = (uint32_t)(((uint64_t)hash32(__float_as_uint(px[k]) ^ hash32(__float_as_uint(py[k]) ^ hash32(__float_as_uint(pz[k])))) * F) >> 32);
        rk[k] = atomicAdd(&cnt[dg[k]], 1u);
      }
    }
    __syncthreads();
    {  // exclusive scan of cnt[0..PF): PF/PTH = 2 per thread
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
      off[2 * tid] = wbase + incl - a - b;
      off[2 * tid + 1] = wbase + incl - b;
      if (a) gb[2 * tid] = atomicAdd(&cursor[2 * tid], a);
      if (b) gb[2 * tid + 1] = atomicAdd(&cursor[2 * tid + 1], b);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const uint32_t i = t * PT + k * PTH + tid;
      if (i < M) {
        const uint32_t p = off[dg[k]] + rk[k];
        rec[p] = make_float4(px[k], py[k], pz[k], __uint_as_float(i));
        dsm[p] = (uint16_t)dg[k];
      }
    }
    __syncthreads();
    const uint32_t n = min((uint32_t)PT, M - t * PT);
    for (uint32_t p = tid; p < n; p += PTH) {
      const uint32_t d = dsm[p];
      out[gb[d] + (p - off[d])] = rec[p];
    }
    __syncthreads();
  }
}

// variant Q: the same partition with atomics-free ranking: warp-private counter rows, peers found
// with match_any (or a ballot loop), the leader lane bumps the row non-atomically.
template<int BITS, int BALLOT>
__device__ __forceinline__ uint32_t peers_of(uint32_t key, bool valid) {
  if (!BALLOT) return __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
  uint32_t m = __ballot_sync(0xffffffffu, valid);
  if (!valid) m = ~m;
#pragma unroll
  for (int b = 0; b < BITS; ++b) {
    const uint32_t v = __ballot_sync(0xffffffffu, (key >> b) & 1u);
    m &= ((key >> b) & 1u) ? v : ~v;
  }
  return m;
}
constexpr int QW = PTH / 32;
template<int BALLOT>
__global__ void __launch_bounds__(PTH) k_part2(const float *__restrict__ x, const float *__restrict__ y,
                                                const float *__restrict__ z, uint32_t M, uint32_t F,
                                                uint32_t *__restrict__ cursor, float4 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char sm[];
  float4 *rec = reinterpret_cast<float4 *>(sm);
  uint16_t *wc = reinterpret_cast<uint16_t *>(sm + PT * 16);          // [QW][PF]
  uint32_t *off = reinterpret_cast<uint32_t *>(wc + QW * PF), *gb = off + PF + 1;
  uint16_t *dsm = reinterpret_cast<uint16_t *>(gb + PF);
  __shared__ uint32_t wsum[QW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ntiles = (M + PT - 1) / PT;
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    for (int d = tid; d < QW * PF / 2; d += PTH) reinterpret_cast<uint32_t *>(wc)[d] = 0;
    __syncthreads();
    float px[PPT], py[PPT], pz[PPT];
    uint32_t dg[PPT], rk[PPT];
    // warp w owns elements [w*PPT*32, (w+1)*PPT*32) of the tile, k-th group of 32
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const uint32_t i = t * PT + (warp * PPT + k) * 32 + lane;
      if (i < M) px[k] = __ldcs(x + i), py[k] = __ldcs(y + i), pz[k] = __ldcs(z + i);
    }
    uint16_t *mine = wc + warp * PF;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const uint32_t i = t * PT + (warp * PPT + k) * 32 + lane;
      const bool valid = i < M;
      dg[k] = valid ? (uint32_t)(((uint64_t)hash32(__float_as_uint(px[k]) ^ hash32(__float_as_uint(py[k]) ^ hash32(__float_as_uint(pz[k])))) * F) >> 32) : 0u;
      const uint32_t peers = peers_of<10, BALLOT>(dg[k], valid);
      const uint32_t c = mine[dg[k]];
      rk[k] = c + __popc(peers & lt);
      __syncwarp();
      if (valid && (peers & lt) == 0) mine[dg[k]] = (uint16_t)(c + __popc(peers));
      __syncwarp();
    }
    __syncthreads();
    {  // per digit: exclusive prefix over the warps, total -> scan over digits, reserve globally
      uint32_t tot[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int d = 2 * tid + h;
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < QW; ++w) {
          const uint32_t c = wc[w * PF + d];
          wc[w * PF + d] = (uint16_t)s;
          s += c;
        }
        tot[h] = s;
      }
      const uint32_t a = tot[0], b = tot[1];
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
      off[2 * tid] = wbase + incl - a - b;
      off[2 * tid + 1] = wbase + incl - b;
      if (a) gb[2 * tid] = atomicAdd(&cursor[2 * tid], a);
      if (b) gb[2 * tid + 1] = atomicAdd(&cursor[2 * tid + 1], b);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const uint32_t i = t * PT + (warp * PPT + k) * 32 + lane;
      if (i < M) {
        const uint32_t p = off[dg[k]] + mine[dg[k]] + rk[k];
        rec[p] = make_float4(px[k], py[k], pz[k], __uint_as_float(i));
        dsm[p] = (uint16_t)dg[k];
      }
    }
    __syncthreads();
    const uint32_t n = min((uint32_t)PT, M - t * PT);
    for (uint32_t p = tid; p < n; p += PTH) {
      const uint32_t d = dsm[p];
      out[gb[d] + (p - off[d])] = rec[p];
    }
    __syncthreads();
  }
}

__global__ void k_init(float *x, float *y, float *z, uint32_t M) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    x[i] = (hash32(3 * i) >> 8) * (1.0f / 16777216.0f);
    y[i] = (hash32(3 * i + 1) >> 8) * (1.0f / 16777216.0f);
    z[i] = (hash32(3 * i + 2) >> 8) * (1.0f / 16777216.0f);
  }
}
__global__ void k_cursors(uint32_t *cursor, uint32_t F, uint32_t cap) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += stride) cursor[i] = i * cap;
}
// the second pass reads the records back as a stream
__global__ void k_readback(const float4 *__restrict__ in, uint64_t n, float *sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  float s = 0;
  for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 v = __ldcs(in + i);
    s += v.x + v.w;
  }
  if (s == 12345.f) *sink = s;
}

int main() {
  const uint32_t M = 100000000u;
  float *x, *y, *z, *sink;
  cudaMalloc(&x, M * 4ull), cudaMalloc(&y, M * 4ull), cudaMalloc(&z, M * 4ull), cudaMalloc(&sink, 4);
  k_init<<<148 * 8, 256>>>(x, y, z, M);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const uint32_t Fs[] = {16641u};
  for (uint32_t F : Fs) {
    const uint32_t mean = M / F;
    const uint32_t cap = ((uint32_t)(mean + 8.0 * sqrt((double)mean) + 64) + 7u) & ~7u;
    float4 *out;
    uint32_t *cursor;
    if (cudaMalloc(&out, (uint64_t)F * cap * 16ull) != cudaSuccess) { printf("alloc failed F=%u\n", F); continue; }
    cudaMalloc(&cursor, F * 4ull);
    for (int mode = 0; mode < 2; ++mode) {
      float best = 1e9f;
      for (int r = 0; r < 3; ++r) {
        k_cursors<<<148, 256>>>(cursor, F, cap);
        cudaEventRecord(e0);
        if (mode == 0) k_ticket<0><<<148 * 8, 256>>>(x, y, z, M, F, cursor, out);
        else k_ticket<1><<<148 * 8, 256>>>(x, y, z, M, F, cursor, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      printf("F=%7u cap=%6u mode=%d: %.3f ms  (2.8 GB algorithmic -> %.0f GB/s)\n", F, cap, mode, best,
             2.8 / best * 1e3);
    }
    {
      uint32_t *pos;
      cudaMalloc(&pos, M * 4ull);
      for (int red = 0; red < 2; ++red) {
        k_cursors<<<148, 256>>>(cursor, F, cap);
        cudaEventRecord(e0);
        k_atom_only<<<148 * 8, 256>>>(x, y, z, M, F, cursor, pos, red);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("F=%7u atomics only (%s): %.3f ms\n", F, red ? "RED, no return" : "ATOM with return", ms);
      }
      k_cursors<<<148, 256>>>(cursor, F, cap);
      k_atom_only<<<148 * 8, 256>>>(x, y, z, M, F, cursor, pos, 0);
      for (int r = 0; r < 2; ++r) {
        cudaEventRecord(e0);
        k_scatter_only<<<148 * 8, 256>>>(x, y, z, M, pos, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("F=%7u scatter only: %.3f ms\n", F, ms);
      }
      cudaFree(pos);
    }
    cudaEventRecord(e0);
    k_readback<<<148 * 8, 256>>>(out, (uint64_t)F * cap, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("F=%7u readback of %.2f GB: %.3f ms\n", F, (double)F * cap * 16e-9, ms);
    cudaFree(out), cudaFree(cursor);
  }
  {  // tile-sorted partition, few buckets
    const size_t shb = PT * 16 + (3 * PF + 1) * 4 + PT * 2;
    cudaFuncSetAttribute(k_part, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb);
    for (uint32_t F : {129u, 256u, 1024u}) {
      const uint32_t mean = M / F;
      const uint32_t cap = ((uint32_t)(mean + 8.0 * sqrt((double)mean) + 64) + 7u) & ~7u;
      float4 *out;
      uint32_t *cursor;
      cudaMalloc(&out, (uint64_t)F * cap * 16ull);
      cudaMalloc(&cursor, PF * 4ull);
      for (int r = 0; r < 3; ++r) {
        k_cursors<<<148, 256>>>(cursor, F, cap);
        cudaEventRecord(e0);
        k_part<<<148 * 2, PTH, shb>>>(x, y, z, M, F, cursor, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("k_part F=%5u: %.3f ms (%.0f GB/s of 2.8 GB)\n", F, ms, 2.8 / ms * 1e3);
      }
      cudaFree(out), cudaFree(cursor);
    }
  }
  {  // atomics-free ranking
    const size_t shb = PT * 16 + QW * PF * 2 + (2 * PF + 1) * 4 + PT * 2;
    cudaFuncSetAttribute(k_part2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb);
    cudaFuncSetAttribute(k_part2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb);
    for (uint32_t F : {129u, 1024u}) {
      const uint32_t mean = M / F;
      const uint32_t cap = ((uint32_t)(mean + 8.0 * sqrt((double)mean) + 64) + 7u) & ~7u;
      float4 *out;
      uint32_t *cursor;
      cudaMalloc(&out, (uint64_t)F * cap * 16ull);
      cudaMalloc(&cursor, PF * 4ull);
      for (int ballot = 0; ballot < 2; ++ballot)
        for (int r = 0; r < 2; ++r) {
          k_cursors<<<148, 256>>>(cursor, F, cap);
          cudaEventRecord(e0);
          if (ballot) k_part2<1><<<148 * 2, PTH, shb>>>(x, y, z, M, F, cursor, out);
          else k_part2<0><<<148 * 2, PTH, shb>>>(x, y, z, M, F, cursor, out);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          printf("k_part2 (%s) F=%5u: %.3f ms (%.0f GB/s of 2.8 GB)\n", ballot ? "ballot" : "match_any", F, ms, 2.8 / ms * 1e3);
        }
      cudaFree(out), cudaFree(cursor);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
