// Two-level permutation of the strengths: see stage.cuh.
#include "stage.cuh"

#include "sort.cuh"

namespace b200 {

int stage_shift(uint64_t M, int elem_bytes) {
  int shift = 25;  // 32 MB
  for (int b = elem_bytes; b > 1; b >>= 1) --shift;
  while (((M + (1ull << shift) - 1) >> shift) > (uint64_t)kStageMaxWindows) ++shift;
  return shift;
}

// counts[w * nunits + u] = sorted positions of unit u whose user index lies in window w.
// One warp per unit; per-warp counters in shared memory, one atomic per distinct window and
// round (warp-aggregated with match_any).
__global__ void __launch_bounds__(256)
k_stage_count(const uint32_t *__restrict__ sidx, uint32_t M, int shift, uint32_t nunits,
              uint32_t *__restrict__ counts) {
  __shared__ uint32_t scnt[8][kStageMaxWindows];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t u = blockIdx.x * (blockDim.x >> 5) + warp; u < nunits; u += nwarps) {
    const uint32_t q0 = u * kStageUnit, q1 = min(M, q0 + kStageUnit);
    scnt[warp][lane] = 0;
    __syncwarp();
    for (uint32_t q = q0 + lane; q < q0 + kStageUnit; q += 32) {
      const bool valid     = q < q1;
      const uint32_t w     = valid ? __ldcs(sidx + q) >> shift : 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, w);
      if (valid && lane == __ffs(peers) - 1) scnt[warp][w] += __popc(peers);
      __syncwarp();
    }
    counts[(size_t)lane * nunits + u] = scnt[warp][lane];
    __syncwarp();
  }
}

// stable placement: slot of q = offsets[w * nunits + u] + number of earlier q of the unit in w
__global__ void __launch_bounds__(256)
k_stage_place(const uint32_t *__restrict__ sidx, uint32_t M, int shift, uint32_t nunits,
              const uint32_t *__restrict__ offsets, uint32_t *__restrict__ perm1,
              uint32_t *__restrict__ perm2, uint32_t *__restrict__ pinv) {
  __shared__ uint32_t srun[8][kStageMaxWindows];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t nwarps  = gridDim.x * (blockDim.x >> 5);
  for (uint32_t u = blockIdx.x * (blockDim.x >> 5) + warp; u < nunits; u += nwarps) {
    const uint32_t q0 = u * kStageUnit, q1 = min(M, q0 + kStageUnit);
    srun[warp][lane] = offsets[(size_t)lane * nunits + u];
    __syncwarp();
    for (uint32_t q = q0 + lane; q < q0 + kStageUnit; q += 32) {
      const bool valid     = q < q1;
      const uint32_t j     = valid ? __ldcs(sidx + q) : 0u;
      const uint32_t w     = valid ? j >> shift : 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, w);
      uint32_t m           = 0;
      if (valid) m = srun[warp][w] + __popc(peers & lt_mask);
      __syncwarp();
      if (valid) {
        perm1[q] = m;
        perm2[m] = j;
        pinv[j]  = m;
        if (lane == __ffs(peers) - 1) srun[warp][w] += __popc(peers);
      }
      __syncwarp();
    }
  }
}

static inline int stage_grid(uint32_t nunits) {
  const uint32_t want = (nunits + 7) / 8;
  return (int)(want < 1 ? 1 : (want > 148u * 8 ? 148u * 8 : want));
}

void build_stage_perms(const uint32_t *sidx, uint32_t M, int shift, uint32_t *counts,
                       uint32_t *offsets, uint32_t *scan_tmp, uint32_t *perm1, uint32_t *perm2,
                       uint32_t *pinv, cudaStream_t st) {
  if (M == 0) return;
  const uint32_t nunits = (M + kStageUnit - 1) / kStageUnit;
  k_stage_count<<<stage_grid(nunits), 256, 0, st>>>(sidx, M, shift, nunits, counts);
  exclusive_scan_u32(counts, offsets, (uint32_t)kStageMaxWindows * nunits, scan_tmp, st);
  k_stage_place<<<stage_grid(nunits), 256, 0, st>>>(sidx, M, shift, nunits, offsets, perm1, perm2,
                                                    pinv);
}

template<class C>
__global__ void __launch_bounds__(256)
k_stage_in(const C *__restrict__ c, const uint32_t *__restrict__ perm2, C *__restrict__ mid,
           uint32_t M) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride)
    mid[m] = __ldg(c + __ldcs(perm2 + m));
}
// user order, coalesced writes; the reads stay inside the window of mid that belongs to the
// window of j being written (a scatter c[perm2[m]] = mid[m] does not merge in the L2: measured
// 3.1 ms and 2.9 GB of DRAM writes at C2, against 0.7 ms for this gather)
template<class C>
__global__ void __launch_bounds__(256)
k_stage_out(const C *__restrict__ mid, const uint32_t *__restrict__ pinv, C *__restrict__ c,
            uint32_t M) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += stride)
    c[j] = __ldg(mid + __ldcs(pinv + j));
}
static inline int stream_grid(uint32_t n) {
  const uint32_t want = (n + 255) / 256;
  return (int)(want < 1 ? 1 : (want > 148u * 16 ? 148u * 16 : want));
}
template<class C>
void launch_stage_in(const C *c, const uint32_t *perm2, C *mid, uint32_t M, cudaStream_t st) {
  if (M) k_stage_in<C><<<stream_grid(M), 256, 0, st>>>(c, perm2, mid, M);
}
template<class C>
void launch_stage_out(const C *mid, const uint32_t *pinv, C *c, uint32_t M, cudaStream_t st) {
  if (M) k_stage_out<C><<<stream_grid(M), 256, 0, st>>>(mid, pinv, c, M);
}
template void launch_stage_in<float2>(const float2 *, const uint32_t *, float2 *, uint32_t,
                                      cudaStream_t);
template void launch_stage_in<double2>(const double2 *, const uint32_t *, double2 *, uint32_t,
                                       cudaStream_t);
template void launch_stage_out<float2>(const float2 *, const uint32_t *, float2 *, uint32_t,
                                       cudaStream_t);
template void launch_stage_out<double2>(const double2 *, const uint32_t *, double2 *, uint32_t,
                                        cudaStream_t);

}  // namespace b200
