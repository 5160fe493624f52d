// Two-level permutation of the strengths ("staging").
//
// The spread kernels read c[sidx[q]] and the interp kernels write c[sidx[q]] for sorted
// positions q: a random 8- or 16-byte access into an array far larger than the L2.  On B200 a
// random read that misses the L2 moves a whole 128-byte line from HBM (measured,
// tools/micro/gather.cu: 1e8 float2 gathers = 12 GB of DRAM reads, 2.45 ms; the scatter costs
// 4.4 ms), 16x the useful bytes, which made the 2D kernels DRAM-bound on the permutation alone.
//
// Staging splits the permutation q -> j into two L2-friendly halves through a buffer `mid`:
//   * the user index range is cut into K <= 32 windows of 2^shift elements (32 MB of c each);
//   * mid holds the strengths grouped by window, inside a window in sorted-position order;
//       perm2[m] = user index j of slot m      (m -> j stays inside one 32 MB window of c)
//       perm1[q] = slot m of sorted position q  (consecutive q of one window -> consecutive m)
//   * type 1:  stage_in   mid[m] = c[perm2[m]]   gathers inside an L2-resident window, every
//              line of c leaves HBM once; the spread kernel then reads mid[perm1[q]], K
//              interleaved sequential streams that live in L1/L2;
//   * type 2:  the interp kernel writes mid[perm1[q]] (K sequential write streams, merged in
//              the L2), stage_out  c[j] = mid[pinv[j]]  gathers inside an L2-resident window of
//              mid and writes c coalesced (pinv = inverse of perm2; the scatter form
//              c[perm2[m]] = mid[m] does not merge in the L2 and was measured slower).
// perm1/perm2/pinv are built once per setpts by a stable K-way partition of the sorted positions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kStageMaxWindows = 32;
constexpr uint32_t kStageUnit  = 1024;  // sorted positions one warp partitions

// shift such that windows of 2^shift elements of `elem_bytes` bytes hold about 32 MB and at
// most kStageMaxWindows windows cover M elements
int stage_shift(uint64_t M, int elem_bytes);

// scratch: counts (K * nunits + 1 words, nunits = ceil(M / kStageUnit)), scan_tmp as for
// exclusive_scan_u32.  Writes perm1[M], perm2[M], pinv[M].
void build_stage_perms(const uint32_t *sidx, uint32_t M, int shift, uint32_t *counts,
                       uint32_t *offsets, uint32_t *scan_tmp, uint32_t *perm1, uint32_t *perm2,
                       uint32_t *pinv, cudaStream_t st);

template<class C>
void launch_stage_in(const C *c, const uint32_t *perm2, C *mid, uint32_t M, cudaStream_t st);
template<class C>
void launch_stage_out(const C *mid, const uint32_t *pinv, C *c, uint32_t M, cudaStream_t st);

}  // namespace b200
