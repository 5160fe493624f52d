// 3D single-precision spreading / interpolation, "row sweep" design (the C3 hot path).
//
// What it computes (reference CPU semantics, identical to spreadinterp.cuh):
//   spread  fw[n] += sum_j c_j phi(X_j-n1) phi(Y_j-n2) phi(Z_j-n3)
//           include/finufft/spreadinterp.hpp:310-485, include/finufft/spread.hpp:301-451
//   interp  c_j = sum_n fw[n] phi phi phi,  include/finufft/interp.hpp:281-355, 457-556
//   stencil start ceil(X - ns/2), ns^3 cells, periodic wrap (spread.hpp:328-336).
//
// Why a different kernel: at ns=7 a point touches 343 cells (686 FMAs) but moves only 24
// bytes, so the kernel is bound by the FP32 pipe and by whatever carries the accumulators.
// Shared-memory read-modify-write costs 16 B of shared traffic per cell update (8 cell
// updates/clk/SM at best, 15 ms for 1e8 points); packed FFMA2 on register accumulators
// reaches 64 cell updates/clk/SM.  So the accumulators live in registers:
//
//   * a warp owns one row of the CPU bin grid, (y bin i2, z bin i3), and sweeps it along x.
//     Bins of a row are consecutive in the bin order, so the row is one contiguous run of
//     the sorted point arrays.  All stencils of the row lie in the tile
//     y in [4*i2-ns/2, +4+ns), z in [4*i3-ns/2, +4+ns);
//   * lane (a, bq), a = lane/4, bq = lane%4, owns, of the current 8-row x window
//     [jw, jw+8), the row x = a (mod 8) and the tile's z rows bq, bq+4, bq+8; for these it
//     keeps the whole y extent of the tile (4+ns cells, complex) in registers.  A point whose
//     stencil starts at jw or jw+1 therefore hits every lane: no atomics and no shared-memory
//     traffic for the accumulators;
//   * setpts orders the points inside each bin by (x window position, y stencil start).  The
//     y offset jb of a run of equal keys is warp-uniform, so a 5-way switch picks a fully
//     unrolled body with static register indices: ns packed FFMA2 (re,im) per owned row, the
//     next record prefetched while they issue;
//   * when the window slides by two x rows, the quarter of the lanes that own the leaving rows
//     park them in a 4-column staging tile in shared memory; every second slide the warp adds
//     the tile to the fine grid with 16-byte vector reductions, one full 32-byte sector per
//     (y,z) line (spread).  Interp runs the mirror image: entering rows are read from the
//     staging tile, which is filled with 16-byte loads, and the per-lane shares of a point are
//     reduced with two shuffles plus a shared-memory transpose.
//
// Point data is prepared thread-per-point (fold, stencil starts, packed Horner windows) into
// shared-memory records, 32 points at a time, from coordinates prefetched two chunks ahead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "devmath.cuh"
#include "sort.cuh"
#include "spreadinterp.cuh"

namespace b200 {

// kernel widths the sweep kernels are instantiated for: every single-precision width whose two
// stencil starts fit the 8-row window (tol >= ~1e-6 at sigma = 2); wider ones (sigma = 1.25 at
// tight tolerances) and double precision use the generic kernels
inline bool sweep3_supported(int ns) { return ns >= 2 && ns <= 7; }

// points in refined bin order plus the work items built at setpts (sort.cuh)
struct SweepPoints {
  const float *xs, *ys, *zs;
  const uint32_t *sidx;
  const SweepItem *items;
  uint32_t nitems;
};
constexpr uint32_t kSweepItemPoints = 16384;  // most points one warp takes

cudaError_t launch_spread3_sweep(int ns, const SweepPoints &pts, const GridGeom<float> &g, int nc,
                                 const float *coef, const float2 *c_in, float2 *fw,
                                 cudaStream_t st);
cudaError_t launch_interp3_sweep(int ns, const SweepPoints &pts, const GridGeom<float> &g, int nc,
                                 const float *coef, float2 *c_out, const float2 *fw,
                                 cudaStream_t st);

}  // namespace b200
