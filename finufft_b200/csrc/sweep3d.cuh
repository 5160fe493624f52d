// 3D single-precision spreading / interpolation, "tube sweep" design (the C3 hot path).
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
//   * a block owns one "tube" of the CPU bin grid: x-bin i1 (16 cells) and z-bin i3 (4 cells),
//     and sweeps it along y, one CPU bin (4 cells in y) per step;
//   * warp w of the block takes the points whose z cell is 4*i3+w.  All their stencils lie in
//     z rows [4*i3+w-ns/2, +ns+1), y rows [j0, j0+ns) and x cells [16*i1-ns/2, +16+ns);
//   * lane (a, bq) owns, at any time, the y row congruent to a (mod ns) of the current
//     ns-row y window and RZ z rows; for these it keeps the whole x row (16+ns cells, complex)
//     in registers.  Every point therefore hits every (a,bq) lane exactly once: no idle
//     lanes inside the stencil, no atomics, no shared-memory traffic for the accumulators;
//   * the x offset of a point is warp-uniform, so a 17-way switch selects a fully unrolled
//     body with static register indices: ns packed FFMA2 (re,im) per row;
//   * when the window slides by one y row, the 1/ns of the lanes that own the leaving row
//     park it in a per-warp staging buffer; once per step the block sums the four warps'
//     overlapping z rows and adds the result to the fine grid with 16-byte vector REDs
//     (spread), or the entering rows are read from a block tile loaded with coalesced
//     16-byte loads (interp).
//
// Point data for a step is prepared thread-per-point (fold, stencil start, Horner windows,
// strength gather) into shared-memory records, bucketed by (z cell, y stencil start).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "devmath.cuh"
#include "sort.cuh"
#include "spreadinterp.cuh"

namespace b200 {

// kernel widths the sweep kernels are instantiated for (others use the generic kernels)
inline bool sweep3_supported(int ns) { return ns == 6 || ns == 7; }

// nsplit: how many y ranges every tube is cut into (one block each)
cudaError_t launch_spread3_sweep(int ns, const PointSet<float> &pts, const GridGeom<float> &g,
                                 int nc, const float *coef, const float2 *c_in, float2 *fw,
                                 cudaStream_t st);
cudaError_t launch_interp3_sweep(int ns, const PointSet<float> &pts, const GridGeom<float> &g,
                                 int nc, const float *coef, float2 *c_out, const float2 *fw,
                                 cudaStream_t st);

}  // namespace b200
