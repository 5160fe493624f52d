// 2D spreading / interpolation, "row sweep" design (float and double, every kernel width).
//
// What it computes (reference CPU semantics, identical to spreadinterp.cuh):
//   spread  fw[n] += sum_j c_j phi(X_j-n1) phi(Y_j-n2)
//           include/finufft/spreadinterp.hpp:310-485, include/finufft/spread.hpp:147-299
//   interp  c_j = sum_n fw[n] phi phi,  include/finufft/interp.hpp:123-279, 457-556
//   stencil start ceil(X - ns/2), ns^2 cells, periodic wrap (spread.hpp:328-336).
//
// A 2D point touches only ns^2 cells, so per-point overhead decides the speed.  The design:
//   * a warp owns a run of points of one row of bins (y bin i2, 4 cells tall) and sweeps it
//     along x.  All stencils of the row lie in the y rows [4*i2-ns/2, +4+ns): YR <= 20 rows;
//   * the x window of W (8 or 16) columns lives in the lanes: lane (g, a), a = lane % W, holds
//     column x = a (mod W) with all YR rows in registers.  The 32/W lane groups g keep
//     SEPARATE copies of the window and take different points, so one step of the inner loop
//     handles 32/W points with YR packed FFMA2 (float) or 2*YR DFMA (double) per lane and no
//     cross-lane traffic;
//   * setpts orders the points inside each bin by (x window position, y stencil start).  A
//     point's y window is stored zero-padded at the window rows it multiplies, so the inner
//     loop does not depend on the y stencil start and a run is keyed by the window position
//     alone; the run boundaries of a 32-point chunk come from one shfl_up + ballot;
//   * when the window moves, the lanes owning the leaving columns add them to the fine grid
//     with RED.64 (spread; every group flushes its own copy), or the lanes owning the entering
//     columns install them (interp; the groups hold identical copies; in single precision
//     they were loaded one window move ahead).  Interp: each lane parks its column's share
//     of a point in shared memory and the chunk is summed thread-per-point afterwards;
//   * point data is prepared thread-per-point, 32 points at a time: fold, stencil starts,
//     Horner windows; the x window (times the strength for spread) is written rotated so that
//     slot a belongs to column a (mod W).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "devmath.cuh"
#include "sort.cuh"

namespace b200 {

// x-window geometry, shared by the kernels and by the refinement of the bin order (sort.cu)
template<int NS> struct Sweep2Win {
  static constexpr int W  = NS <= 8 ? 8 : 16;         // columns in the lane window
  static constexpr int S  = NS + 1 <= W ? 2 : 1;      // columns the window moves per position
  static constexpr int G  = 32 / W;                   // lane groups = points per step
  static constexpr int XB = 16;                       // position p covers starts S*p - XB + [0,S)
  static_assert(NS + S - 1 <= W && XB % S == 0 && XB >= NS / 2, "window");
};

template<class T> inline bool sweep2_supported(int ns) {
  return ns >= 2 && ns <= (sizeof(T) == 4 ? 12 : 16);
}

template<class T> struct Sweep2Points {
  const T *xs, *ys;
  const uint32_t *sidx;
  const SweepItem *items;
  uint32_t nitems;
};
constexpr uint32_t kSweep2ItemPoints = 2048;  // most points one warp takes (measured: 1024..4096 within 2 %)

template<class T>
cudaError_t launch_spread2_sweep(int ns, const Sweep2Points<T> &pts, const GridGeom<T> &g, int nc,
                                 const T *coef, const typename CxOf<T>::type *c_in,
                                 typename CxOf<T>::type *fw, cudaStream_t st);
template<class T>
cudaError_t launch_interp2_sweep(int ns, const Sweep2Points<T> &pts, const GridGeom<T> &g, int nc,
                                 const T *coef, typename CxOf<T>::type *c_out,
                                 const typename CxOf<T>::type *fw, cudaStream_t st);

// Orders the points inside every bin (chunks of at most kRefineChunk points) by (x window
// position, y stencil start, index) and gathers the coordinates; sidx is updated in place.
template<class T>
void launch_refine_bins2(int ns, const Packed4<T> *packed, T *xs, T *ys, uint32_t *sidx,
                         const uint32_t *binstart, const uint32_t *chunk_bin,
                         const uint32_t *chunk_off, const uint32_t *nchunks, uint32_t max_chunks,
                         const GridGeom<T> &g, cudaStream_t st);

}  // namespace b200
