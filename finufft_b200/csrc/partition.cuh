// setpts as a multi-level partition with sequential streams (the default bin sort).
//
// What it produces is what the reference CPU library's bin sort produces
// (include/finufft/spread.hpp:459-584): the points grouped by the 16x4x4-cell bin
//   bin = i1 + nb1*(i2 + nb2*i3),  i_d = trunc(fold_rescale(x_d, N_d) * (1/binsize_d)),
// with `binstart` delimiting the bins; the bin of every point is bit-identical to the CPU
// code's.  Inside a bin the points are ordered by (window class, user index): the window class
// is what the sweep kernels key their runs on (sweep3d.cuh, sweep2d.cuh), 0 for the generic
// kernels; sorting a bin's indices ascending gives the reference's stable permutation.
//
// How (B200): every pass reads and writes sequential streams, the 16/32-byte point record
// (x, y, z, index) moves with its key.  Measured on B200 (tools/micro/ticket.cu, 1e8 points):
// a random 16-byte scatter costs 4.6 ms and a random 16-byte gather 12 GB of DRAM reads, but a
// block that ranks a 4096-point tile by a digit of <= 1024 values in shared memory and appends
// every digit's run to a global cursor moves the same data in 1.0 ms with DRAM traffic equal
// to the algorithmic 2.8 GB (the <= 1024 write frontiers merge in the L2).  So:
//
//   1. k_bin_hist      bins counted with warp-aggregated REDs          -> scan -> binstart
//   2. k_seg_prep      segments = 2^ss consecutive bins (~1400 points): cursors, largest segment
//   3. k_part (level A) raw coordinates -> records grouped by the high digit of the segment
//   4. k_part (level B) records -> records grouped by segment          (skipped if <= 1024 segments)
//   5. k_seg_sort      one block per segment: shared-memory counting sort by (bin, window
//                      class), index order inside small groups, coordinates and indices written
//                      as streams into the plan's sorted arrays
//
// A segment must fit the shared-memory sort (kSegCap points); point sets with denser segments
// (clustered input) take the counting-sort path of sort.cu instead (engine.cu decides after
// step 2, the one host synchronisation of setpts).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sort.cuh"

namespace b200 {

constexpr int kPartFanout = 1024;  // digit values one partition pass separates
#ifndef B200_SEGCAP_F32
#define B200_SEGCAP_F32 2048  // 4 resident sort blocks per SM (measured: setpts 5.0 -> 4.9 ms)
#endif
template<class T> struct SegCap {  // most points of one segment
  static constexpr uint32_t value = sizeof(T) == 4 ? B200_SEGCAP_F32 : B200_SEGCAP_F32 / 2;
};

// window classes of the sweep kernels (how the points of one bin are ordered)
enum PartClass : int { kClassNone = 0, kClassSweep3 = 1, kClassSweep2 = 2 };
// number of classes per bin for kernel width ns
int part_class_count(int cls, int ns);

struct PartPlan {
  bool ok       = false;  // geometry fits (else: counting-sort path)
  int ss        = 0;      // log2(bins per segment)
  int sb        = 0;      // log2(segments per level-A bucket), 0 with one level
  int levels    = 1;
  uint32_t nseg = 0, nA = 0;
  int cls = kClassNone, ns = 0, ncls = 1;
};
// Chooses the segment size for M points in nbins bins.  `force` skips the size heuristics
// (tests exercise the path on small inputs).
PartPlan plan_partition(uint64_t M, uint32_t nbins, int cls, int ns, bool is_double, bool force);

// Runs the whole sort.  binstart (nbins+1), xs/ys/zs/sidx (M) are outputs.  Returns false, with
// nothing the caller may rely on written, when a segment holds more than SegCap<T>::value points
// (seen on a 65536-point sample first, so that clustered input leaves before the full histogram);
// the caller then takes the counting-sort path.  Synchronises the stream once or twice.
template<class T>
bool partition_sort(int dim, const T *x, const T *y, const T *z, uint32_t M,
                    const GridGeom<T> &g, const PartPlan &pp, uint32_t *binstart, T *xs, T *ys,
                    T *zs, uint32_t *sidx, uint32_t *scan_tmp, int device, cudaStream_t st);

}  // namespace b200
