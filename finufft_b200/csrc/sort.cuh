// setpts on the device: bin keys, stable LSD radix sort of (bin key, point index), bin
// boundaries, gather of coordinates into sorted order, subproblem list.
//
// The permutation produced is bit-identical to the reference CPU library's stable counting
// sort over 16x4x4-cell bins (include/finufft/spread.hpp:459-584 with the bin sizes of
// include/finufft/spreadinterp.hpp:159): bin = i1 + nb1*(i2 + nb2*i3),
// i_d = trunc(fold_rescale(x_d, N_d) * (1/binsize_d)), nb_d = trunc(T(N_d)/binsize_d + 1),
// ties in ascending original index.  A stable radix sort of ascending indices by that key
// is exactly that order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kBinX = 16, kBinY = 4, kBinZ = 4;

template<class T> struct GridGeom {
  int nf[3];      // fine grid size per dim (1 for unused dims)
  T nf_t[3];      // the same, converted to T the way the CPU code does (T(N))
  int nb[3];      // bins per dim
  uint32_t nbins; // sort bins in total = nbins1 * nchunks
  // Host-pointer plans may split the points into nchunks groups of consecutive USER indices
  // (capi.cu: the strengths of group k arrive while group k-1 is being spread).  The sort key is
  // then group-major: key = group * nbins1 + bin, every group a complete bin-sorted point set of
  // its own; kernels recover the bin with a modulo.  Group k holds the user indices
  // [gb[k-1], gb[k]) (gb[-1] = 0); the groups need not be equal (engine.cu: sort_points);
  // unused entries are 0xffffffff.
  uint32_t nbins1  = 0;  // bins of the grid = nb[0]*nb[1]*nb[2]
  uint32_t nchunks = 1;
  static constexpr int kMaxGroups = 8;
  uint32_t gb[kMaxGroups - 1] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu,
                                 0xffffffffu, 0xffffffffu, 0xffffffffu};
  // z window (3D, sharded plans: slab.cu): the grid array holds only zwin_n consecutive planes
  // of the periodic nf[2]-plane grid, starting at global plane zwin_org.  Folding, bins and
  // stencils stay those of the whole grid; a cell of global plane P lives in array plane
  // (P - zwin_org) mod nf[2], which must be < zwin_n for every cell a point of this plan
  // touches.  zwin_n = 0: the array is the whole grid.
  int zwin_org = 0, zwin_n = 0;
};

// group of user index i (call only when g.nchunks > 1)
template<class T>
__host__ __device__ __forceinline__ uint32_t point_group(const GridGeom<T> &g, uint32_t i) {
  uint32_t k = 0;
#pragma unroll
  for (int j = 0; j < GridGeom<T>::kMaxGroups - 1; ++j) k += i >= g.gb[j] ? 1u : 0u;
  return k;
}

// array plane of global plane gz (already wrapped into [0, nf[2])), or -1 outside the window
template<class T> __host__ __device__ __forceinline__ int grid_plane(const GridGeom<T> &g, int gz) {
  if (g.zwin_n == 0) return gz;
  int w = gz - g.zwin_org;
  if (w < 0) w += g.nf[2];
  return w < g.zwin_n ? w : -1;
}

// --- launchers (defined in sort.cu) -------------------------------------------------------
template<class T>
void launch_bin_keys(int dim, const T *x, const T *y, const T *z, uint32_t M,
                     const GridGeom<T> &g, uint32_t *keys, cudaStream_t st);

// Stable sort of keys (only the low `nbits` bits are significant) carrying values; on entry
// values are implicit 0..M-1.  Uses keys_a/keys_b and vals_a/vals_b as ping-pong buffers and
// `hist` (>= 256*kRadixMaxBlocks+1 words).  Returns which buffer (0 = a, 1 = b) holds the
// result.
constexpr uint32_t kRadixMaxBlocks = 148 * 8;
int radix_sort_pairs(uint32_t *keys_a, uint32_t *keys_b, uint32_t *vals_a, uint32_t *vals_b,
                     uint32_t M, int nbits, uint32_t *hist, uint32_t *scan_tmp, cudaStream_t st);

// binstart[b] = first sorted position whose key >= b, b = 0..nbins (binstart[nbins] = M).
void launch_bin_bounds(const uint32_t *sorted_keys, uint32_t M, uint32_t nbins,
                       uint32_t *binstart, cudaStream_t st);

template<class T>
void launch_gather_coords(int dim, const T *x, const T *y, const T *z, const uint32_t *sidx,
                          uint32_t M, T *xs, T *ys, T *zs, cudaStream_t st);

// Exclusive scan of n words; out[n] receives the total (out has n+1 entries). tmp >= n/4096+2.
void exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tmp,
                        cudaStream_t st);

// nsub[b] = ceil(count_b / maxsub);  then (after a scan) the per-subproblem (bin, start) list.
void launch_sub_count(const uint32_t *binstart, uint32_t nbins, uint32_t maxsub, uint32_t *nsub,
                      cudaStream_t st);
void launch_sub_fill(const uint32_t *binstart, const uint32_t *substart, uint32_t nbins,
                     uint32_t maxsub, uint32_t *sub_bin, uint32_t *sub_off, cudaStream_t st);

void launch_iota(uint32_t *v, uint32_t n, cudaStream_t st);

// perm (a copy of the sort order) -> ascending user index inside every bin of at most 1024
// points; *nbig (zero on entry) counts the larger bins, which are left untouched.
void launch_canon_bins(uint32_t *perm, const uint32_t *binstart, uint32_t nbins, uint32_t *nbig,
                       cudaStream_t st);

// --- counting sort by bin (the default setpts path) ----------------------------------------
// Bin counting with warp-aggregated atomics, prefix scan of the counts, placement of every
// index in its bin.  The order inside a bin is the order the atomics happened to return; the
// bins and their ranges are exactly the reference's, and sorting each bin's indices ascending
// gives the reference's stable permutation (include/finufft/spread.hpp:559-581).
template<class T> struct alignas(sizeof(T) * 4) Packed4 {
  T x, y, z, w;
};
// keys[i] = bin of point i, ranks[i] = its arrival number inside the bin, cnt[bin] += 1,
// packed[i] = (x,y,z,0).  cnt must be zero on entry.
template<class T>
void launch_bin_count(int dim, const T *x, const T *y, const T *z, uint32_t M,
                      const GridGeom<T> &g, uint32_t *keys, uint32_t *ranks, uint32_t *cnt,
                      Packed4<T> *packed, cudaStream_t st);
// sidx[binstart[keys[i]] + ranks[i]] = i
void launch_bin_place(const uint32_t *keys, const uint32_t *ranks, const uint32_t *binstart,
                      uint32_t M, uint32_t *sidx, cudaStream_t st);
// coordinates into sorted order from the packed copy (one 16/32-byte read per point)
template<class T>
void launch_gather_packed(int dim, const Packed4<T> *packed, const uint32_t *sidx, uint32_t M,
                          T *xs, T *ys, T *zs, cudaStream_t st);

// --- 3D sweep support (sweep3d.cuh) --------------------------------------------------------
// Orders the points inside every bin by (x window position, y stencil start, index), so that a
// row of bins read front to back is sorted by x window position, and gathers their coordinates.
// The order stays a refinement of the bin order: bins keep their ranges.  Deterministic.
// ns = kernel width (fixes the stencil start rule ceil(X - ns/2)).  sidx is updated in place.
// Work units are chunks of at most kRefineChunk consecutive points of one bin, (chunk_bin,
// chunk_off) with the count on the device, so a bin holding most of the points (clustered
// input) is shared by many warps; the order is refined inside each chunk.
constexpr uint32_t kRefineChunk = 512;
void launch_refine_bins3(int ns, const Packed4<float> *packed, float *xs, float *ys, float *zs,
                         uint32_t *sidx, const uint32_t *binstart, const uint32_t *chunk_bin,
                         const uint32_t *chunk_off, const uint32_t *nchunks, uint32_t max_chunks,
                         const GridGeom<float> &g, cudaStream_t st);

// Work items of the sweep kernels: every row of bins (i2, i3) cut into runs of at most
// `maxpts` consecutive points.  item = {row, first point, one past last point}.
struct SweepItem {
  uint32_t row, qa, qb;
};
void launch_row_item_count(const uint32_t *binstart, uint32_t nrows, uint32_t nb1,
                           uint32_t maxpts, uint32_t *nitems, cudaStream_t st);
void launch_row_item_fill(const uint32_t *binstart, const uint32_t *itemstart, uint32_t nrows,
                          uint32_t nb1, uint32_t maxpts, SweepItem *items, cudaStream_t st);

}  // namespace b200
