// One large 3D type-1 / type-2 transform sharded over the GPUs of one box by z-slabs of the fine
// grid (no reference code exists for this; the reference only runs independent
// plans per device, test/cuda/cufinufft_multigpu_test.cu:29-132).  One process per GPU, NCCL over
// NVLink for the two exchange steps the path really has:
//
//   type 1   spread the rank's points into a WINDOW of the fine grid: the nz = nf3/W planes the
//            rank owns plus ns/2 ghost planes below and ns - ns/2 above (nothing else is
//            allocated; the spreader is the plan's own engine in spread-only mode with a z
//            window, sort.cuh GridGeom)
//            -> ghost planes to the two ring neighbours in ONE grouped send/recv, added there
//            -> batched 2D cuFFT (x, y) of the owned planes
//            -> crop to the ms1 x ms2 wanted modes, divide by phihat1*phihat2, pack by
//               destination                                           [one kernel]
//            -> slab -> pencil transpose: grouped ncclSend/ncclRecv   [all-to-all]
//            -> batched 1D cuFFT along z on the pencils (nf3 x my x ms1)
//            -> keep ms3 modes, divide by phihat3, mode order         [one kernel]
//            => this rank's block fk[:, ylo:yhi, :] of the mode array
//   type 2   the mirror image, ending with interpolation from the window.
//
// Between processes on one NVLink box the exchanges do not go through NCCL at all: every rank maps
// the other ranks' window, pencil and strength buffers (CUDA IPC) and the kernels read / write
// them directly - the ghost planes are ADDED straight from the neighbours' windows, the pack
// kernel STORES the cropped modes into the destination ranks' pencils, the routing kernel stores
// the strengths into the owners' arrays - with three stream barriers per execute (a one-word
// ncclAllReduce each) in place of the send/recv pairs.
//
// Points: `routed = 1` promises that every point already folds into the rank's slab.  Otherwise
// setpts routes them: plane histogram -> all-reduce -> each point goes to the rank that owns its
// plane (all-to-all of coordinates once, of the strengths / values at every execute; both are
// inside the stage timings).  If the points are so clustered that one slab would hold more than
// 1.5x its share and all of them fit a window of at most half the grid, the plan switches to
// REPLICATED-WINDOW mode instead: no routing, every rank spreads its own share of the points
// into a private copy of that window and the copies are summed onto the owning ranks with
// grouped ncclReduce (type 2: owners broadcast their planes of the window) - all GPUs stay busy
// on clustered input.
//
// Grid geometry, mode ordering and deconvolution factors are the single-GPU path's
// (include/finufft/execute.hpp:69-237, makeplan.hpp:39-108), so the result equals the unsharded
// transform up to FFT rounding.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <memory>
#include <vector>

#include "engine.hpp"

namespace b200 {

// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy already in the process if the host
// application, e.g. torch, brought one), so the library has no link-time dependency on it.
int nccl_unique_id(void *out128);  // 0 or ERR_CUDA_FAILURE

template<class T> class SlabPlan {
 public:
  using C = typename CxOf<T>::type;
  // nmodes = {ms1, ms2, ms3}, x fastest like the reference API; uid = the 128 bytes rank 0 got
  // from nccl_unique_id and sent to every rank (world = 1 needs none)
  SlabPlan(int type, const int64_t *nmodes, int iflag, double tol, int rank, int world,
           const void *uid, const EngineOpts &o);
  ~SlabPlan();
  void setpts(int64_t M, const T *x, const T *y, const T *z, int routed);
  // type 1: c = strengths of the rank's own M points (the order given to setpts) ->
  //         fk_block = fk[:, ylo:yhi, :], ms3 x (yhi-ylo) x ms1, x fastest
  // type 2: the reverse
  void execute(C *c, C *fk_block);
  // the full ms3 x ms2 x ms1 mode array on every rank from the blocks
  void gather_modes(const C *fk_block, C *fk_full);
  // this rank's block cut out of a full mode array (type-2 input)
  void slice_modes(const C *fk_full, C *fk_block);
  // [0] spread|interp [1] ghost exchange [2] 2D FFT [3] pack|unpack [4] transpose
  // [5] 1D FFT [6] deconvolve|amplify [7] routing of strengths/values [8] total, last execute;
  // [9] last setpts (incl. routing of the coordinates)
  void stage_ms(float out[10]);
  cudaStream_t stream() const { return st_; }

  int type, rank, world, sign, ns = 0, nc = 0, modeord = 0;
  double sigma = 2.0, tol = 0, beta = 0;
  int64_t ms[3], nf[3];
  int nz = 0, z0 = 0, below = 0, above = 0;
  int64_t ylo = 0, yhi = 0;
  int mode = 0;           // 0 slabs (routed points), 1 replicated window
  int64_t M = 0, Ml = 0;  // user's local points, points this rank spreads
  int win_org = 0, win_n = 0;
  uint64_t launches = 0;  // this file's own kernels (the engine counts its own)
  uint64_t engine_launches() const { return eng_ ? eng_->launches : 0; }
  EngineOpts opts;

 private:
  struct Seg {  // planes [g0, g0+n) of the periodic grid, owned by rank `owner`
    int owner, g0, n;
  };
  void make_engine(int org, int n);
  void exchange_ghosts(bool add);
  void window_collective(bool reduce);
  void transpose(bool to_pencil);
  void route_points(const T *x, const T *y, const T *z);
  void route_values(bool to_owner, C *user);
  // peer-memory path (ranks in separate processes on one NVLink / NVSwitch box): the window,
  // the pencils and the routed strengths of every rank are mapped into every other rank with
  // CUDA IPC, and the exchanges ride inside the kernels as loads / stores over NVLink, ordered
  // by stream barriers (slab.cu).  Falls back to NCCL send/recv when the mapping fails.
  void setup_peers();
  void close_peers();
  void barrier();
  void mark(int i);
  int owner_of_plane(int p) const;

  std::unique_ptr<Engine<T>> eng_;
  int eng_org_ = -1, eng_n_ = -1;
  void *comm_ = nullptr;
  cudaStream_t st_ = nullptr;
  DevBuf<C> win_, own_, send_, pencil_, gprev_, gnext_, croute_, clocal_, gath_;
  C *ownp_ = nullptr;  // first owned plane (inside win_ in slab mode, own_ in replicated mode)
  DevBuf<T> ph_[3], xr_, yr_, zr_, xs_, ys_, zs_;
  DevBuf<uint32_t> order_;
  std::vector<int> zstart_, ystart_;  // world+1 entries each
  std::vector<Seg> segs_;
  std::vector<uint64_t> sendcnt_, recvcnt_, sendoff_, recvoff_;  // points, routing
  bool routed_ = true;
  bool p2p_    = false;
  C *peer_win_[16] = {}, *peer_pencil_[16] = {}, *peer_clocal_[16] = {};
  std::vector<void *> opened_, retired_;  // peers' mappings; own buffers replaced while mapped
  void *published_[3] = {nullptr, nullptr, nullptr};
  bool peers_tried_   = false;
  void grow_exported(DevBuf<C> &b, size_t count);
  std::vector<uint64_t> peer_off_;  // where this rank's routed block starts in rank d's clocal_
  DevBuf<uint32_t> bar_;
  cufftHandle fft2_ = 0, fft1_ = 0;
  bool have2_ = false, have1_ = false;
  cudaEvent_t ev_[12] = {};
};

}  // namespace b200
