// The plan object behind the C ABI: owns every device buffer of one NUFFT plan and drives the
// hot path  setpts (fold-rescale + bin sort) -> spread | interp -> cuFFT -> deconvolve.
//
// Mirrors the guru sequence of the reference (makeplan / setpts / execute / destroy):
//   CPU  include/finufft/makeplan.hpp:317-465, setpts.hpp:107-321, execute.hpp:318-563
//   GPU  src/cuda/makeplan.cu:223-398, src/cuda/setpts.cu:25-75, src/cuda/execute.cu:106-271
// but is a different design: one sort that is bit-identical to the CPU library's, coordinates
// gathered once into sorted order at setpts, fine grid and cuFFT plan kept in the plan, type-2
// zero-padding fused into the amplify kernel, warp-private shared-memory tiles instead of
// shared-memory atomics.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <functional>
#include <memory>
#include <vector>

#include "devmath.cuh"
#include "direct.cuh"
#include "errors.hpp"
#include "gridops.cuh"
#include "sort.cuh"
#include "spreadinterp.cuh"
#include "stage.cuh"
#include "sweep2d.cuh"

namespace b200 {

struct EngineOpts {
  double upsampfac     = 0.0;  // 0 = choose (2.0)
  int spreadinterponly = 0;
  int maxbatch         = 0;    // 0 = min(ntr, 8)
  int device           = 0;
  cudaStream_t stream  = nullptr;
  int modeord          = 0;
  int maxsub           = 1024;  // most points one warp takes from one bin
  int debug            = 0;
  int allow_eps_too_small = 1;
  int sort_radix       = 0;     // 1: stable radix sort (reference permutation on the device)
  int prune            = 1;     // 3D: skip the x,y transforms of the z-planes outside the mode band
  int sweep            = 1;     // 3D float: tube-sweep kernels (0 = generic kernels)
  int stage            = -1;    // two-level strength permutation (stage.cuh): -1 auto, 0 off, 1 on
  int check_sigma      = 0;     // host (finufft_*) entry points apply the CPU feasibility rule
  double group_frac[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // debugging aid: sizes of the point groups
  int auto_sigma       = 0;     // upsampfac = 0 on the host API: choose sigma at setpts (types 1, 2)
  int partition        = 1;     // setpts: 1 partition sort where it pays, 2 always, 0 counting sort
  // bin sort at setpts: 1 sort, 0 keep the user's order and run the point-driven kernels
  // (direct.cuh), 2 = library's choice (this engine: sort).  gpu_sort / spread_sort of the
  // reference's option structs (include/cufinufft_opts.h:11, include/finufft_opts.h:41)
  int sort             = 1;
  // 3D spreadinterponly plans of a sharded transform (slab.cu): the grid handed to execute holds
  // only zwin_n planes of the periodic grid, from global plane zwin_org (sort.cuh, GridGeom)
  int zwin_org = 0, zwin_n = 0;
};

// Callbacks of a pipelined execute (host-pointer plans, capi.cu): the engine announces when it
// is about to consume / has just produced a unit of user data, so the caller can order its
// host<->device copies against the plan's stream.  A point unit is (vector v, group k) with k
// indexing the nchunks groups of consecutive user indices (k = 0 when the plan is not split).
struct ExecHooks {
  std::function<void(int v, int k)> before_points, after_points;
  std::function<void(int b0, int nb)> before_modes, after_modes;
};

template<class T> struct DevBuf {
  T *p     = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf &)            = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void alloc(size_t count);
  void release();
};

// NVTX range for the duration of a scope (visible in Nsight Systems / Compute timelines; a no-op
// without a tool attached).  Header-only NVTX 3: no link dependency.
struct NvtxRange {
  explicit NvtxRange(const char *name);
  ~NvtxRange();
};

struct DeviceGuard {  // make the plan's device current for the duration of a call
  int prev = -1;
  explicit DeviceGuard(int dev);
  ~DeviceGuard();
};

template<class T> class Engine {
 public:
  using C = typename CxOf<T>::type;
  Engine(int type, int dim, const int64_t *nmodes, int iflag, int ntr, double tol,
         const EngineOpts &o);
  ~Engine();
  // all pointers are device pointers on opts.device
  void setpts(int64_t M, const T *x, const T *y, const T *z, int64_t N, const T *s, const T *t,
              const T *u);
  void execute(C *c, C *fk, bool adjoint, const ExecHooks *hooks = nullptr);
  // split the next setpts into k groups of consecutive user indices (types 1 and 2 only)
  void set_point_groups(int k) {
    want_groups_ = k < 1 ? 1 : (k > GridGeom<T>::kMaxGroups ? GridGeom<T>::kMaxGroups : k);
  }
  int point_groups() const { return (int)geom.nchunks; }
  // user indices [group_begin(k), group_begin(k+1)) form group k of the last setpts
  int64_t group_begin(int k) const {
    return k <= 0 ? 0 : (k >= (int)geom.nchunks ? M : (int64_t)geom.gb[k - 1]);
  }
  const Engine<T> *inner() const { return inner_.get(); }  // type 3: the inner type-2 plan

  // ---- introspection (tests, benches) ----
  int type, dim, ntr, sign;
  EngineOpts opts;
  int ns = 0, nc = 0;
  double beta = 0, sigma = 0, tol = 0;
  int64_t ms[3] = {1, 1, 1};
  int64_t nf[3] = {1, 1, 1};
  int64_t M = 0, nk = 0;
  std::vector<T> coef;  // nc x ns, host copy
  GridGeom<T> geom{};
  uint32_t nsub = 0;
  int batch = 1;
  int64_t grid_cells() const { return nf[0] * nf[1] * (geom.zwin_n ? geom.zwin_n : nf[2]); }
  int64_t mode_count() const { return ms[0] * ms[1] * ms[2]; }
  // raw = false: the reference permutation (bins in order, ascending index inside a bin);
  // raw = true: the order the kernels work in
  void copy_sort_to_host(uint32_t *out, bool raw = false) const;
  // 0 counting sort, 1 partition sort, 2 stable radix sort, 3 not sorted (identity order)
  int sort_path() const { return unsorted_ ? 3 : (radix_order_ ? 2 : (part_used_ ? 1 : 0)); }
  bool did_sort() const { return !unsorted_; }
  void copy_phihat_to_host(int d, T *out) const;
  cudaStream_t stream() const { return opts.stream; }
  // stage timing (CUDA events on the plan's stream) and launch accounting for benches
  void enable_profiling(bool on);
  // ms of the last execute: [0] spread or interp, [1] FFT, [2] deconvolve/amplify, [3] total;
  // ms of the last setpts in [4]
  void stage_ms(float out[5]);
  uint64_t launches = 0;  // kernels of this library launched so far (cuFFT/memset excluded)

 private:
  void plan_kernel();
  void plan_grid();
  void sort_points(const T *x, const T *y, const T *z);
  bool use_sweep3(const void *grid) const;
  cudaError_t sweep_run(bool spread, C *c, C *fw, const uint32_t *ix, uint32_t it0,
                        uint32_t nit);
  void build_sweep_items(uint32_t *scan_tmp);
  void run_spread(const C *c, C *fw, int group = -1);
  void run_interp(C *c, const C *fw, int group = -1);
  void spread_path(C *c, C *fk, int fsign, const ExecHooks *hooks);
  void interp_path(C *c, C *fk, int fsign, const ExecHooks *hooks);
  void exec_type3(C *c, C *fk, bool adjoint);
  void setpts_type3(int64_t M, const T *x, const T *y, const T *z, int64_t N, const T *s,
                    const T *t, const T *u);

  bool prof_ = false;
  cudaEvent_t ev_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int order_[4] = {0, 1, 2, 3};
  void mark(int i);
  DevBuf<T> phihat_[3];
  DevBuf<C> fw_;
  cufftHandle fft_ = 0;
  bool have_fft_   = false;
  // pruned 3D FFT: only ms3 of the nf3 z-planes carry wanted (type 1) or non-zero (type 2) data
  // across the x,y transforms, so the 3D FFT is run as one batched 1D FFT along z plus batched
  // 2D FFTs of those planes only (two contiguous ranges: kz >= 0 and kz < 0)
  cufftHandle fftz_ = 0, fftxy_[2] = {0, 0};
  int64_t xy_first_[2] = {0, 0};
  int xy_count_[2]     = {0, 0};
  bool pruned_         = false;
  void fft_grid(C *grid, int nb, int fsign, bool spreading);
  void destroy_fft();
  // point state
  DevBuf<T> xs_, ys_, zs_;
  DevBuf<uint32_t> sidx_, binstart_, sub_bin_, sub_off_;
  DevBuf<SweepItem> items_;  // 3D float sweep kernels: work items, refined bin order in use
  uint32_t nitems_ = 0;
  int want_groups_ = 1;
  std::vector<uint32_t> group_item_, group_sub_;  // first work item / subproblem of every group
  bool swept_      = false;
  bool swept2_     = false;  // 2D sweep kernels (sweep2d.cuh) in use
  // two-level permutation of the strengths (stage.cuh)
  DevBuf<uint32_t> perm1_, perm2_, pinv_;
  DevBuf<C> mid_;
  bool staged_ = false;
  void build_staging();
  bool radix_order_ = false;  // sidx_ is the reference permutation as it stands
  bool part_used_   = false;  // the last setpts took the partition sort (partition.cuh)
  bool pool_held_   = false;
  double tol_req_   = 0;      // the caller's tolerance (tol is what the kernel choice clamped it to)
  bool unsorted_    = false;  // the last setpts kept the user's order (opts.sort = 0)
  DevBuf<T> coef_dev_;        // polynomial table for the point-driven kernels
  // type 3
  DevBuf<T> xp_[3], sp_[3];
  DevBuf<C> prephase_, deconv_, cp_, ck_;
  std::unique_ptr<Engine<T>> inner_;
  T t3C_[3] = {0, 0, 0}, t3D_[3] = {0, 0, 0}, t3h_[3] = {0, 0, 0}, t3gam_[3] = {1, 1, 1};
};

}  // namespace b200
