// Plan object: makeplan / setpts / execute for types 1 and 2 (type 3 lives in type3.cu).
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include <mutex>

#include <nvtx3/nvToolsExt.h>

#include "partition.cuh"
#include "planmath.hpp"
#include "scratch.hpp"
#include "spreadinterp.cuh"
#include "sweep3d.cuh"

namespace b200 {

// ------------------------------------------------------------------ small utilities
static void cuda_check(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return;
  fprintf(stderr, "[finufft_b200] CUDA error in %s: %s\n", what, cudaGetErrorString(e));
  cudaGetLastError();  // clear
  throw Failure{e == cudaErrorMemoryAllocation ? ERR_ALLOC : ERR_CUDA_FAILURE};
}
#define CU(x) cuda_check((x), #x)

// ------------------------------------------------------------------ scratch pool (scratch.hpp)
namespace {
constexpr int kMaxDevices = 64;
std::mutex pool_mutex;
cudaMemPool_t pools[kMaxDevices] = {};
bool pool_tried[kMaxDevices]     = {};
int pool_users[kMaxDevices]      = {};
}  // namespace
cudaMemPool_t scratch_pool(int device) {
  if (device < 0 || device >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lock(pool_mutex);
  if (!pool_tried[device]) {
    pool_tried[device] = true;
    cudaMemPoolProps props{};
    props.allocType     = cudaMemAllocationTypePinned;
    props.handleTypes   = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id   = device;
    cudaMemPool_t pool  = nullptr;
    if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
      uint64_t keep = ~0ull;  // freed blocks stay cached while plans live; trimmed at release
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      pools[device] = pool;
    }
    cudaGetLastError();
  }
  return pools[device];
}
void scratch_pool_retain(int device) {
  if (device < 0 || device >= kMaxDevices) return;
  std::lock_guard<std::mutex> lock(pool_mutex);
  ++pool_users[device];
}
void scratch_pool_release(int device) {
  if (device < 0 || device >= kMaxDevices) return;
  std::lock_guard<std::mutex> lock(pool_mutex);
  if (--pool_users[device] <= 0) {
    pool_users[device] = 0;
    if (pools[device]) cudaMemPoolTrimTo(pools[device], 0);  // back to the driver
    cudaGetLastError();
  }
}

template<class T> void DevBuf<T>::alloc(size_t count) {
  if (count <= n && p) return;
  release();
  if (count == 0) return;
  CU(cudaMalloc((void **)&p, count * sizeof(T)));
  n = count;
}
template<class T> void DevBuf<T>::release() {
  if (p) cudaFree(p);
  p = nullptr;
  n = 0;
}
template struct DevBuf<float>;
template struct DevBuf<double>;
template struct DevBuf<float2>;
template struct DevBuf<double2>;
template struct DevBuf<uint32_t>;
template struct DevBuf<SweepItem>;

NvtxRange::NvtxRange(const char *name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

DeviceGuard::DeviceGuard(int dev) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) {
    cudaGetLastError();
    throw Failure{ERR_CUDA_FAILURE};
  }
  CU(cudaGetDevice(&prev));
  if (prev != dev) CU(cudaSetDevice(dev));
  else prev = -1;
}
DeviceGuard::~DeviceGuard() {
  if (prev >= 0) cudaSetDevice(prev);
}

template<class T> static cufftType fft_kind();
template<> cufftType fft_kind<float>() { return CUFFT_C2C; }
template<> cufftType fft_kind<double>() { return CUFFT_Z2Z; }
static void fft_exec(cufftHandle h, float2 *d, int dir) {
  if (cufftExecC2C(h, d, d, dir) != CUFFT_SUCCESS) throw Failure{ERR_CUDA_FAILURE};
}
static void fft_exec(cufftHandle h, double2 *d, int dir) {
  if (cufftExecZ2Z(h, d, d, dir) != CUFFT_SUCCESS) throw Failure{ERR_CUDA_FAILURE};
}

// ------------------------------------------------------------------ makeplan
template<class T>
Engine<T>::Engine(int type_, int dim_, const int64_t *nmodes, int iflag, int ntr_, double tol_,
                  const EngineOpts &o)
    : type(type_), dim(dim_), ntr(ntr_), sign(iflag >= 0 ? 1 : -1), opts(o) {
  if (type < 1 || type > 3) throw Failure{ERR_TYPE_NOTVALID};
  if (dim < 1 || dim > 3) throw Failure{ERR_DIM_NOTVALID};
  if (ntr < 1) throw Failure{ERR_NTRANS_NOTVALID};
  DeviceGuard guard(opts.device);
  scratch_pool_retain(opts.device);  // setpts scratch: the library's own pool (scratch.hpp)
  pool_held_ = true;
  tol = tol_req_ = tol_;
  sigma = opts.upsampfac == 0.0 ? 2.0 : opts.upsampfac;
  if (opts.upsampfac != 0.0 || opts.spreadinterponly) opts.auto_sigma = 0;
  batch = opts.maxbatch > 0 ? std::min(opts.maxbatch, ntr) : std::min(ntr, 8);
  if (opts.maxsub < 32) opts.maxsub = 32;
  if (const char *env = getenv("B200_NUFFT_SWEEP")) opts.sweep = atoi(env);  // debugging aids
  if (const char *env = getenv("B200_NUFFT_SORT")) opts.sort_radix = atoi(env) == 2;
  if (const char *env = getenv("B200_NUFFT_STAGE")) opts.stage = atoi(env);
  if (const char *env = getenv("B200_NUFFT_PRUNE")) opts.prune = atoi(env);
  if (const char *env = getenv("B200_NUFFT_PART")) opts.partition = atoi(env);
  if (const char *env = getenv("B200_NUFFT_GROUP_FRACS")) {  // "0.1,0.3,0.3,0.3"
    int j = 0;
    for (const char *q = env; *q && j < 8; ++j) {
      char *end = nullptr;
      opts.group_frac[j] = strtod(q, &end);
      if (end == q) break;
      q = *end == ',' ? end + 1 : end;
    }
  }
  plan_kernel();
  if (type != 3) {
    for (int d = 0; d < dim; ++d) ms[d] = nmodes[d];
    plan_grid();
  }
}

template<class T> void Engine<T>::destroy_fft() {
  if (have_fft_) cufftDestroy(fft_);
  have_fft_ = false;
  if (pruned_) {
    cufftDestroy(fftz_);
    for (int r = 0; r < 2; ++r)
      if (xy_count_[r]) cufftDestroy(fftxy_[r]);
  }
  pruned_      = false;
  xy_count_[0] = xy_count_[1] = 0;
}

// In-place FFT of nb fine grids.  Pruned form (3D): for type 1 the z transform runs first and the
// x,y transforms only on the planes whose kz is kept by the deconvolve step; for type 2 the x,y
// transforms run first, only on the planes amplify filled (the others are zero and stay zero),
// then z.  The reference's CPU path prunes the same way with DUCC0 (src/fft.cpp:299-362).
template<class T> void Engine<T>::fft_grid(C *grid, int nb, int fsign, bool spreading) {
  if (!pruned_) {
    fft_exec(fft_, grid, fsign);
    return;
  }
  const int64_t G = grid_cells(), plane = nf[0] * nf[1];
  for (int i = 0; i < nb; ++i) {
    C *g = grid + (int64_t)i * G;
    if (spreading) fft_exec(fftz_, g, fsign);
    for (int r = 0; r < 2; ++r)
      if (xy_count_[r]) fft_exec(fftxy_[r], g + xy_first_[r] * plane, fsign);
    if (!spreading) fft_exec(fftz_, g, fsign);
  }
}

template<class T> Engine<T>::~Engine() {
  if (pool_held_) {
    // pending frees of this plan's stream must land before the pool can give memory back
    cudaStreamSynchronize(opts.stream);
    scratch_pool_release(opts.device);
  }
  destroy_fft();
  for (auto &e : ev_)
    if (e) cudaEventDestroy(e);
}

template<class T> void Engine<T>::enable_profiling(bool on) {
  DeviceGuard guard(opts.device);
  prof_ = on;
  if (on)
    for (auto &e : ev_)
      if (!e) CU(cudaEventCreate(&e));
}
template<class T> void Engine<T>::mark(int i) {
  if (prof_) cudaEventRecord(ev_[i], opts.stream);
}
template<class T> void Engine<T>::stage_ms(float out[5]) {
  for (int i = 0; i < 5; ++i) out[i] = 0.f;
  if (!prof_) return;
  DeviceGuard guard(opts.device);
  cudaStreamSynchronize(opts.stream);
  // order_ maps {spread/interp, fft, deconv} to the event intervals of the last execute
  float a = 0, b = 0, c = 0;
  if (cudaEventElapsedTime(&a, ev_[0], ev_[1]) != cudaSuccess) a = 0;
  if (cudaEventElapsedTime(&b, ev_[1], ev_[2]) != cudaSuccess) b = 0;
  if (cudaEventElapsedTime(&c, ev_[2], ev_[3]) != cudaSuccess) c = 0;
  const float iv[3] = {a, b, c};
  out[0] = iv[order_[0]];
  out[1] = iv[order_[1]];
  out[2] = iv[order_[2]];
  out[3] = a + b + c;
  float sp = 0;
  if (cudaEventElapsedTime(&sp, ev_[4], ev_[5]) == cudaSuccess) out[4] = sp;
  cudaGetLastError();
}

template<class T> void Engine<T>::plan_kernel() {
  constexpr bool is_f = std::is_same<T, float>::value;
  double tol_used;
  int err = choose_kernel(tol, dim, type, sigma, is_f, opts.allow_eps_too_small != 0, ns, beta,
                          tol_used);
  if (err) throw Failure{err};
  tol = tol_used;
  err = build_horner_table<T>(ns, beta, (T)tol, coef, nc);
  if (err) throw Failure{err};
  if (opts.debug)
    printf("[finufft_b200] type %d dim %d: sigma=%.3g ns=%d beta=%.6g nc=%d\n", type, dim, sigma,
           ns, beta, nc);
}

// fine grid, window Fourier series, cuFFT plan, bin geometry (types 1/2; type 3 calls this
// from setpts once the grid is known)
template<class T> void Engine<T>::plan_grid() {
  int64_t total = 1;
  for (int d = 0; d < dim; ++d) {
    if (opts.spreadinterponly) nf[d] = ms[d];
    else if (type != 3) {
      nf[d] = fine_grid_size(sigma, ms[d], ns);
      if (nf[d] < 0) throw Failure{ERR_MAXNALLOC};
    }
    total *= nf[d];
    if (nf[d] > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
  }
  if (total > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
  for (int d = 0; d < 3; ++d) {
    geom.nf[d]   = (int)nf[d];
    geom.nf_t[d] = (T)nf[d];
  }
  const double bs[3] = {(double)kBinX, (double)kBinY, (double)kBinZ};
  uint64_t nbins = 1;
  for (int d = 0; d < 3; ++d) {
    geom.nb[d] = d < dim ? (int)(int64_t)((double)geom.nf_t[d] / bs[d] + 1) : 1;
    nbins *= (uint64_t)geom.nb[d];
  }
  if (nbins > 0x7fffffffull) throw Failure{ERR_NDATA_NOTVALID};
  geom.nbins = geom.nbins1 = (uint32_t)nbins;
  geom.zwin_org = geom.zwin_n = 0;
  if (opts.zwin_n > 0) {
    if (dim != 3 || !opts.spreadinterponly || opts.zwin_n > nf[2] || opts.zwin_org < 0 ||
        opts.zwin_org >= nf[2])
      throw Failure{ERR_INVALID_ARGUMENT};
    geom.zwin_org = opts.zwin_org;
    geom.zwin_n   = opts.zwin_n == nf[2] ? 0 : opts.zwin_n;
  }
  if (opts.spreadinterponly) return;
  if (type == 3) {  // spread grid only: the inner type-2 plan owns the FFT and the series
    fw_.alloc((size_t)total);
    return;
  }

  cudaStream_t st = opts.stream;
  for (int d = 0; d < dim; ++d) {  // plan-time, O(nf * ns): host, in the reference's arithmetic
    phihat_[d].alloc(nf[d] / 2 + 1);
    int same = -1;
    for (int e = 0; e < d; ++e)
      if (nf[e] == nf[d]) same = e;
    if (same >= 0) {
      CU(cudaMemcpyAsync(phihat_[d].p, phihat_[same].p, sizeof(T) * (nf[d] / 2 + 1),
                         cudaMemcpyDeviceToDevice, st));
      continue;
    }
    std::vector<T> ph;
    fseries_wound<T>(nf[d], ns, nc, coef.data(), ph);
    CU(cudaMemcpyAsync(phihat_[d].p, ph.data(), sizeof(T) * ph.size(), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // ph is a temporary
  }
  fw_.alloc((size_t)total * batch);
  destroy_fft();
  int n[3];
  for (int d = 0; d < dim; ++d) n[d] = (int)nf[dim - 1 - d];  // slowest first
  // pruned form when at most 3/4 of the z-planes are needed (sigma = 2: half of them)
  if (dim == 3 && opts.prune && 4 * ms[2] <= 3 * nf[2] && nf[0] * nf[1] > 1) {
    const int plane = (int)(nf[0] * nf[1]);
    int nz[1] = {(int)nf[2]}, emb[1] = {(int)nf[2]};
    bool ok = cufftPlanMany(&fftz_, 1, nz, emb, plane, 1, emb, plane, 1, fft_kind<T>(), plane) ==
              CUFFT_SUCCESS;
    xy_first_[0] = 0, xy_count_[0] = (int)((ms[2] - 1) / 2 + 1);        // kz = 0 .. kmax
    xy_first_[1] = nf[2] - ms[2] / 2, xy_count_[1] = (int)(ms[2] / 2);  // kz = -ms/2 .. -1
    int nxy[2] = {(int)nf[1], (int)nf[0]};
    int made = 0;
    for (int r = 0; r < 2 && ok; ++r) {
      if (!xy_count_[r]) continue;
      ok = cufftPlanMany(&fftxy_[r], 2, nxy, nullptr, 1, plane, nullptr, 1, plane, fft_kind<T>(),
                         xy_count_[r]) == CUFFT_SUCCESS;
      if (ok) ++made;
    }
    if (!ok) {  // fall back to the plain 3D plan
      if (fftz_) cufftDestroy(fftz_);
      for (int r = 0, k = 0; r < 2 && k < made; ++r)
        if (xy_count_[r]) cufftDestroy(fftxy_[r]), ++k;
      xy_count_[0] = xy_count_[1] = 0;
    } else {
      pruned_ = true;
      bool sok = cufftSetStream(fftz_, st) == CUFFT_SUCCESS;
      for (int r = 0; r < 2; ++r)
        if (xy_count_[r]) sok = sok && cufftSetStream(fftxy_[r], st) == CUFFT_SUCCESS;
      if (!sok) throw Failure{ERR_CUDA_FAILURE};
      return;
    }
  }
  if (cufftPlanMany(&fft_, dim, n, nullptr, 1, (int)total, nullptr, 1, (int)total, fft_kind<T>(),
                    batch) != CUFFT_SUCCESS)
    throw Failure{ERR_CUDA_FAILURE};
  have_fft_ = true;
  if (cufftSetStream(fft_, st) != CUFFT_SUCCESS) throw Failure{ERR_CUDA_FAILURE};
}

// ------------------------------------------------------------------ setpts
static uint32_t sweep2_item_points() {
  const char *e = getenv("B200_SWEEP2_ITEM");
  const long v  = e ? atol(e) : 0;
  return v >= 32 ? (uint32_t)v : kSweep2ItemPoints;
}
// Most points one warp of the 3D sweep kernels takes: a whole row of bins when the point set is
// large (each item pays one window start and one final flush), smaller pieces when the set is
// small (a rank's share of a sharded transform), so that the 148 x 16 resident warps get at
// least ~6 waves of items and the last wave does not dominate.
static uint32_t sweep3_item_points(uint64_t M) {
  const char *e = getenv("B200_SWEEP3_ITEM");
  const long v  = e ? atol(e) : 0;
  if (v >= 32) return (uint32_t)v;
  const uint64_t want = M / (148ull * 16 * 6);
  return (uint32_t)std::min<uint64_t>(kSweepItemPoints, std::max<uint64_t>(1024, want));
}
// refined order inside the bins for the sweep kernels; work units = 512-point chunks of bins
static void refine_impl(int ns, const Packed4<float> *packed, float *xs, float *ys, float *zs,
                        uint32_t *sidx, const uint32_t *binstart, const GridGeom<float> &g,
                        uint64_t M, uint32_t *scan_tmp, cudaStream_t st, int dev) {
  const uint32_t max_chunks =
      (uint32_t)(M / kRefineChunk + std::min<uint64_t>(g.nbins, M));
  Scratch<uint32_t> nch(g.nbins, st, dev), chstart((size_t)g.nbins + 1, st, dev);
  Scratch<uint32_t> chunk_bin(max_chunks, st, dev), chunk_off(max_chunks, st, dev);
  launch_sub_count(binstart, g.nbins, kRefineChunk, nch.p, st);
  exclusive_scan_u32(nch.p, chstart.p, g.nbins, scan_tmp, st);
  launch_sub_fill(binstart, chstart.p, g.nbins, kRefineChunk, chunk_bin.p, chunk_off.p, st);
  launch_refine_bins3(ns, packed, xs, ys, zs, sidx, binstart, chunk_bin.p, chunk_off.p,
                      chstart.p + g.nbins, max_chunks, g, st);
}
static void refine_impl(int, const Packed4<double> *, double *, double *, double *, uint32_t *,
                        const uint32_t *, const GridGeom<double> &, uint64_t, uint32_t *,
                        cudaStream_t, int) {}
// the same for the 2D sweep kernels (any precision)
template<class T>
static void refine2_impl(int ns, const Packed4<T> *packed, T *xs, T *ys, uint32_t *sidx,
                         const uint32_t *binstart, const GridGeom<T> &g, uint64_t M,
                         uint32_t *scan_tmp, cudaStream_t st, int dev) {
  const uint32_t max_chunks =
      (uint32_t)(M / kRefineChunk + std::min<uint64_t>(g.nbins, M));
  Scratch<uint32_t> nch(g.nbins, st, dev), chstart((size_t)g.nbins + 1, st, dev);
  Scratch<uint32_t> chunk_bin(max_chunks, st, dev), chunk_off(max_chunks, st, dev);
  launch_sub_count(binstart, g.nbins, kRefineChunk, nch.p, st);
  exclusive_scan_u32(nch.p, chstart.p, g.nbins, scan_tmp, st);
  launch_sub_fill(binstart, chstart.p, g.nbins, kRefineChunk, chunk_bin.p, chunk_off.p, st);
  launch_refine_bins2<T>(ns, packed, xs, ys, sidx, binstart, chunk_bin.p, chunk_off.p,
                         chstart.p + g.nbins, max_chunks, g, st);
}

template<class T> void Engine<T>::sort_points(const T *x, const T *y, const T *z) {
  cudaStream_t st = opts.stream;
  const int dev   = opts.device;
  const uint32_t m = (uint32_t)M;
  // point groups (sort.cuh): group-major sort keys; needs the counting sort
  {
    uint32_t k = (uint32_t)want_groups_;
    if (type == 3 || M < 2 * (int64_t)k || (uint64_t)geom.nbins1 * k > 0x7fffffffull) k = 1;
    geom.nchunks = k;
    geom.nbins   = geom.nbins1 * k;
    // Group sizes (fractions of the points); measured at C3 through the host API
    // (profiles/r2_e2e_group_layouts.txt).  Type 2: the download of group k runs under the
    // interpolation of group k+1 and the LAST download under nothing, so a smaller last group
    // shortens the step (24.9 -> 23.9 ms).  Type 1: every layout tried (small first group,
    // small last group, 3 / 4 / 5 groups) is equal to or slower than equal groups (21.6 ms).
    double frac[GridGeom<T>::kMaxGroups];
    for (uint32_t j = 0; j < k; ++j) frac[j] = 1.0 / k;
    if (type == 2 && k == 4) frac[0] = 0.30, frac[1] = 0.30, frac[2] = 0.25, frac[3] = 0.15;
    if (opts.group_frac[0] > 0.0)
      for (uint32_t j = 0; j < k; ++j) frac[j] = std::max(opts.group_frac[j], 1e-3);
    double tot = 0.0, acc = 0.0;
    for (uint32_t j = 0; j < k; ++j) tot += frac[j];
    for (int j = 0; j < GridGeom<T>::kMaxGroups - 1; ++j) geom.gb[j] = 0xffffffffu;
    uint32_t prev = 0;
    for (uint32_t j = 0; j + 1 < k; ++j) {
      acc += frac[j];
      uint32_t b = (uint32_t)std::llround((double)M * acc / tot);
      b          = std::min<uint32_t>(std::max<uint32_t>(b, prev + 1), (uint32_t)M - (k - 1 - j));
      geom.gb[j] = prev = b;
    }
  }
  xs_.alloc(M);
  if (dim > 1) ys_.alloc(M);
  if (dim > 2) zs_.alloc(M);
  sidx_.alloc(M);
  unsorted_ = opts.sort == 0 && type != 3;
  if (unsorted_) {
    // the caller asked for no sort: identity permutation (reference indexSort,
    // spreadinterp.hpp:186-191), coordinates kept as given, point-driven kernels at execute
    geom.nchunks = 1, geom.nbins = geom.nbins1;
    for (int j = 0; j < GridGeom<T>::kMaxGroups - 1; ++j) geom.gb[j] = 0xffffffffu;
    swept_ = swept2_ = staged_ = radix_order_ = part_used_ = false;
    nsub = 0, nitems_ = 0;
    group_sub_.assign(2, 0);
    group_item_.assign(2, 0);
    if (M) {
      launch_iota(sidx_.p, m, st);
      CU(cudaMemcpyAsync(xs_.p, x, sizeof(T) * M, cudaMemcpyDeviceToDevice, st));
      if (dim > 1) CU(cudaMemcpyAsync(ys_.p, y, sizeof(T) * M, cudaMemcpyDeviceToDevice, st));
      if (dim > 2) CU(cudaMemcpyAsync(zs_.p, z, sizeof(T) * M, cudaMemcpyDeviceToDevice, st));
    }
    if (!coef_dev_.p) {
      coef_dev_.alloc(coef.size());
      CU(cudaMemcpyAsync(coef_dev_.p, coef.data(), sizeof(T) * coef.size(),
                         cudaMemcpyHostToDevice, st));
    }
    CU(cudaStreamSynchronize(st));
    return;
  }
  binstart_.alloc((size_t)geom.nbins + 1);
  const size_t scan_n = std::max<size_t>(geom.nbins + 1, 256 * (size_t)kRadixMaxBlocks + 1);
  Scratch<uint32_t> scan_tmp(scan_n / 4096 + 8, st, dev);
  // 3D float with a supported width: the sweep kernels want the order inside the bins refined
  swept_ = std::is_same<T, float>::value && dim == 3 && sweep3_supported(ns) && opts.sweep &&
           nf[0] % 2 == 0 && M > 0;
  swept2_ = dim == 2 && sweep2_supported<T>(ns) && opts.sweep && M > 0;
  radix_order_ = opts.sort_radix != 0 && geom.nchunks == 1;

  // default: multi-level partition with sequential streams (partition.cuh); point sets with
  // very dense segments (clustered input) and small or sparse ones take the counting sort below
  bool partitioned = false;
  part_used_       = false;
  if (!radix_order_ && opts.partition) {
    const int cls = swept_ ? kClassSweep3 : (swept2_ ? kClassSweep2 : kClassNone);
    const PartPlan pp = plan_partition((uint64_t)M, geom.nbins, cls, ns, sizeof(T) == 8,
                                       opts.partition > 1);
    if (pp.ok)
      partitioned = partition_sort<T>(dim, x, y, z, m, geom, pp, binstart_.p, xs_.p, ys_.p, zs_.p,
                                      sidx_.p, scan_tmp.p, dev, st);
    part_used_ = partitioned;
  }
  if (partitioned) {
  } else if (!radix_order_) {
    // counting sort: bin counts (warp-aggregated atomics) -> scan -> placement -> gather
    Scratch<uint32_t> keys(M, st, dev), ranks(M, st, dev), cnt(geom.nbins, st, dev);
    Scratch<Packed4<T>> packed(M, st, dev);
    CU(cudaMemsetAsync(cnt.p, 0, sizeof(uint32_t) * geom.nbins, st));
    launch_bin_count<T>(dim, x, y, z, m, geom, keys.p, ranks.p, cnt.p, packed.p, st);
    exclusive_scan_u32(cnt.p, binstart_.p, geom.nbins, scan_tmp.p, st);
    launch_bin_place(keys.p, ranks.p, binstart_.p, m, sidx_.p, st);
    if (swept_)
      refine_impl(ns, packed.p, xs_.p, ys_.p, zs_.p, sidx_.p, binstart_.p, geom, (uint64_t)M,
                  scan_tmp.p, st, dev);
    else if (swept2_)
      refine2_impl<T>(ns, packed.p, xs_.p, ys_.p, sidx_.p, binstart_.p, geom, (uint64_t)M,
                      scan_tmp.p, st, dev);
    else
      launch_gather_packed<T>(dim, packed.p, sidx_.p, m, xs_.p, ys_.p, zs_.p, st);
    CU(cudaGetLastError());
  } else {
    // stable LSD radix sort of (bin key, index): yields the reference permutation directly
    Scratch<uint32_t> keys_a(M, st, dev), keys_b(M, st, dev), vals_b(M, st, dev);
    Scratch<uint32_t> hist(256 * (size_t)kRadixMaxBlocks + 1, st, dev);
    launch_bin_keys<T>(dim, x, y, z, m, geom, keys_a.p, st);
    int nbits = 0;
    while ((1ull << nbits) < (uint64_t)geom.nbins) ++nbits;
    const int which = radix_sort_pairs(keys_a.p, keys_b.p, sidx_.p, vals_b.p, m, nbits, hist.p,
                                       scan_tmp.p, st);
    const uint32_t *sorted_keys = which ? keys_b.p : keys_a.p;
    if (which && M)  // result landed in the scratch value buffer
      CU(cudaMemcpyAsync(sidx_.p, vals_b.p, sizeof(uint32_t) * M, cudaMemcpyDeviceToDevice, st));
    launch_bin_bounds(sorted_keys, m, geom.nbins, binstart_.p, st);
    if (swept_ || swept2_) {
      Scratch<Packed4<T>> packed(M, st, dev);
      Scratch<uint32_t> keys(M, st, dev), ranks(M, st, dev), cnt(geom.nbins, st, dev);  // packs the coordinates
      CU(cudaMemsetAsync(cnt.p, 0, sizeof(uint32_t) * geom.nbins, st));
      launch_bin_count<T>(dim, x, y, z, m, geom, keys.p, ranks.p, cnt.p, packed.p, st);
      if (swept_)
        refine_impl(ns, packed.p, xs_.p, ys_.p, zs_.p, sidx_.p, binstart_.p, geom, (uint64_t)M,
                    scan_tmp.p, st, dev);
      else
        refine2_impl<T>(ns, packed.p, xs_.p, ys_.p, sidx_.p, binstart_.p, geom, (uint64_t)M,
                        scan_tmp.p, st, dev);
    } else
      launch_gather_coords<T>(dim, x, y, z, sidx_.p, m, xs_.p, ys_.p, zs_.p, st);
    CU(cudaGetLastError());
  }

  if (swept_ || swept2_) build_sweep_items(scan_tmp.p);
  build_staging();

  // subproblem list of the generic kernels: every bin in chunks of at most maxsub points
  Scratch<uint32_t> nsubs(geom.nbins, st, dev), substart((size_t)geom.nbins + 1, st, dev);
  launch_sub_count(binstart_.p, geom.nbins, (uint32_t)opts.maxsub, nsubs.p, st);
  exclusive_scan_u32(nsubs.p, substart.p, geom.nbins, scan_tmp.p, st);
  uint32_t total = 0;
  CU(cudaMemcpyAsync(&total, substart.p + geom.nbins, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                     st));
  CU(cudaStreamSynchronize(st));
  nsub = total;
  group_sub_.assign(geom.nchunks + 1, 0);
  group_sub_[geom.nchunks] = total;
  for (uint32_t k = 1; k < geom.nchunks; ++k)
    CU(cudaMemcpyAsync(&group_sub_[k], substart.p + (size_t)k * geom.nbins1, sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, st));
  sub_bin_.alloc(std::max<uint32_t>(nsub, 1));
  sub_off_.alloc(std::max<uint32_t>(nsub, 1));
  if (nsub)
    launch_sub_fill(binstart_.p, substart.p, geom.nbins, (uint32_t)opts.maxsub, sub_bin_.p,
                    sub_off_.p, st);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
}

// Two-level permutation (stage.cuh): worth it once the strength vector is far larger than the
// L2, where every random 8/16-byte access costs a 128-byte line of HBM traffic.
template<class T> void Engine<T>::build_staging() {
  cudaStream_t st = opts.stream;
  const int dev   = opts.device;
  const uint64_t bytes = (uint64_t)M * sizeof(C);
  staged_ = opts.stage > 0;  // opt-in: pays for repeated 2D type-2 executes only
  (void)bytes;
  if (M == 0 || geom.nchunks > 1) staged_ = false;
  if (!staged_) return;
  int shift = stage_shift((uint64_t)M, (int)sizeof(C));
  if (const char *env = getenv("B200_NUFFT_STAGE_SHIFT")) {  // tests: many windows at small M
    shift = std::max(1, atoi(env));
    while ((((uint64_t)M + (1ull << shift) - 1) >> shift) > (uint64_t)kStageMaxWindows) ++shift;
  }
  const uint32_t nunits = ((uint32_t)M + kStageUnit - 1) / kStageUnit;
  const size_t ncnt     = (size_t)kStageMaxWindows * nunits;
  Scratch<uint32_t> counts(ncnt, st, dev), offsets(ncnt + 1, st, dev), tmp(ncnt / 4096 + 8, st, dev);
  perm1_.alloc(M);
  perm2_.alloc(M);
  pinv_.alloc(M);
  mid_.alloc(M);
  build_stage_perms(sidx_.p, (uint32_t)M, shift, counts.p, offsets.p, tmp.p, perm1_.p, perm2_.p,
                    pinv_.p, st);
  CU(cudaGetLastError());
}

template<class T> void Engine<T>::build_sweep_items(uint32_t *scan_tmp) {
  cudaStream_t st = opts.stream;
  const int dev   = opts.device;
  const uint32_t nrows1 = (uint32_t)geom.nb[1] * (uint32_t)geom.nb[2];
  const uint32_t nrows  = nrows1 * geom.nchunks;  // rows are group-major like the bins
  Scratch<uint32_t> nit(nrows, st, dev), itstart((size_t)nrows + 1, st, dev);
  const uint32_t maxpts = swept2_ ? sweep2_item_points() : sweep3_item_points((uint64_t)M);
  launch_row_item_count(binstart_.p, nrows, (uint32_t)geom.nb[0], maxpts, nit.p, st);
  exclusive_scan_u32(nit.p, itstart.p, nrows, scan_tmp, st);
  uint32_t total = 0;
  CU(cudaMemcpyAsync(&total, itstart.p + nrows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  group_item_.assign(geom.nchunks + 1, 0);
  for (uint32_t k = 1; k < geom.nchunks; ++k)
    CU(cudaMemcpyAsync(&group_item_[k], itstart.p + (size_t)k * nrows1, sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  group_item_[geom.nchunks] = total;
  nitems_ = total;
  items_.alloc(std::max<uint32_t>(total, 1));
  if (total)
    launch_row_item_fill(binstart_.p, itstart.p, nrows, (uint32_t)geom.nb[0], maxpts, items_.p,
                         st);
}

template<class T>
void Engine<T>::setpts(int64_t M_, const T *x, const T *y, const T *z, int64_t N, const T *s,
                       const T *t, const T *u) {
  DeviceGuard guard(opts.device);
  if (M_ < 0) throw Failure{ERR_NUM_NU_PTS_INVALID};
  if (M_ > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
  if (type == 3) {
    setpts_type3(M_, x, y, z, N, s, t, u);
    return;
  }
  M = M_;
  if (opts.auto_sigma) {
    // the reference CPU library picks sigma here from the number of points
    // (include/finufft/setpts.hpp:107-161); same candidates, this device's cost model
    const double s_new = choose_sigma(tol_req_, dim, type, std::is_same<T, float>::value, ms,
                                      (double)M);
    if (std::abs(s_new - sigma) > 1e-12) {
      CU(cudaStreamSynchronize(opts.stream));
      sigma = s_new;
      tol   = tol_req_;
      plan_kernel();
      plan_grid();
      coef_dev_.release();
    }
  }
  if (opts.check_sigma) {  // include/finufft/setpts.hpp:29-53
    const double eps  = std::numeric_limits<T>::epsilon();
    const double glen = (double)*std::max_element(nf, nf + dim);
    const bool floor_ = tol <= 0.5 * 0.48 * eps * glen;
    if ((floor_ || least_sigma(tol, dim, ns, eps, glen) > sigma) && !opts.allow_eps_too_small)
      throw Failure{ERR_EPS_TOO_SMALL};
  }
  for (int d = 0; d < dim; ++d)
    if (nf[d] < 2 * ns) throw Failure{ERR_SPREAD_BOX_SMALL};
  NvtxRange range("b200::setpts (fold + bin sort)");
  mark(4);
  sort_points(x, y, z);
  mark(5);
}

// ------------------------------------------------------------------ execute
// The row-sweep kernels (sweep3d.cuh) take 3D single-precision grids whose x lines can be
// addressed in aligned 16-byte pairs; setpts has then refined the bin order for them.
template<class T> bool Engine<T>::use_sweep3(const void *grid) const {
  return swept_ && (reinterpret_cast<uintptr_t>(grid) & 15) == 0;
}
static cudaError_t sweep_impl(bool spread, int ns, const SweepPoints &pts,
                              const GridGeom<float> &g, int nc, const float *coef, float2 *c,
                              float2 *fw, cudaStream_t st) {
  return spread ? launch_spread3_sweep(ns, pts, g, nc, coef, c, fw, st)
                : launch_interp3_sweep(ns, pts, g, nc, coef, c, fw, st);
}
static cudaError_t sweep_impl(bool, int, const SweepPoints &, const GridGeom<double> &, int,
                              const double *, double2 *, double2 *, cudaStream_t) {
  return cudaErrorInvalidValue;
}
template<class T>
cudaError_t Engine<T>::sweep_run(bool spread, C *c, C *fw, const uint32_t *ix, uint32_t it0,
                                 uint32_t nit) {
  SweepPoints sp{reinterpret_cast<const float *>(xs_.p), reinterpret_cast<const float *>(ys_.p),
                 reinterpret_cast<const float *>(zs_.p), ix, items_.p + it0, nit};
  return sweep_impl(spread, ns, sp, geom, nc, coef.data(), c, fw, opts.stream);
}
// group = -1: all points; else the work items / subproblems of that point group only
template<class T> void Engine<T>::run_spread(const C *c, C *fw, int group) {
  if (unsorted_) {
    if (group > 0) return;  // one group only
    CU(launch_direct<T>(true, dim, ns, nc, coef_dev_.p, xs_.p, ys_.p, zs_.p, (uint32_t)M, geom, c,
                        nullptr, fw, opts.stream));
    if (M) ++launches;
    return;
  }
  if (nsub == 0) return;
  uint32_t s0 = 0, s1 = nsub, it0 = 0, it1 = nitems_;
  if (group >= 0) {
    s0 = group_sub_[group], s1 = group_sub_[group + 1];
    if (swept_ || swept2_) it0 = group_item_[group], it1 = group_item_[group + 1];
  }
  if (s1 == s0) return;
  PointSet<T> pts{xs_.p, ys_.p, zs_.p, sidx_.p, binstart_.p, sub_bin_.p + s0, sub_off_.p + s0,
                  s1 - s0, (uint32_t)opts.maxsub};
  cudaError_t e;
  const uint32_t *ix = sidx_.p;
  if (staged_) {  // strengths grouped by window of the user index; kernels then index mid
    launch_stage_in<C>(c, perm2_.p, mid_.p, (uint32_t)M, opts.stream);
    ++launches;
    c  = mid_.p;
    ix = perm1_.p;
    pts.sidx = ix;
  }
  if (use_sweep3(fw))
    e = sweep_run(true, const_cast<C *>(c), fw, ix, it0, it1 - it0);
  else if (swept2_)
    e = launch_spread2_sweep<T>(ns, Sweep2Points<T>{xs_.p, ys_.p, ix, items_.p + it0, it1 - it0},
                                geom, nc, coef.data(), c, fw, opts.stream);
  else if (dim == 1)
    e = launch_spreadinterp<T, 1>(true, ns, pts, geom, nc, coef.data(), c, nullptr, fw, opts.stream);
  else if (dim == 2)
    e = launch_spreadinterp<T, 2>(true, ns, pts, geom, nc, coef.data(), c, nullptr, fw, opts.stream);
  else
    e = launch_spreadinterp<T, 3>(true, ns, pts, geom, nc, coef.data(), c, nullptr, fw, opts.stream);
  if (e == cudaErrorInvalidConfiguration) throw Failure{ERR_INSUFFICIENT_SHMEM};
  CU(e);
  ++launches;
}
template<class T> void Engine<T>::run_interp(C *c, const C *fw, int group) {
  if (unsorted_) {
    if (group > 0) return;
    CU(launch_direct<T>(false, dim, ns, nc, coef_dev_.p, xs_.p, ys_.p, zs_.p, (uint32_t)M, geom,
                        nullptr, c, const_cast<C *>(fw), opts.stream));
    if (M) ++launches;
    return;
  }
  if (nsub == 0) return;
  uint32_t s0 = 0, s1 = nsub, it0 = 0, it1 = nitems_;
  if (group >= 0) {
    s0 = group_sub_[group], s1 = group_sub_[group + 1];
    if (swept_ || swept2_) it0 = group_item_[group], it1 = group_item_[group + 1];
  }
  if (s1 == s0) return;
  PointSet<T> pts{xs_.p, ys_.p, zs_.p, sidx_.p, binstart_.p, sub_bin_.p + s0, sub_off_.p + s0,
                  s1 - s0, (uint32_t)opts.maxsub};
  cudaError_t e;
  C *fwm = const_cast<C *>(fw);
  C *cuser = c;
  const uint32_t *ix = sidx_.p;
  if (staged_) {  // kernels write mid in window order; stage_out scatters inside the windows
    c  = mid_.p;
    ix = perm1_.p;
    pts.sidx = ix;
  }
  if (use_sweep3(fw))
    e = sweep_run(false, c, fwm, ix, it0, it1 - it0);
  else if (swept2_)
    e = launch_interp2_sweep<T>(ns, Sweep2Points<T>{xs_.p, ys_.p, ix, items_.p + it0, it1 - it0},
                                geom, nc, coef.data(), c, fw, opts.stream);
  else if (dim == 1)
    e = launch_spreadinterp<T, 1>(false, ns, pts, geom, nc, coef.data(), nullptr, c, fwm, opts.stream);
  else if (dim == 2)
    e = launch_spreadinterp<T, 2>(false, ns, pts, geom, nc, coef.data(), nullptr, c, fwm, opts.stream);
  else
    e = launch_spreadinterp<T, 3>(false, ns, pts, geom, nc, coef.data(), nullptr, c, fwm, opts.stream);
  if (e == cudaErrorInvalidConfiguration) throw Failure{ERR_INSUFFICIENT_SHMEM};
  CU(e);
  ++launches;
  if (staged_) {
    launch_stage_out<C>(mid_.p, pinv_.p, cuser, (uint32_t)M, opts.stream);
    ++launches;
  }
}

// NU strengths -> modes: spread, FFT, deconvolve (include/finufft/execute.hpp:376-417, type 1)
template<class T>
void Engine<T>::spread_path(C *c, C *fk, int fsign, const ExecHooks *hooks) {
  cudaStream_t st    = opts.stream;
  const int64_t G    = grid_cells(), Nm = mode_count();
  ModeGeom<T> mg;
  for (int d = 0; d < 3; ++d) {
    mg.ms[d] = (int)ms[d];
    mg.nf[d] = (int)nf[d];
    mg.ph[d] = phihat_[d].p;
  }
  mg.modeord = opts.modeord;
  for (int b0 = 0; b0 < ntr; b0 += batch) {
    const int nb = std::min(batch, ntr - b0);
    C *grid      = opts.spreadinterponly ? fk + (int64_t)b0 * Nm : fw_.p;
    order_[0] = 0, order_[1] = 1, order_[2] = 2;
    mark(0);
    {
    NvtxRange range("b200::spread");
    CU(cudaMemsetAsync(grid, 0, sizeof(C) * (size_t)G * nb, st));
    const int ngrp = (int)geom.nchunks;
    for (int i = 0; i < nb; ++i) {
      const C *cv = c + (int64_t)(b0 + i) * M;
      C *gv       = grid + (int64_t)i * G;
      if (ngrp == 1 && !hooks) {
        run_spread(cv, gv);
        continue;
      }
      for (int k = 0; k < ngrp; ++k) {  // group k's strengths may still be on their way
        if (hooks && hooks->before_points) hooks->before_points(b0 + i, k);
        run_spread(cv, gv, ngrp == 1 ? -1 : k);
      }
    }
    }
    mark(1);
    if (opts.spreadinterponly) {
      mark(2);
      mark(3);
      if (hooks && hooks->after_modes) hooks->after_modes(b0, nb);
      continue;
    }
    {
      NvtxRange range("b200::fft (cuFFT)");
      fft_grid(fw_.p, nb, fsign, true);
    }
    mark(2);
    NvtxRange range("b200::deconvolve");
    launch_grid_to_modes<T>(dim, nb, fw_.p, fk + (int64_t)b0 * Nm, mg, st);
    ++launches;
    mark(3);
    if (hooks && hooks->after_modes) hooks->after_modes(b0, nb);
  }
  CU(cudaGetLastError());
}

// modes -> NU values: amplify + zero-pad, FFT, interpolate (type 2)
template<class T>
void Engine<T>::interp_path(C *c, C *fk, int fsign, const ExecHooks *hooks) {
  cudaStream_t st    = opts.stream;
  const int64_t G    = grid_cells(), Nm = mode_count();
  ModeGeom<T> mg;
  for (int d = 0; d < 3; ++d) {
    mg.ms[d] = (int)ms[d];
    mg.nf[d] = (int)nf[d];
    mg.ph[d] = phihat_[d].p;
  }
  mg.modeord = opts.modeord;
  for (int b0 = 0; b0 < ntr; b0 += batch) {
    const int nb  = std::min(batch, ntr - b0);
    const C *grid = fw_.p;
    order_[0] = 2, order_[1] = 1, order_[2] = 0;  // intervals: amplify, fft, interp
    if (hooks && hooks->before_modes) hooks->before_modes(b0, nb);
    mark(0);
    if (opts.spreadinterponly) {
      grid = fk + (int64_t)b0 * Nm;
      mark(1);
      mark(2);
    } else {
      {
        NvtxRange range("b200::amplify");
        launch_modes_to_grid<T>(dim, nb, fk + (int64_t)b0 * Nm, fw_.p, mg, st);
      }
      ++launches;
      mark(1);
      {
        NvtxRange range("b200::fft (cuFFT)");
        fft_grid(fw_.p, nb, fsign, false);
      }
      mark(2);
    }
    const int ngrp = (int)geom.nchunks;
    NvtxRange range("b200::interp");
    for (int i = 0; i < nb; ++i) {
      C *cv       = c + (int64_t)(b0 + i) * M;
      const C *gv = grid + (int64_t)i * G;
      if (ngrp == 1 && !hooks) {
        run_interp(cv, gv);
        continue;
      }
      for (int k = 0; k < ngrp; ++k) {  // group k's values can leave while k+1 is computed
        run_interp(cv, gv, ngrp == 1 ? -1 : k);
        if (hooks && hooks->after_points) hooks->after_points(b0 + i, k);
      }
    }
    mark(3);
  }
  CU(cudaGetLastError());
}

template<class T>
void Engine<T>::execute(C *c, C *fk, bool adjoint, const ExecHooks *hooks) {
  DeviceGuard guard(opts.device);
  if (type == 3) {
    exec_type3(c, fk, adjoint);
    return;
  }
  const bool spreading = (type == 1) != adjoint;
  const int fsign      = adjoint ? -sign : sign;
  if (spreading) spread_path(c, fk, fsign, hooks);
  else interp_path(c, fk, fsign, hooks);
}

template<class T> void Engine<T>::copy_sort_to_host(uint32_t *out, bool raw) const {
  if (M == 0) return;
  cudaStream_t st = opts.stream;
  if (raw || radix_order_ || unsorted_) {  // the device order as it stands
    cudaStreamSynchronize(st);
    cuda_check(cudaMemcpy(out, sidx_.p, sizeof(uint32_t) * M, cudaMemcpyDeviceToHost), "copy sort");
    return;
  }
  // The bins and their ranges are the reference's; inside a bin the device order is (window
  // class, index) or arrival order.  The reference's stable counting sort leaves a bin in
  // ascending index order (include/finufft/spread.hpp:559-581): restore that on the device, on
  // a copy.
  Scratch<uint32_t> perm(M, st, opts.device), nbig(1, st, opts.device);
  cuda_check(cudaMemcpyAsync(perm.p, sidx_.p, sizeof(uint32_t) * M, cudaMemcpyDeviceToDevice, st),
             "copy sort");
  cuda_check(cudaMemsetAsync(nbig.p, 0, sizeof(uint32_t), st), "copy sort");
  launch_canon_bins(perm.p, binstart_.p, geom.nbins, nbig.p, st);
  uint32_t big = 0;
  cuda_check(cudaMemcpyAsync(&big, nbig.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "copy sort");
  cuda_check(cudaMemcpyAsync(out, perm.p, sizeof(uint32_t) * M, cudaMemcpyDeviceToHost, st),
             "copy sort");
  cuda_check(cudaStreamSynchronize(st), "copy sort");
  if (big) {  // bins beyond the kernel's shared-memory capacity (clustered input): host
    std::vector<uint32_t> bs((size_t)geom.nbins + 1);
    cuda_check(cudaMemcpy(bs.data(), binstart_.p, sizeof(uint32_t) * bs.size(),
                          cudaMemcpyDeviceToHost), "copy binstart");
    for (uint32_t b = 0; b < geom.nbins; ++b)
      if (bs[b + 1] - bs[b] > 1024) std::sort(out + bs[b], out + bs[b + 1]);
  }
}
template<class T> void Engine<T>::copy_phihat_to_host(int d, T *out) const {
  cudaStreamSynchronize(opts.stream);
  cuda_check(cudaMemcpy(out, phihat_[d].p, sizeof(T) * (nf[d] / 2 + 1), cudaMemcpyDeviceToHost),
             "copy phihat");
}

template class Engine<float>;
template class Engine<double>;

}  // namespace b200
