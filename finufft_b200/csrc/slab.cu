// z-slab sharded 3D transform (slab.hpp): C++ host code driving the plan's own spread / interp
// engine, cuFFT and NCCL.  Everything between the exchanges is this library's kernels; nothing
// here runs on the host except plan-time bookkeeping.
#include "slab.hpp"

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is loaded with dlopen below

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>

#include "gridops.cuh"
#include "planmath.hpp"
#include "scratch.hpp"

namespace b200 {

// ------------------------------------------------------------------ NCCL, loaded at run time
namespace {
struct NcclApi {
  bool ok = false;
  decltype(&ncclGetUniqueId) GetUniqueId   = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy   = nullptr;
  decltype(&ncclGroupStart) GroupStart     = nullptr;
  decltype(&ncclGroupEnd) GroupEnd         = nullptr;
  decltype(&ncclSend) Send                 = nullptr;
  decltype(&ncclRecv) Recv                 = nullptr;
  decltype(&ncclAllReduce) AllReduce       = nullptr;
  decltype(&ncclReduce) Reduce             = nullptr;
  decltype(&ncclBroadcast) Broadcast       = nullptr;
  decltype(&ncclAllGather) AllGather       = nullptr;
};
NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy the host application already loaded (torch ships its own), else the system one
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
#define B200_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name))
    B200_SYM(GetUniqueId);
    B200_SYM(CommInitRank);
    B200_SYM(CommDestroy);
    B200_SYM(GroupStart);
    B200_SYM(GroupEnd);
    B200_SYM(Send);
    B200_SYM(Recv);
    B200_SYM(AllReduce);
    B200_SYM(Reduce);
    B200_SYM(Broadcast);
    B200_SYM(AllGather);
#undef B200_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart &&
             api.GroupEnd && api.Send && api.Recv && api.AllReduce && api.Reduce &&
             api.Broadcast && api.AllGather;
  });
  return api;
}
void nccl_check(ncclResult_t r, const char *what) {
  if (r == ncclSuccess) return;
  fprintf(stderr, "[finufft_b200] NCCL error %d in %s\n", (int)r, what);
  throw Failure{ERR_CUDA_FAILURE};
}
#define NC(x) nccl_check((x), #x)
void cu(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return;
  fprintf(stderr, "[finufft_b200] CUDA error in %s: %s\n", what, cudaGetErrorString(e));
  cudaGetLastError();
  throw Failure{e == cudaErrorMemoryAllocation ? ERR_ALLOC : ERR_CUDA_FAILURE};
}
#define CU(x) cu((x), #x)

template<class T> cufftType fft_kind();
template<> cufftType fft_kind<float>() { return CUFFT_C2C; }
template<> cufftType fft_kind<double>() { return CUFFT_Z2Z; }
void fft_exec(cufftHandle h, float2 *d, int dir) {
  if (cufftExecC2C(h, d, d, dir) != CUFFT_SUCCESS) throw Failure{ERR_CUDA_FAILURE};
}
void fft_exec(cufftHandle h, double2 *d, int dir) {
  if (cufftExecZ2Z(h, d, d, dir) != CUFFT_SUCCESS) throw Failure{ERR_CUDA_FAILURE};
}
template<class T> ncclDataType_t nccl_real();
template<> ncclDataType_t nccl_real<float>() { return ncclFloat32; }
template<> ncclDataType_t nccl_real<double>() { return ncclFloat64; }

std::vector<int> split_even(int64_t n, int parts) {
  std::vector<int> s(parts + 1, 0);
  const int64_t base = n / parts, extra = n % parts;
  for (int r = 0; r < parts; ++r) s[r + 1] = s[r] + (int)(base + (r < extra ? 1 : 0));
  return s;
}

constexpr int kMaxWorld = 16;
struct SlabGeom {
  int ms[3], nf[3], modeord;
  int nz, my, ylo, world;
  int ystart[kMaxWorld + 1];
};
template<class T> struct SlabPh {
  const T *ph[3];
};

constexpr int kRows = 8, kThreads = 256;

__device__ __forceinline__ int y_owner(const SlabGeom &g, int py) {
  int r = 0;
  while (r + 1 < g.world && py >= g.ystart[r + 1]) ++r;
  return r;
}

// type 1, after the 2D FFT of the owned planes: keep the ms1 x ms2 wanted modes of every plane
// and lay them out by destination rank, each block [plane][y in the rank's range][x]
template<class C>
__global__ void __launch_bounds__(kThreads)
k_slab_pack(const C *__restrict__ own, C *__restrict__ send, SlabGeom g) {
  const int nrows = g.nz * g.ms[1];
  for (int i = 0; i < kRows; ++i) {
    const int row = blockIdx.x * kRows + i;
    if (row >= nrows) return;
    const int py = row % g.ms[1], zl = row / g.ms[1];
    const int ky = mode_freq(py, g.ms[1], g.modeord);
    const int r  = y_owner(g, py);
    const int ylen = g.ystart[r + 1] - g.ystart[r];
    const C *srow  = own + ((int64_t)zl * g.nf[1] + (ky >= 0 ? ky : g.nf[1] + ky)) * g.nf[0];
    C *drow = send + (int64_t)g.nz * g.ms[0] * g.ystart[r] +
              ((int64_t)zl * ylen + (py - g.ystart[r])) * g.ms[0];
    for (int px = threadIdx.x; px < g.ms[0]; px += kThreads) {
      const int kx = mode_freq(px, g.ms[0], g.modeord);
      drow[px]     = srow[kx >= 0 ? kx : g.nf[0] + kx];
    }
  }
}

// Peer-memory forms (slab.hpp): the same crop / zero-pad, but the rows are stored into (type 1)
// or loaded from (type 2) the pencil arrays of the ranks that own their y range, over NVLink.
// Pencil of rank r: [z global][y in r's range][x].
template<class C> struct PeerPtrs {
  C *p[kMaxWorld];
};
template<class C>
__global__ void __launch_bounds__(kThreads)
k_slab_pack_peer(const C *__restrict__ own, PeerPtrs<C> pencil, SlabGeom g, int z0) {
  const int nrows = g.nz * g.ms[1];
  for (int i = 0; i < kRows; ++i) {
    const int row = blockIdx.x * kRows + i;
    if (row >= nrows) return;
    // rows of one destination are consecutive: py fastest inside a plane
    const int py = row % g.ms[1], zl = row / g.ms[1];
    const int ky = mode_freq(py, g.ms[1], g.modeord);
    const int r  = y_owner(g, py);
    const int ylen = g.ystart[r + 1] - g.ystart[r];
    const C *srow  = own + ((int64_t)zl * g.nf[1] + (ky >= 0 ? ky : g.nf[1] + ky)) * g.nf[0];
    C *drow = pencil.p[r] + ((int64_t)(z0 + zl) * ylen + (py - g.ystart[r])) * g.ms[0];
    for (int px = threadIdx.x; px < g.ms[0]; px += kThreads) {
      const int kx = mode_freq(px, g.ms[0], g.modeord);
      drow[px]     = srow[kx >= 0 ? kx : g.nf[0] + kx];
    }
  }
}
template<class C, class T>
__global__ void __launch_bounds__(kThreads)
k_slab_unpack_peer(PeerPtrs<C> pencil, C *__restrict__ own, SlabGeom g, int z0) {
  const int nrows = g.nz * g.nf[1];
  for (int i = 0; i < kRows; ++i) {
    const int row = blockIdx.x * kRows + i;
    if (row >= nrows) return;
    const int cy = row % g.nf[1], zl = row / g.nf[1];
    C *drow = own + (int64_t)row * g.nf[0];
    int ky  = 0;
    if (!cell_freq(cy, g.ms[1], g.nf[1], ky)) {
      for (int cx = threadIdx.x; cx < g.nf[0]; cx += kThreads) drow[cx] = C{(T)0, (T)0};
      continue;
    }
    const int py   = mode_pos(ky, g.ms[1], g.modeord);
    const int r    = y_owner(g, py);
    const int ylen = g.ystart[r + 1] - g.ystart[r];
    const C *srow  = pencil.p[r] + ((int64_t)(z0 + zl) * ylen + (py - g.ystart[r])) * g.ms[0];
    for (int cx = threadIdx.x; cx < g.nf[0]; cx += kThreads) {
      int kx = 0;
      C v    = C{(T)0, (T)0};
      if (cell_freq(cx, g.ms[0], g.nf[0], kx)) v = srow[mode_pos(kx, g.ms[0], g.modeord)];
      drow[cx] = v;
    }
  }
}

// type 2, after the transpose: the mirror image, zero-padding the planes to nf2 x nf1
template<class C, class T>
__global__ void __launch_bounds__(kThreads)
k_slab_unpack(const C *__restrict__ recv, C *__restrict__ own, SlabGeom g) {
  const int nrows = g.nz * g.nf[1];
  for (int i = 0; i < kRows; ++i) {
    const int row = blockIdx.x * kRows + i;
    if (row >= nrows) return;
    const int cy = row % g.nf[1], zl = row / g.nf[1];
    C *drow = own + (int64_t)row * g.nf[0];
    int ky  = 0;
    if (!cell_freq(cy, g.ms[1], g.nf[1], ky)) {
      for (int cx = threadIdx.x; cx < g.nf[0]; cx += kThreads) drow[cx] = C{(T)0, (T)0};
      continue;
    }
    const int py   = mode_pos(ky, g.ms[1], g.modeord);
    const int r    = y_owner(g, py);
    const int ylen = g.ystart[r + 1] - g.ystart[r];
    const C *srow  = recv + (int64_t)g.nz * g.ms[0] * g.ystart[r] +
                    ((int64_t)zl * ylen + (py - g.ystart[r])) * g.ms[0];
    for (int cx = threadIdx.x; cx < g.nf[0]; cx += kThreads) {
      int kx = 0;
      C v    = C{(T)0, (T)0};
      if (cell_freq(cx, g.ms[0], g.nf[0], kx)) v = srow[mode_pos(kx, g.ms[0], g.modeord)];
      drow[cx] = v;
    }
  }
}

// type 1, after the 1D FFT along z of the pencils [z][y local][x]: keep ms3 modes, divide by
// phihat3*phihat2*phihat1 (nested real divisions like gridops.cu), mode order
template<class C, class T>
__global__ void __launch_bounds__(kThreads)
k_slab_deconv(const C *__restrict__ pencil, C *__restrict__ fk, SlabGeom g, SlabPh<T> ph) {
  const int nrows = g.ms[2] * g.my;
  for (int i = 0; i < kRows; ++i) {
    const int row = blockIdx.x * kRows + i;
    if (row >= nrows) return;
    const int yl = row % g.my, pz = row / g.my;
    const int kz = mode_freq(pz, g.ms[2], g.modeord);
    const int ky = mode_freq(g.ylo + yl, g.ms[1], g.modeord);
    T p          = (T)1 / ph.ph[2][kz >= 0 ? kz : -kz];
    p            = p / ph.ph[1][ky >= 0 ? ky : -ky];
    const C *srow = pencil + ((int64_t)(kz >= 0 ? kz : g.nf[2] + kz) * g.my + yl) * g.ms[0];
    C *drow       = fk + (int64_t)row * g.ms[0];
    for (int px = threadIdx.x; px < g.ms[0]; px += kThreads) {
      const int kx = mode_freq(px, g.ms[0], g.modeord);
      const T div  = ph.ph[0][kx >= 0 ? kx : -kx];
      const C v    = srow[px];
      drow[px]     = C{mul_rn(p, v.x) / div, mul_rn(p, v.y) / div};
    }
  }
}

// type 2 start: this rank's block of modes, amplified, zero-padded along z into the pencils
template<class C, class T>
__global__ void __launch_bounds__(kThreads)
k_slab_amplify(const C *__restrict__ fk, C *__restrict__ pencil, SlabGeom g, SlabPh<T> ph) {
  const int nrows = g.nf[2] * g.my;
  for (int i = 0; i < kRows; ++i) {
    const int row = blockIdx.x * kRows + i;
    if (row >= nrows) return;
    const int yl = row % g.my, cz = row / g.my;
    C *drow = pencil + (int64_t)row * g.ms[0];
    int kz  = 0;
    if (!cell_freq(cz, g.ms[2], g.nf[2], kz)) {
      for (int px = threadIdx.x; px < g.ms[0]; px += kThreads) drow[px] = C{(T)0, (T)0};
      continue;
    }
    const int ky = mode_freq(g.ylo + yl, g.ms[1], g.modeord);
    T p          = (T)1 / ph.ph[2][kz >= 0 ? kz : -kz];
    p            = p / ph.ph[1][ky >= 0 ? ky : -ky];
    const C *srow = fk + ((int64_t)mode_pos(kz, g.ms[2], g.modeord) * g.my + yl) * g.ms[0];
    for (int px = threadIdx.x; px < g.ms[0]; px += kThreads) {
      const int kx = mode_freq(px, g.ms[0], g.modeord);
      const T div  = ph.ph[0][kx >= 0 ? kx : -kx];
      const C v    = srow[px];
      drow[px]     = C{mul_rn(p, v.x) / div, mul_rn(p, v.y) / div};
    }
  }
}

// blocks [pz][y of rank r][x] (concatenated by rank) <-> full array [pz][py][px]
template<class C>
__global__ void __launch_bounds__(kThreads)
k_slab_interleave(const C *__restrict__ blocks, C *__restrict__ full, SlabGeom g, int to_full) {
  const int nrows = g.ms[2] * g.ms[1];
  for (int i = 0; i < kRows; ++i) {
    const int row = blockIdx.x * kRows + i;
    if (row >= nrows) return;
    const int py = row % g.ms[1], pz = row / g.ms[1];
    const int r  = y_owner(g, py);
    const int ylen = g.ystart[r + 1] - g.ystart[r];
    const int64_t b = (int64_t)g.ms[2] * g.ms[0] * g.ystart[r] +
                      ((int64_t)pz * ylen + (py - g.ystart[r])) * g.ms[0];
    const int64_t f = (int64_t)row * g.ms[0];
    if (to_full)
      for (int px = threadIdx.x; px < g.ms[0]; px += kThreads) full[f + px] = blocks[b + px];
    else
      for (int px = threadIdx.x; px < g.ms[0]; px += kThreads)
        const_cast<C *>(blocks)[b + px] = full[f + px];
  }
}

template<class T>
__global__ void k_add(T *__restrict__ dst, const T *__restrict__ src, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] += src[i];
}

int blocks_for(int64_t n, int threads, int per_sm = 8) {
  const int64_t want = (n + threads - 1) / threads, cap = 148LL * per_sm;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// ------------------------------------------------------------------ points: planes, routing
// plane of a point = floor of its fold-rescaled z (the fold of the library, devmath.cuh); the
// rounding case fold == nf3 belongs to the last plane, whose owner's window covers its stencil
template<class T> __device__ __forceinline__ int plane_of(T z, T nf3t, int nf3) {
  const int p = (int)fold_rescale<T>(z, nf3t);
  return p < nf3 ? (p < 0 ? 0 : p) : nf3 - 1;
}

// hist[p] += points in plane p; hist[nf3] += points outside [zlo, zhi)
template<class T>
__global__ void __launch_bounds__(256)
k_plane_hist(const T *__restrict__ z, uint32_t M, T nf3t, int nf3, int zlo, int zhi,
             uint32_t *__restrict__ hist) {
  extern __shared__ uint32_t sh[];
  for (int p = threadIdx.x; p <= nf3; p += blockDim.x) sh[p] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  const int lane        = threadIdx.x & 31;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < M; i0 += stride) {
    const uint32_t i = i0 + threadIdx.x;
    const bool live  = i < M;
    const int p      = live ? plane_of<T>(z[i], nf3t, nf3) : -1;
    const uint32_t peers = __match_any_sync(0xffffffffu, p);
    if (live && lane == __ffs(peers) - 1) {
      atomicAdd(&sh[p], (uint32_t)__popc(peers));
      if (p < zlo || p >= zhi) atomicAdd(&sh[nf3], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p <= nf3; p += blockDim.x)
    if (sh[p]) atomicAdd(&hist[p], sh[p]);
}

struct RouteGeom {
  int world, nf3;
  int zstart[kMaxWorld + 1];
};
__device__ __forceinline__ int z_owner(const RouteGeom &g, int p) {
  int r = 0;
  while (r + 1 < g.world && p >= g.zstart[r + 1]) ++r;
  return r;
}
constexpr uint32_t kRouteChunk = 2048;  // points per block

// counts[d * nblk + b] = points of block b's chunk that go to rank d
template<class T>
__global__ void __launch_bounds__(256)
k_route_count(const T *__restrict__ z, uint32_t M, T nf3t, RouteGeom g, uint32_t nblk,
              uint32_t *__restrict__ counts) {
  __shared__ uint32_t cnt[kMaxWorld];
  if (threadIdx.x < kMaxWorld) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kRouteChunk;
  const int lane      = threadIdx.x & 31;
  for (uint32_t k = 0; k < kRouteChunk; k += 256) {
    const uint32_t i = base + k + threadIdx.x;
    const int d      = i < M ? z_owner(g, plane_of<T>(z[i], nf3t, g.nf3)) : -1;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (d >= 0 && lane == __ffs(peers) - 1) atomicAdd(&cnt[d], (uint32_t)__popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < g.world) counts[threadIdx.x * nblk + blockIdx.x] = cnt[threadIdx.x];
}

// stable placement: send position of point i = offsets[d * nblk + b] + rank of i among the
// block's earlier points with the same destination; coordinates written in send order
template<class T>
__global__ void __launch_bounds__(256)
k_route_fill(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
             uint32_t M, T nf3t, RouteGeom g, uint32_t nblk, const uint32_t *__restrict__ offsets,
             uint32_t *__restrict__ order, T *__restrict__ xs, T *__restrict__ ys,
             T *__restrict__ zs) {
  __shared__ uint32_t base_[kMaxWorld], wcnt[8][kMaxWorld];
  if (threadIdx.x < g.world) base_[threadIdx.x] = offsets[threadIdx.x * nblk + blockIdx.x];
  const uint32_t base = blockIdx.x * kRouteChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t k = 0; k < kRouteChunk; k += 256) {
    if (threadIdx.x < 8 * kMaxWorld) (&wcnt[0][0])[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = base + k + threadIdx.x;
    T zi             = (T)0;
    int d            = -1;
    if (i < M) {
      zi = z[i];
      d  = z_owner(g, plane_of<T>(zi, nf3t, g.nf3));
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank  = (uint32_t)__popc(peers & ((1u << lane) - 1u));
    if (d >= 0 && lane == __ffs(peers) - 1) wcnt[warp][d] = (uint32_t)__popc(peers);
    __syncthreads();
    if (d >= 0) {
      uint32_t pos = base_[d] + rank;
      for (int w = 0; w < warp; ++w) pos += wcnt[w][d];
      order[pos] = i;
      xs[pos]    = x[i];
      ys[pos]    = y[i];
      zs[pos]    = zi;
    }
    __syncthreads();
    if (threadIdx.x < g.world) {
      uint32_t tot = 0;
      for (int w = 0; w < 8; ++w) tot += wcnt[w][threadIdx.x];
      base_[threadIdx.x] += tot;
    }
    __syncthreads();  // wcnt is cleared at the top of the next round
  }
}

// Routing over peer memory: send position pos (destination-major, k_route_fill) belongs to rank
// d = the block of `sendoff` holding it; its slot in d's array is peer_off[d] + pos - sendoff[d].
// push: strengths of the caller's points are stored into the owners' arrays (type 1);
// pull: interpolated values are loaded from them (type 2).
struct RouteTable {
  int world;
  uint32_t sendoff[kMaxWorld + 1];
  uint32_t peer_off[kMaxWorld];
};
template<class C, bool PUSH>
__global__ void k_route_peer(C *__restrict__ user, const uint32_t *__restrict__ order,
                             PeerPtrs<C> owner, RouteTable t, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += stride) {
    int d = 0;
    while (d + 1 < t.world && pos >= t.sendoff[d + 1]) ++d;
    C *slot = owner.p[d] + (t.peer_off[d] + (pos - t.sendoff[d]));
    if (PUSH) *slot = user[order[pos]];
    else user[order[pos]] = *slot;
  }
}

template<class C>
__global__ void k_gather_by(const C *__restrict__ src, const uint32_t *__restrict__ order,
                            C *__restrict__ dst, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = src[order[i]];
}
template<class C>
__global__ void k_scatter_by(const C *__restrict__ src, const uint32_t *__restrict__ order,
                             C *__restrict__ dst, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[order[i]] = src[i];
}

}  // namespace

int nccl_unique_id(void *out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
  NcclApi &api = nccl();
  if (!api.ok || !out128) return ERR_CUDA_FAILURE;
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) return ERR_CUDA_FAILURE;
  std::memcpy(out128, &id, sizeof(id));
  return 0;
}

// ------------------------------------------------------------------ plan
template<class T>
SlabPlan<T>::SlabPlan(int type_, const int64_t *nmodes, int iflag, double tol_, int rank_,
                      int world_, const void *uid, const EngineOpts &o)
    : type(type_), rank(rank_), world(world_), sign(iflag >= 0 ? 1 : -1), opts(o) {
  constexpr bool is_f = std::is_same<T, float>::value;
  if (type != 1 && type != 2) throw Failure{ERR_TYPE_NOTVALID};
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || !nmodes)
    throw Failure{ERR_INVALID_ARGUMENT};
  DeviceGuard guard(opts.device);
  st_     = opts.stream;
  modeord = opts.modeord;
  sigma   = opts.upsampfac == 0.0 ? 2.0 : opts.upsampfac;
  double tol_used = tol_;
  int err = choose_kernel(tol_, 3, type, sigma, is_f, true, ns, beta, tol_used);
  if (err) throw Failure{err};
  tol = tol_used;
  std::vector<T> coef;
  err = build_horner_table<T>(ns, beta, (T)tol, coef, nc);
  if (err) throw Failure{err};
  int64_t total = 1;
  for (int d = 0; d < 3; ++d) {
    ms[d] = nmodes[d];
    if (ms[d] < 1) throw Failure{ERR_NDATA_NOTVALID};
    nf[d] = fine_grid_size(sigma, ms[d], ns);
    if (nf[d] < 0) throw Failure{ERR_MAXNALLOC};
    if (nf[d] < 2 * ns) throw Failure{ERR_SPREAD_BOX_SMALL};
    total *= nf[d];
  }
  if (total / world + (int64_t)ns * nf[0] * nf[1] > std::numeric_limits<int32_t>::max())
    throw Failure{ERR_NDATA_NOTVALID};
  zstart_ = split_even(nf[2], world);
  ystart_ = split_even(ms[1], world);
  z0  = zstart_[rank];
  nz  = zstart_[rank + 1] - z0;
  ylo = ystart_[rank];
  yhi = ystart_[rank + 1];
  below = world > 1 ? ns / 2 : 0;
  above = world > 1 ? ns - ns / 2 : 0;
  if (world > 1) {
    // ghosts reach the immediate neighbours only; the window must not meet itself
    for (int r = 0; r < world; ++r)
      if (zstart_[r + 1] - zstart_[r] < std::max(below, above)) throw Failure{ERR_INVALID_ARGUMENT};
    if (nz + ns > nf[2]) throw Failure{ERR_INVALID_ARGUMENT};
  }
  for (int d = 0; d < 3; ++d) {
    std::vector<T> ph;
    fseries_wound<T>(nf[d], ns, nc, coef.data(), ph);
    ph_[d].alloc(ph.size());
    CU(cudaMemcpy(ph_[d].p, ph.data(), sizeof(T) * ph.size(), cudaMemcpyHostToDevice));
  }
  const int64_t plane = nf[0] * nf[1], my = yhi - ylo;
  send_.alloc((size_t)std::max<int64_t>(1, (int64_t)nz * ms[1] * ms[0]));
  pencil_.alloc((size_t)std::max<int64_t>(1, nf[2] * my * ms[0]));
  {
    int nxy[2] = {(int)nf[1], (int)nf[0]};
    if (cufftPlanMany(&fft2_, 2, nxy, nullptr, 1, (int)plane, nullptr, 1, (int)plane,
                      fft_kind<T>(), nz) != CUFFT_SUCCESS)
      throw Failure{ERR_CUDA_FAILURE};
    have2_ = true;
    if (cufftSetStream(fft2_, st_) != CUFFT_SUCCESS) throw Failure{ERR_CUDA_FAILURE};
    if (my > 0) {
      int n3[1] = {(int)nf[2]}, emb[1] = {(int)nf[2]};
      const int lines = (int)(my * ms[0]);
      if (cufftPlanMany(&fft1_, 1, n3, emb, lines, 1, emb, lines, 1, fft_kind<T>(), lines) !=
          CUFFT_SUCCESS)
        throw Failure{ERR_CUDA_FAILURE};
      have1_ = true;
      if (cufftSetStream(fft1_, st_) != CUFFT_SUCCESS) throw Failure{ERR_CUDA_FAILURE};
    }
  }
  for (auto &e : ev_) CU(cudaEventCreate(&e));
  if (world > 1) {
    NcclApi &api = nccl();
    if (!api.ok || !uid) throw Failure{ERR_CUDA_FAILURE};
    ncclUniqueId id;
    std::memcpy(&id, uid, sizeof(id));
    ncclComm_t comm = nullptr;
    NC(api.CommInitRank(&comm, world, id, rank));
    comm_ = comm;
    gprev_.alloc((size_t)above * plane);
    gnext_.alloc((size_t)below * plane);
  }
  make_engine((int)((z0 - below + nf[2]) % nf[2]), nz + below + above);
}

template<class T> SlabPlan<T>::~SlabPlan() {
  cudaStreamSynchronize(st_);
  if (peers_tried_) {
    // no rank may free a buffer another rank still has mapped: unmap, then meet (collective)
    close_peers();
    try {
      barrier();
    } catch (...) {
    }
    cudaStreamSynchronize(st_);
  }
  for (void *q : retired_) cudaFree(q);
  eng_.reset();
  if (comm_) nccl().CommDestroy((ncclComm_t)comm_);
  if (have2_) cufftDestroy(fft2_);
  if (have1_) cufftDestroy(fft1_);
  for (auto &e : ev_)
    if (e) cudaEventDestroy(e);
}

template<class T> int SlabPlan<T>::owner_of_plane(int p) const {
  int r = 0;
  while (r + 1 < world && p >= zstart_[r + 1]) ++r;
  return r;
}

// A buffer other ranks may have mapped is never freed while mapped: when it has to grow, the old
// allocation is parked until the next collective unmapping (setup_peers).
template<class T> void SlabPlan<T>::grow_exported(DevBuf<C> &b, size_t count) {
  if (count <= b.n && b.p) return;
  if (b.p && peers_tried_) {
    retired_.push_back(b.p);
    b.p = nullptr;
    b.n = 0;
  }
  b.alloc(count);
}

// the spread / interp engine on a window of `n` planes from global plane `org`
template<class T> void SlabPlan<T>::make_engine(int org, int n) {
  if (n >= nf[2]) org = 0, n = (int)nf[2];
  win_org = org;
  win_n   = n;
  grow_exported(win_, (size_t)n * nf[0] * nf[1]);
  if (eng_ && eng_org_ == org && eng_n_ == n) return;
  EngineOpts eo       = opts;
  eo.spreadinterponly = 1;
  eo.upsampfac        = sigma;
  eo.stream           = st_;
  eo.zwin_org         = org;
  eo.zwin_n           = n;
  eng_.reset();
  eng_.reset(new Engine<T>(type, 3, nf, sign, 1, tol, eo));
  if (eng_->ns != ns) throw Failure{ERR_UNKNOWN_EXCEPTION};
  eng_org_ = org;
  eng_n_   = n;
}

template<class T> void SlabPlan<T>::mark(int i) { cudaEventRecord(ev_[i], st_); }

template<class T> void SlabPlan<T>::stage_ms(float out[10]) {
  for (int i = 0; i < 10; ++i) out[i] = 0.f;
  DeviceGuard guard(opts.device);
  cudaStreamSynchronize(st_);
  float iv[8] = {};
  for (int i = 0; i < 8; ++i)
    if (cudaEventElapsedTime(&iv[i], ev_[i], ev_[i + 1]) != cudaSuccess) iv[i] = 0;
  // type 1 runs route, spread, ghosts, fft2, pack, transpose, fft1, deconvolve; type 2 backwards
  static const int slot1[8] = {7, 0, 1, 2, 3, 4, 5, 6}, slot2[8] = {6, 5, 4, 3, 2, 1, 0, 7};
  for (int i = 0; i < 8; ++i) {
    out[(type == 1 ? slot1 : slot2)[i]] = iv[i];
    out[8] += iv[i];
  }
  float sp = 0;
  if (cudaEventElapsedTime(&sp, ev_[9], ev_[10]) == cudaSuccess) out[9] = sp;
  cudaGetLastError();
}

// ------------------------------------------------------------------ setpts
template<class T>
void SlabPlan<T>::route_points(const T *x, const T *y, const T *z) {
  NcclApi &api     = nccl();
  ncclComm_t comm  = (ncclComm_t)comm_;
  const uint32_t m = (uint32_t)M;
  const int dev    = opts.device;
  RouteGeom rg{};
  rg.world = world;
  rg.nf3   = (int)nf[2];
  for (int r = 0; r <= world; ++r) rg.zstart[r] = zstart_[r];
  const uint32_t nblk = std::max<uint32_t>(1, (m + kRouteChunk - 1) / kRouteChunk);
  const size_t ncnt   = (size_t)world * nblk;
  Scratch<uint32_t> counts(ncnt, st_, dev), offsets(ncnt + 1, st_, dev),
      tmp(ncnt / 4096 + 8, st_, dev), mine(world, st_, dev), all((size_t)world * world, st_, dev);
  k_route_count<T><<<nblk, 256, 0, st_>>>(z, m, (T)nf[2], rg, nblk, counts.p);
  exclusive_scan_u32(counts.p, offsets.p, (uint32_t)ncnt, tmp.p, st_);
  xs_.alloc(std::max<size_t>(1, M));
  ys_.alloc(std::max<size_t>(1, M));
  zs_.alloc(std::max<size_t>(1, M));
  order_.alloc(std::max<size_t>(1, M));
  k_route_fill<T><<<nblk, 256, 0, st_>>>(x, y, z, m, (T)nf[2], rg, nblk, offsets.p, order_.p,
                                        xs_.p, ys_.p, zs_.p);
  launches += 2;
  CU(cudaGetLastError());
  // send offsets per destination = offsets[d * nblk]; counts of every rank to every rank
  std::vector<uint32_t> off(world + 1, 0);
  CU(cudaMemcpy2DAsync(off.data(), sizeof(uint32_t), offsets.p, sizeof(uint32_t) * nblk,
                       sizeof(uint32_t), world, cudaMemcpyDeviceToHost, st_));
  CU(cudaStreamSynchronize(st_));
  off[world] = m;
  std::vector<uint32_t> cnt(world);
  for (int d = 0; d < world; ++d) cnt[d] = off[d + 1] - off[d];
  CU(cudaMemcpyAsync(mine.p, cnt.data(), sizeof(uint32_t) * world, cudaMemcpyHostToDevice, st_));
  NC(api.AllGather(mine.p, all.p, world, ncclUint32, comm, st_));
  std::vector<uint32_t> mat((size_t)world * world);
  CU(cudaMemcpyAsync(mat.data(), all.p, sizeof(uint32_t) * mat.size(), cudaMemcpyDeviceToHost,
                     st_));
  CU(cudaStreamSynchronize(st_));
  sendcnt_.assign(world, 0), recvcnt_.assign(world, 0);
  sendoff_.assign(world + 1, 0), recvoff_.assign(world + 1, 0);
  for (int r = 0; r < world; ++r) {
    sendcnt_[r]     = cnt[r];
    recvcnt_[r]     = mat[(size_t)r * world + rank];
    sendoff_[r + 1] = sendoff_[r] + sendcnt_[r];
    recvoff_[r + 1] = recvoff_[r] + recvcnt_[r];
  }
  Ml = (int64_t)recvoff_[world];
  if (Ml > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
  // where this rank's block starts inside rank d's received array: the blocks of ranks < rank
  peer_off_.assign(world, 0);
  for (int d = 0; d < world; ++d)
    for (int s = 0; s < rank; ++s) peer_off_[d] += mat[(size_t)s * world + d];
  xr_.alloc(std::max<size_t>(1, Ml));
  yr_.alloc(std::max<size_t>(1, Ml));
  zr_.alloc(std::max<size_t>(1, Ml));
  const T *src[3] = {xs_.p, ys_.p, zs_.p};
  T *dst[3]       = {xr_.p, yr_.p, zr_.p};
  NC(api.GroupStart());
  for (int a = 0; a < 3; ++a)
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      if (sendcnt_[r])
        NC(api.Send(src[a] + sendoff_[r], sendcnt_[r] * sizeof(T), ncclChar, r, comm, st_));
      if (recvcnt_[r])
        NC(api.Recv(dst[a] + recvoff_[r], recvcnt_[r] * sizeof(T), ncclChar, r, comm, st_));
    }
  NC(api.GroupEnd());
  for (int a = 0; a < 3; ++a)
    if (sendcnt_[rank])
      CU(cudaMemcpyAsync(dst[a] + recvoff_[rank], src[a] + sendoff_[rank],
                         sendcnt_[rank] * sizeof(T), cudaMemcpyDeviceToDevice, st_));
  croute_.alloc(std::max<size_t>(1, M));
  grow_exported(clocal_, std::max<size_t>(1, Ml));
}

template<class T>
void SlabPlan<T>::setpts(int64_t M_, const T *x, const T *y, const T *z, int routed) {
  DeviceGuard guard(opts.device);
  if (M_ < 0) throw Failure{ERR_NUM_NU_PTS_INVALID};
  if (M_ > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
  M = M_;
  NvtxRange range("b200::slab setpts (route + sort)");
  mark(9);
  const int nf3 = (int)nf[2];
  const int dev = opts.device;
  // planes the points fold into: local histogram, summed over the ranks
  std::vector<uint32_t> hist((size_t)nf3 + 1, 0);
  {
    Scratch<uint32_t> h((size_t)nf3 + 1, st_, dev);
    CU(cudaMemsetAsync(h.p, 0, sizeof(uint32_t) * (nf3 + 1), st_));
    if (M) {
      const size_t shmem = sizeof(uint32_t) * (nf3 + 1);
      if (shmem > 48 * 1024)
        CU(cudaFuncSetAttribute(k_plane_hist<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)shmem));
      k_plane_hist<T><<<blocks_for(M, 256, 4), 256, shmem, st_>>>(z, (uint32_t)M, (T)nf3, nf3, z0,
                                                                 z0 + nz, h.p);
      ++launches;
      CU(cudaGetLastError());
    }
    if (world > 1)
      NC(nccl().AllReduce(h.p, h.p, (size_t)nf3 + 1, ncclUint32, ncclSum, (ncclComm_t)comm_, st_));
    CU(cudaMemcpyAsync(hist.data(), h.p, sizeof(uint32_t) * (nf3 + 1), cudaMemcpyDeviceToHost,
                       st_));
    CU(cudaStreamSynchronize(st_));
  }
  // every rank takes the same decision from the same global histogram
  if (routed && hist[nf3]) throw Failure{ERR_INVALID_ARGUMENT};  // a point outside its slab
  routed_ = routed != 0 || world == 1;
  mode    = 0;
  uint64_t total = 0, worst = 0;
  for (int r = 0; r < world; ++r) {
    uint64_t load = 0;
    for (int p = zstart_[r]; p < zstart_[r + 1]; ++p) load += hist[p];
    total += load;
    worst = std::max(worst, load);
  }
  int worg = 0, wn = nf3;
  if (world > 1 && !routed && total > 0 && 2 * worst * world > 3 * total) {
    // clustered: smallest periodic window of planes that holds every point's stencil =
    // complement of the longest run of empty planes
    int best_len = 0, best_end = 0, run = 0;
    for (int k = 0; k < 2 * nf3; ++k) {
      const int p = k % nf3;
      if (hist[p] == 0) {
        ++run;
        if (run > best_len && run <= nf3) best_len = run, best_end = p;
      } else
        run = 0;
    }
    const int first = (best_end + 1) % nf3;  // first occupied plane after the gap
    const int occ   = nf3 - best_len;
    wn   = occ + ns;
    worg = ((first - ns / 2) % nf3 + nf3) % nf3;
    if (2 * wn <= nf3) mode = 1;
  }
  const T *xl = x, *yl = y, *zl = z;
  Ml = M;
  segs_.clear();
  if (mode == 1) {
    // replicated window: every rank keeps its own points; owners are found per plane range
    make_engine(worg, wn);
    own_.alloc((size_t)nz * nf[0] * nf[1]);
    ownp_ = own_.p;
    for (int r = 0; r < world; ++r)
      for (int piece = 0; piece < 2; ++piece) {  // the window as one or two straight intervals
        const int a = piece == 0 ? worg : 0;
        const int b = piece == 0 ? std::min(worg + wn, nf3) : worg + wn - nf3;
        const int lo = std::max(a, zstart_[r]), hi = std::min(b, zstart_[r + 1]);
        if (hi > lo) segs_.push_back(Seg{r, lo, hi - lo});
      }
  } else {
    make_engine((int)((z0 - below + nf3) % nf3), nz + below + above);
    ownp_ = win_.p + (size_t)below * nf[0] * nf[1];
    if (!routed_) {
      route_points(x, y, z);
      xl = xr_.p, yl = yr_.p, zl = zr_.p;
    }
  }
  eng_->setpts(Ml, xl, yl, zl, 0, nullptr, nullptr, nullptr);
  if (world > 1) setup_peers();
  mark(10);
}

// ------------------------------------------------------------------ peer memory (CUDA IPC)
template<class T> void SlabPlan<T>::close_peers() {
  for (void *q : opened_) cudaIpcCloseMemHandle(q);
  opened_.clear();
  for (int r = 0; r < kMaxWorld; ++r) peer_win_[r] = peer_pencil_[r] = peer_clocal_[r] = nullptr;
  cudaGetLastError();
  p2p_ = false;
}

// Every rank publishes the IPC handles of its window, pencil and routed-strength buffers and maps
// the others'.  Collective (called from setpts).  Any failure on any rank (threads of one process,
// no peer access, B200_NUFFT_SLAB_P2P=0) leaves the NCCL send/recv path in charge on all ranks.
template<class T> void SlabPlan<T>::setup_peers() {
  struct Pub {
    cudaIpcMemHandle_t h[3];
    uint32_t have[3];
    uint32_t ok;
  };
  NcclApi &api    = nccl();
  ncclComm_t comm = (ncclComm_t)comm_;
  bar_.alloc(1);
  void *bufs[3] = {win_.p, pencil_.p, (mode == 0 && !routed_) ? (void *)clocal_.p : nullptr};
  {  // mapping is slow (milliseconds): keep the existing one unless some rank's buffers moved
    uint32_t changed = !peers_tried_;
    for (int k = 0; k < 3; ++k) changed |= bufs[k] != published_[k];
    Scratch<uint32_t> flag(1, st_, opts.device);
    CU(cudaMemcpyAsync(flag.p, &changed, sizeof(uint32_t), cudaMemcpyHostToDevice, st_));
    NC(api.AllReduce(flag.p, flag.p, 1, ncclUint32, ncclMax, comm, st_));
    CU(cudaMemcpyAsync(&changed, flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st_));
    CU(cudaStreamSynchronize(st_));
    if (!changed) return;
  }
  // unmap everywhere, then meet, and only then free what was replaced while mapped
  close_peers();
  barrier();
  CU(cudaStreamSynchronize(st_));
  for (void *q : retired_) cudaFree(q);
  retired_.clear();
  peers_tried_ = true;
  for (int k = 0; k < 3; ++k) published_[k] = bufs[k];
  const char *env = getenv("B200_NUFFT_SLAB_P2P");
  Pub mine{};
  mine.ok = (env && atoi(env) == 0) ? 0u : 1u;
  for (int k = 0; k < 3; ++k) {
    mine.have[k] = bufs[k] != nullptr;
    if (bufs[k] && mine.ok && cudaIpcGetMemHandle(&mine.h[k], bufs[k]) != cudaSuccess) {
      cudaGetLastError();
      mine.ok = 0;
    }
  }
  Scratch<Pub> dsend(1, st_, opts.device), dall(world, st_, opts.device);
  std::vector<Pub> all(world);
  CU(cudaMemcpyAsync(dsend.p, &mine, sizeof(Pub), cudaMemcpyHostToDevice, st_));
  NC(api.AllGather(dsend.p, dall.p, sizeof(Pub), ncclChar, comm, st_));
  CU(cudaMemcpyAsync(all.data(), dall.p, sizeof(Pub) * world, cudaMemcpyDeviceToHost, st_));
  CU(cudaStreamSynchronize(st_));
  uint32_t ok = 1;
  for (int r = 0; r < world; ++r) ok &= all[r].ok;
  if (ok) {
    C **tab[3] = {peer_win_, peer_pencil_, peer_clocal_};
    for (int r = 0; r < world && ok; ++r)
      for (int k = 0; k < 3 && ok; ++k) {
        if (r == rank) {
          tab[k][r] = static_cast<C *>(bufs[k]);
          continue;
        }
        if (!all[r].have[k]) continue;
        void *q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
          break;
        }
        opened_.push_back(q);
        tab[k][r] = static_cast<C *>(q);
      }
  }
  // every rank must take the same path: agree on the outcome of the mapping
  Scratch<uint32_t> flag(1, st_, opts.device);
  CU(cudaMemcpyAsync(flag.p, &ok, sizeof(uint32_t), cudaMemcpyHostToDevice, st_));
  NC(api.AllReduce(flag.p, flag.p, 1, ncclUint32, ncclMin, comm, st_));
  CU(cudaMemcpyAsync(&ok, flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st_));
  CU(cudaStreamSynchronize(st_));
  if (!ok) {
    close_peers();
    return;
  }
  p2p_ = true;
}

// all ranks' streams meet here: work queued before it on any rank is complete on every rank's
// side before work queued after it starts (one-word all-reduce on the plan's stream)
template<class T> void SlabPlan<T>::barrier() {
  NC(nccl().AllReduce(bar_.p, bar_.p, 1, ncclUint32, ncclSum, (ncclComm_t)comm_, st_));
}

// ------------------------------------------------------------------ exchanges
// slab mode.  add = true (type 1): the ghost planes this rank spread into go to the ring
// neighbours that own them and are added there; add = false (type 2): the neighbours' edge
// planes are copied into this rank's ghost planes.  One grouped send/recv for both faces.
template<class T> void SlabPlan<T>::exchange_ghosts(bool add) {
  NcclApi &api    = nccl();
  ncclComm_t comm = (ncclComm_t)comm_;
  const int nxt = (rank + 1) % world, prv = (rank + world - 1) % world;
  const size_t plane = (size_t)nf[0] * nf[1];
  C *w = win_.p;
  if (p2p_) {
    // windows of the neighbours: [below ghost | their nz owned planes | above ghost]
    const int nz_prv = zstart_[prv + 1] - zstart_[prv];
    barrier();  // every rank has finished writing what its neighbours are about to read
    if (add) {
      // my first `above` planes += prev's above-ghost planes, my last `below` += next's below-ghost
      const int64_t na = (int64_t)above * plane * 2, nb = (int64_t)below * plane * 2;
      k_add<T><<<blocks_for(na / 4, 256), 256, 0, st_>>>(
          reinterpret_cast<T *>(ownp_),
          reinterpret_cast<const T *>(peer_win_[prv] + (size_t)(below + nz_prv) * plane), na);
      k_add<T><<<blocks_for(nb / 4, 256), 256, 0, st_>>>(
          reinterpret_cast<T *>(ownp_ + (size_t)(nz - below) * plane),
          reinterpret_cast<const T *>(peer_win_[nxt]), nb);
      launches += 2;
    } else {
      // my below ghost = prev's last `below` owned planes, my above ghost = next's first `above`
      CU(cudaMemcpyAsync(w, peer_win_[prv] + (size_t)nz_prv * plane, below * plane * sizeof(C),
                         cudaMemcpyDefault, st_));
      CU(cudaMemcpyAsync(w + (size_t)(below + nz) * plane, peer_win_[nxt] + (size_t)below * plane,
                         above * plane * sizeof(C), cudaMemcpyDefault, st_));
    }
    return;
  }
  NC(api.GroupStart());
  if (add) {
    NC(api.Send(w + (size_t)(below + nz) * plane, above * plane * sizeof(C), ncclChar, nxt, comm, st_));
    NC(api.Send(w, below * plane * sizeof(C), ncclChar, prv, comm, st_));
    NC(api.Recv(gprev_.p, above * plane * sizeof(C), ncclChar, prv, comm, st_));
    NC(api.Recv(gnext_.p, below * plane * sizeof(C), ncclChar, nxt, comm, st_));
  } else {
    NC(api.Send(ownp_ + (size_t)(nz - below) * plane, below * plane * sizeof(C), ncclChar, nxt, comm, st_));
    NC(api.Send(ownp_, above * plane * sizeof(C), ncclChar, prv, comm, st_));
    NC(api.Recv(w, below * plane * sizeof(C), ncclChar, prv, comm, st_));
    NC(api.Recv(w + (size_t)(below + nz) * plane, above * plane * sizeof(C), ncclChar, nxt, comm, st_));
  }
  NC(api.GroupEnd());
  if (add) {
    const int64_t na = (int64_t)above * plane * 2, nb = (int64_t)below * plane * 2;
    k_add<T><<<blocks_for(na / 4, 256), 256, 0, st_>>>(reinterpret_cast<T *>(ownp_),
                                                       reinterpret_cast<const T *>(gprev_.p), na);
    k_add<T><<<blocks_for(nb / 4, 256), 256, 0, st_>>>(
        reinterpret_cast<T *>(ownp_ + (size_t)(nz - below) * plane),
        reinterpret_cast<const T *>(gnext_.p), nb);
    launches += 2;
  }
}

// replicated-window mode.  reduce = true (type 1): the private copies of the window are summed
// onto the ranks that own the planes; reduce = false (type 2): owners broadcast their planes.
template<class T> void SlabPlan<T>::window_collective(bool reduce) {
  NcclApi &api    = nccl();
  ncclComm_t comm = (ncclComm_t)comm_;
  const size_t plane = (size_t)nf[0] * nf[1];
  const int nf3      = (int)nf[2];
  if (reduce) CU(cudaMemsetAsync(own_.p, 0, sizeof(C) * (size_t)nz * plane, st_));
  NC(api.GroupStart());
  for (const Seg &s : segs_) {
    C *inwin = win_.p + (size_t)((s.g0 - win_org + nf3) % nf3) * plane;
    C *inown = own_.p + (size_t)(s.g0 - zstart_[s.owner]) * plane;  // valid on the owner only
    const size_t count = (size_t)s.n * plane * 2;
    if (reduce)
      NC(api.Reduce(inwin, s.owner == rank ? (void *)inown : (void *)inwin, count, nccl_real<T>(),
                    ncclSum, s.owner, comm, st_));
    else
      NC(api.Broadcast(s.owner == rank ? (const void *)inown : (const void *)inwin, inwin, count,
                       nccl_real<T>(), s.owner, comm, st_));
  }
  NC(api.GroupEnd());
}

// slab <-> pencil: all-to-all of the wanted modes as grouped pairwise send/recv
template<class T> void SlabPlan<T>::transpose(bool to_pencil) {
  const int64_t my = yhi - ylo, m1 = ms[0];
  auto slab_part = [&](int r) {  // my planes, rank r's y range
    return send_.p + (int64_t)nz * m1 * ystart_[r];
  };
  auto slab_count = [&](int r) { return (size_t)nz * (ystart_[r + 1] - ystart_[r]) * m1; };
  auto pen_part  = [&](int r) { return pencil_.p + (int64_t)zstart_[r] * my * m1; };
  auto pen_count = [&](int r) { return (size_t)(zstart_[r + 1] - zstart_[r]) * my * m1; };
  if (world > 1) {
    NcclApi &api    = nccl();
    ncclComm_t comm = (ncclComm_t)comm_;
    NC(api.GroupStart());
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      C *sp = to_pencil ? slab_part(r) : pen_part(r);
      C *rp = to_pencil ? pen_part(r) : slab_part(r);
      const size_t sc = to_pencil ? slab_count(r) : pen_count(r);
      const size_t rc = to_pencil ? pen_count(r) : slab_count(r);
      if (sc) NC(api.Send(sp, sc * sizeof(C), ncclChar, r, comm, st_));
      if (rc) NC(api.Recv(rp, rc * sizeof(C), ncclChar, r, comm, st_));
    }
    NC(api.GroupEnd());
  }
  const size_t self = slab_count(rank);
  if (self)
    CU(cudaMemcpyAsync(to_pencil ? pen_part(rank) : slab_part(rank),
                       to_pencil ? slab_part(rank) : pen_part(rank), self * sizeof(C),
                       cudaMemcpyDeviceToDevice, st_));
}

// strengths to the ranks that own the points (type 1) / values back to the callers (type 2)
template<class T> void SlabPlan<T>::route_values(bool to_owner, C *user) {
  NcclApi &api    = nccl();
  ncclComm_t comm = (ncclComm_t)comm_;
  const uint32_t m = (uint32_t)M;
  if (p2p_) {
    RouteTable t{};
    t.world = world;
    for (int d = 0; d <= world; ++d) t.sendoff[d] = (uint32_t)sendoff_[d];
    for (int d = 0; d < world; ++d) t.peer_off[d] = (uint32_t)peer_off_[d];
    PeerPtrs<C> own{};
    for (int r = 0; r < world; ++r) own.p[r] = peer_clocal_[r];
    if (to_owner) {
      if (m) k_route_peer<C, true><<<blocks_for(m, 256), 256, 0, st_>>>(user, order_.p, own, t, m);
      barrier();  // the owners' arrays are complete
    } else {
      barrier();  // every owner has interpolated
      if (m) k_route_peer<C, false><<<blocks_for(m, 256), 256, 0, st_>>>(user, order_.p, own, t, m);
    }
    if (m) ++launches;
    return;
  }
  if (to_owner && m) {
    k_gather_by<C><<<blocks_for(m, 256), 256, 0, st_>>>(user, order_.p, croute_.p, m);
    ++launches;
  }
  NC(api.GroupStart());
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    const uint64_t sc = to_owner ? sendcnt_[r] : recvcnt_[r];
    const uint64_t rc = to_owner ? recvcnt_[r] : sendcnt_[r];
    C *sp = to_owner ? croute_.p + sendoff_[r] : clocal_.p + recvoff_[r];
    C *rp = to_owner ? clocal_.p + recvoff_[r] : croute_.p + sendoff_[r];
    if (sc) NC(api.Send(sp, sc * sizeof(C), ncclChar, r, comm, st_));
    if (rc) NC(api.Recv(rp, rc * sizeof(C), ncclChar, r, comm, st_));
  }
  NC(api.GroupEnd());
  if (sendcnt_[rank])
    CU(cudaMemcpyAsync(to_owner ? clocal_.p + recvoff_[rank] : croute_.p + sendoff_[rank],
                       to_owner ? croute_.p + sendoff_[rank] : clocal_.p + recvoff_[rank],
                       sendcnt_[rank] * sizeof(C), cudaMemcpyDeviceToDevice, st_));
  if (!to_owner && m) {
    k_scatter_by<C><<<blocks_for(m, 256), 256, 0, st_>>>(croute_.p, order_.p, user, m);
    ++launches;
  }
}

// ------------------------------------------------------------------ execute
template<class T> void SlabPlan<T>::execute(C *c, C *fk_block) {
  DeviceGuard guard(opts.device);
  NvtxRange range(type == 1 ? "b200::slab execute (type 1)" : "b200::slab execute (type 2)");
  if (!eng_) throw Failure{ERR_PLAN_NOTVALID};
  SlabGeom g{};
  for (int d = 0; d < 3; ++d) g.ms[d] = (int)ms[d], g.nf[d] = (int)nf[d];
  g.modeord = modeord;
  g.nz      = nz;
  g.my      = (int)(yhi - ylo);
  g.ylo     = (int)ylo;
  g.world   = world;
  for (int r = 0; r <= world; ++r) g.ystart[r] = ystart_[r];
  SlabPh<T> ph{{ph_[0].p, ph_[1].p, ph_[2].p}};
  const bool routing = mode == 0 && !routed_;
  auto rows = [](int64_t n) { return (unsigned)std::max<int64_t>(1, (n + kRows - 1) / kRows); };
  C *cl = routing ? clocal_.p : c;

  PeerPtrs<C> pen{};
  for (int r = 0; r < world; ++r) pen.p[r] = peer_pencil_[r];
  // replicated-window mode has no ghost barrier: its reduce / broadcast do not order every pair
  // of ranks, so the peers' reads of the previous execute are fenced explicitly
  if (p2p_ && mode == 1) barrier();

  if (type == 1) {
    mark(0);
    if (routing) route_values(true, c);
    mark(1);
    eng_->execute(cl, win_.p, false);  // zero the window, spread
    mark(2);
    if (world > 1) {
      if (mode == 0) exchange_ghosts(true);
      else window_collective(true);
    }
    mark(3);
    fft_exec(fft2_, ownp_, sign);
    mark(4);
    if (p2p_) {  // crop + transpose in one kernel: rows stored into the owners' pencils over NVLink
      if (mode == 1) barrier();  // the owners have finished with their pencils (see above)
      k_slab_pack_peer<C><<<rows((int64_t)nz * ms[1]), kThreads, 0, st_>>>(ownp_, pen, g, z0);
      ++launches;
      mark(5);
      barrier();  // every pencil is complete
    } else {
      k_slab_pack<C><<<rows((int64_t)nz * ms[1]), kThreads, 0, st_>>>(ownp_, send_.p, g);
      ++launches;
      mark(5);
      transpose(true);
    }
    mark(6);
    if (have1_) fft_exec(fft1_, pencil_.p, sign);
    mark(7);
    if (g.my) {
      k_slab_deconv<C, T><<<rows(ms[2] * g.my), kThreads, 0, st_>>>(pencil_.p, fk_block, g, ph);
      ++launches;
    }
    mark(8);
  } else {
    mark(0);
    if (g.my) {
      k_slab_amplify<C, T><<<rows(nf[2] * g.my), kThreads, 0, st_>>>(fk_block, pencil_.p, g, ph);
      ++launches;
    }
    mark(1);
    if (have1_) fft_exec(fft1_, pencil_.p, sign);
    mark(2);
    if (p2p_) {  // transpose + zero-pad in one kernel: rows loaded from the owners' pencils
      barrier();   // every pencil is ready (and nobody still reads this rank's planes)
      mark(3);
      k_slab_unpack_peer<C, T><<<rows((int64_t)nz * nf[1]), kThreads, 0, st_>>>(pen, ownp_, g, z0);
    } else {
      transpose(false);
      mark(3);
      k_slab_unpack<C, T><<<rows((int64_t)nz * nf[1]), kThreads, 0, st_>>>(send_.p, ownp_, g);
    }
    ++launches;
    mark(4);
    fft_exec(fft2_, ownp_, sign);
    mark(5);
    if (world > 1) {
      if (mode == 0) exchange_ghosts(false);
      else window_collective(false);
    } else if (mode == 1) {
    }
    mark(6);
    eng_->execute(cl, win_.p, false);  // interpolate from the window
    mark(7);
    if (routing) route_values(false, c);
    mark(8);
  }
  CU(cudaGetLastError());
}

template<class T> void SlabPlan<T>::gather_modes(const C *fk_block, C *fk_full) {
  DeviceGuard guard(opts.device);
  SlabGeom g{};
  for (int d = 0; d < 3; ++d) g.ms[d] = (int)ms[d], g.nf[d] = (int)nf[d];
  g.world = world;
  for (int r = 0; r <= world; ++r) g.ystart[r] = ystart_[r];
  const int64_t slice = ms[2] * ms[0];
  gath_.alloc((size_t)(slice * ms[1]));
  if (world > 1) {
    NcclApi &api    = nccl();
    ncclComm_t comm = (ncclComm_t)comm_;
    const size_t mine = (size_t)(slice * (yhi - ylo));
    NC(api.GroupStart());
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      const size_t theirs = (size_t)(slice * (ystart_[r + 1] - ystart_[r]));
      if (mine) NC(api.Send(fk_block, mine * sizeof(C), ncclChar, r, comm, st_));
      if (theirs)
        NC(api.Recv(gath_.p + slice * ystart_[r], theirs * sizeof(C), ncclChar, r, comm, st_));
    }
    NC(api.GroupEnd());
  }
  if (yhi > ylo)
    CU(cudaMemcpyAsync(gath_.p + slice * ylo, fk_block, sizeof(C) * (size_t)(slice * (yhi - ylo)),
                       cudaMemcpyDeviceToDevice, st_));
  k_slab_interleave<C><<<(unsigned)((ms[2] * ms[1] + kRows - 1) / kRows), kThreads, 0, st_>>>(
      gath_.p, fk_full, g, 1);
  ++launches;
  CU(cudaGetLastError());
}

template<class T> void SlabPlan<T>::slice_modes(const C *fk_full, C *fk_block) {
  DeviceGuard guard(opts.device);
  SlabGeom g{};
  for (int d = 0; d < 3; ++d) g.ms[d] = (int)ms[d], g.nf[d] = (int)nf[d];
  g.world = world;
  for (int r = 0; r <= world; ++r) g.ystart[r] = ystart_[r];
  const int64_t slice = ms[2] * ms[0];
  gath_.alloc((size_t)(slice * ms[1]));
  k_slab_interleave<C><<<(unsigned)((ms[2] * ms[1] + kRows - 1) / kRows), kThreads, 0, st_>>>(
      gath_.p, const_cast<C *>(fk_full), g, 0);
  ++launches;
  if (yhi > ylo)
    CU(cudaMemcpyAsync(fk_block, gath_.p + slice * ylo, sizeof(C) * (size_t)(slice * (yhi - ylo)),
                       cudaMemcpyDeviceToDevice, st_));
  CU(cudaGetLastError());
}

template class SlabPlan<float>;
template class SlabPlan<double>;

}  // namespace b200
