// extern "C" boundary of finufft_b200: the cufinufft_* (device pointer) and finufft_* (host
// pointer) guru + simple entry points declared in include/b200_cufinufft.h and
// include/b200_finufft.h, plus the introspection calls of include/b200_introspect.h.
// Exceptions thrown by the engine become the reference's integer codes here, the convention
// of reference include/finufft_common/safe_call.h:57-80.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "../../include/b200_cufinufft.h"
#include "../../include/b200_finufft.h"
#include "../../include/b200_introspect.h"
#include "engine.hpp"
#include "planmath.hpp"

using namespace b200;

namespace {

constexpr uint32_t kMagic = 0xB2005EEDu;

struct PlanBase {
  uint32_t magic = kMagic;
  bool is_float  = false;
  bool host_api  = false;
  virtual ~PlanBase() { magic = 0; }
};

template<class T> struct DevicePlan : PlanBase {
  Engine<T> eng;
  DevicePlan(int type, int dim, const int64_t *nm, int iflag, int ntr, double tol,
             const EngineOpts &o)
      : eng(type, dim, nm, iflag, ntr, tol, o) {
    is_float = std::is_same<T, float>::value;
  }
};

// Host-pointer plan: keeps device mirrors of the user's arrays.
template<class T> struct HostPlan : DevicePlan<T> {
  using C = typename CxOf<T>::type;
  DevBuf<T> x, y, z, s, t, u;
  DevBuf<C> c, fk;
  int64_t M = 0, N = 0;
  // pipelined execute: copy streams for the two directions and events per unit of user data
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev;
  HostPlan(int type, int dim, const int64_t *nm, int iflag, int ntr, double tol,
           const EngineOpts &o)
      : DevicePlan<T>(type, dim, nm, iflag, ntr, tol, o) {
    this->host_api = true;
  }
  ~HostPlan() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
  }
  cudaEvent_t event(size_t i) {
    while (ev.size() <= i) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
        throw Failure{ERR_CUDA_FAILURE};
      ev.push_back(e);
    }
    return ev[i];
  }
};

// Point groups of a host-pointer plan (engine.hpp set_point_groups): with one transform vector
// of at least 64 MB, 4 groups let the spread of group k overlap the upload of group k+1 (and
// the download of group k overlap the interpolation of k+1).  B200_NUFFT_HOST_GROUPS overrides.
template<class T> int host_point_groups(int64_t M, int ntr, int type, int64_t grid_cells) {
  if (const char *env = getenv("B200_NUFFT_HOST_GROUPS")) return std::max(1, atoi(env));
  if (type == 3 || ntr > 1) return 1;  // many vectors pipeline vector by vector instead
  // every group sweeps the whole grid, so splitting pays only for dense point sets (measured at
  // 0.75 and 6 points per fine-grid cell); sparse ones keep one group
  if (2 * M < grid_cells) return 1;
  return (uint64_t)M * sizeof(typename CxOf<T>::type) >= (64ull << 20) ? 4 : 1;
}

template<class F> int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const Failure &e) {
    return e.code;
  } catch (const std::bad_alloc &) {
    return ERR_ALLOC;
  } catch (...) {
    return ERR_UNKNOWN_EXCEPTION;
  }
}

void check_cuda(cudaError_t e) {
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Failure{e == cudaErrorMemoryAllocation ? ERR_ALLOC : ERR_CUDA_FAILURE};
  }
}

// reference src/cuda/c_interface.cpp:14-31: device API indexes with 32-bit ints
void validate_modes_gpu(int type, int dim, const int64_t *nm) {
  if (dim < 1 || dim > 3) throw Failure{ERR_DIM_NOTVALID};
  if (type == 3) return;
  if (!nm) throw Failure{ERR_INVALID_ARGUMENT};
  int64_t tot = 1;
  for (int d = 0; d < dim; ++d) {
    if (nm[d] <= 0 || nm[d] > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
    tot *= nm[d];
    if (tot > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
  }
}

EngineOpts from_gpu_opts(const cufinufft_opts *o) {
  cufinufft_opts d;
  if (o) d = *o;
  else cufinufft_default_opts(&d);
  EngineOpts e;
  e.upsampfac        = d.upsampfac;
  e.spreadinterponly = d.gpu_spreadinterponly;
  e.maxbatch         = d.gpu_maxbatchsize;
  e.device           = d.gpu_device_id;
  e.stream           = (cudaStream_t)d.gpu_stream;
  e.modeord          = d.modeord;
  e.maxsub           = d.gpu_maxsubprobsize > 0 ? d.gpu_maxsubprobsize : 1024;
  e.debug            = d.debug;
  e.allow_eps_too_small = 1;  // the device API has no such switch; clamp and proceed
  e.check_sigma         = 0;
  e.sort                = d.gpu_sort ? 1 : 0;  // include/cufinufft_opts.h:11
  return e;
}

EngineOpts from_host_opts(const finufft_opts *o) {
  finufft_opts d;
  if (o) d = *o;
  else finufft_default_opts(&d);
  if (d.spread_kerformula != 0) throw Failure{ERR_KERFORMULA_NOTVALID};
  EngineOpts e;
  e.upsampfac        = d.upsampfac;
  e.spreadinterponly = d.spreadinterponly;
  e.maxbatch         = d.maxbatchsize;
  int dev            = 0;
  check_cuda(cudaGetDevice(&dev));
  e.device  = dev;
  e.stream  = nullptr;
  e.modeord = d.modeord;
  e.maxsub  = d.spread_max_sp_size > 0 ? d.spread_max_sp_size : 1024;
  e.debug   = d.debug;
  e.allow_eps_too_small = d.allow_eps_too_small;
  e.check_sigma         = 1;
  e.auto_sigma          = d.upsampfac == 0.0;  // include/finufft_opts.h:44: 0 = auto
  // include/finufft_opts.h:41: 0 don't sort, 1 sort, 2 heuristic choice.  The reference's
  // heuristic (spreadinterp.hpp:161-163) skips the sort for 1D type 2 / tiny grids because a
  // CPU thread streams such points well; on this device sorted points always win, so the
  // library's own choice for 2 is to sort.
  e.sort = d.spread_sort == 0 ? 0 : 1;
  return e;
}

// Option contract of the reference device API that callers (and the reference's own
// test/cuda/test_makeplan.c, cufinufft_error_handling.cu) rely on, although this engine has a
// single native method and its own bins:
//  * a nonstandard upsampfac (not 0, 2 or 1.25) with gpu_kerevalmeth = 1 is error 8, upsampfac
//    <= 1 otherwise error 7 (src/cuda/makeplan.cu:60-72);
//  * a user-chosen shared-memory method (gpu_method 2 or 3) whose user-chosen bins would not fit
//    the device's shared memory, or an unknown gpu_method, is error 19
//    (src/cuda/makeplan.cu:307-327, src/cuda/heuristics.cu:34-42,326-445).
template<class T> void validate_opts_gpu(int dim, double tol, const cufinufft_opts *o) {
  if (!o) return;
  const double sigma = o->upsampfac;
  if (sigma != 0.0 && sigma != 2.0 && sigma != 1.25) {
    if (o->gpu_kerevalmeth == 1) throw Failure{ERR_HORNER_WRONG_BETA};
    if (sigma <= 1.0) throw Failure{ERR_UPSAMPFAC_TOO_SMALL};
  }
  if (o->gpu_method < 0 || o->gpu_method > 4) throw Failure{ERR_INSUFFICIENT_SHMEM};
  const bool user_bins = (o->gpu_binsizex | o->gpu_binsizey | o->gpu_binsizez) != 0;
  if ((o->gpu_method == 2 || o->gpu_method == 3) && user_bins) {
    int limit = 0;
    if (cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin,
                               o->gpu_device_id) != cudaSuccess) {
      cudaGetLastError();
      throw Failure{ERR_CUDA_FAILURE};
    }
    // width the reference device library would pick (src/cuda/makeplan.cu:88-93)
    const double eps = std::max(tol, (double)std::numeric_limits<T>::epsilon());
    int ns = (int)std::ceil(-std::log10(eps / 10.0));
    if (sigma != 0.0 && sigma != 2.0)
      ns = (int)std::ceil(-std::log(eps) / (3.14159265358979323846 * std::sqrt(1 - 1 / sigma)));
    ns = std::max(2, ns);
    const int64_t pad = 2 * ((ns + 1) / 2);
    const int64_t bx = o->gpu_binsizex ? o->gpu_binsizex : 1, by = o->gpu_binsizey ? o->gpu_binsizey : 1,
                  bz = o->gpu_binsizez ? o->gpu_binsizez : 1;
    if (bx < 0 || by < 0 || bz < 0) throw Failure{ERR_BINSIZE_NOTVALID};
    double cells = (double)(bx + pad);
    if (dim > 1) cells *= (double)(by + pad);
    if (dim > 2) cells *= (double)(bz + pad);
    if (cells * 2.0 * sizeof(T) > (double)limit) throw Failure{ERR_INSUFFICIENT_SHMEM};
  }
}

template<class T>
int gpu_makeplan(int type, int dim, const int64_t *nm, int iflag, int ntr, double tol, void **out,
                 const cufinufft_opts *o) {
  return guarded([&] {
    if (!out) throw Failure{ERR_INVALID_ARGUMENT};
    *out = nullptr;
    validate_modes_gpu(type, dim, nm);
    if (type < 1 || type > 3) throw Failure{ERR_TYPE_NOTVALID};
    if (ntr < 1) throw Failure{ERR_NTRANS_NOTVALID};
    validate_opts_gpu<T>(dim, tol, o);
    *out = new DevicePlan<T>(type, dim, nm, iflag, ntr, tol, from_gpu_opts(o));
  });
}

template<class T> DevicePlan<T> *as_plan(void *p) {
  auto *b = static_cast<PlanBase *>(p);
  if (!b || b->magic != kMagic || b->is_float != std::is_same<T, float>::value)
    throw Failure{ERR_PLAN_NOTVALID};
  return static_cast<DevicePlan<T> *>(b);
}

template<class T>
int gpu_setpts(void *plan, int64_t M, const T *x, const T *y, const T *z, int64_t N, const T *s,
               const T *t, const T *u) {
  return guarded([&] {
    auto *p = as_plan<T>(plan);
    if (M > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NDATA_NOTVALID};
    p->eng.setpts(M, x, y, z, N, s, t, u);
  });
}
template<class T> int gpu_execute(void *plan, void *c, void *fk, bool adjoint = false) {
  using C = typename CxOf<T>::type;
  return guarded([&] { as_plan<T>(plan)->eng.execute((C *)c, (C *)fk, adjoint); });
}
template<class T> int gpu_destroy(void *plan) {
  return guarded([&] {
    if (!plan) throw Failure{ERR_PLAN_NOTVALID};
    delete as_plan<T>(plan);
  });
}

// one-shot: plan + setpts + execute + destroy (reference src/cuda/c_interface.cpp:188-218)
template<class T>
int gpu_simple(int dim, int type, int ntr, int64_t M, const T *x, const T *y, const T *z,
               void *c, int iflag, T eps, int64_t m0, int64_t m1, int64_t m2, int64_t nk,
               const T *s, const T *t, const T *u, void *fk, const cufinufft_opts *o) {
  const int64_t nm[3] = {m0, m1, m2};
  void *plan          = nullptr;
  int err             = gpu_makeplan<T>(type, dim, nm, iflag, ntr, (double)eps, &plan, o);
  if (err) return err;
  err = gpu_setpts<T>(plan, M, x, y, z, nk, s, t, u);
  if (!err) err = gpu_execute<T>(plan, c, fk);
  if (!err) {
    auto *p = static_cast<DevicePlan<T> *>(static_cast<PlanBase *>(plan));
    if (cudaStreamSynchronize(p->eng.stream()) != cudaSuccess) err = ERR_CUDA_FAILURE;
  }
  gpu_destroy<T>(plan);
  return err;
}

// ---------------------------------------------------------------- host-pointer plans
template<class T>
int host_makeplan(int type, int dim, const int64_t *nm, int iflag, int ntr, double tol,
                  void **out, const finufft_opts *o) {
  return guarded([&] {
    if (!out) throw Failure{ERR_INVALID_ARGUMENT};
    *out = nullptr;
    if (type < 1 || type > 3) throw Failure{ERR_TYPE_NOTVALID};
    if (dim < 1 || dim > 3) throw Failure{ERR_DIM_NOTVALID};
    if (ntr < 1) throw Failure{ERR_NTRANS_NOTVALID};
    if (type != 3) validate_modes_gpu(type, dim, nm);
    *out = new HostPlan<T>(type, dim, nm, iflag, ntr, tol, from_host_opts(o));
  });
}
template<class T> HostPlan<T> *as_host_plan(void *p) {
  auto *q = as_plan<T>(p);
  if (!q->host_api) throw Failure{ERR_PLAN_NOTVALID};
  return static_cast<HostPlan<T> *>(q);
}
template<class T> void upload(DevBuf<T> &d, const T *h, int64_t n, cudaStream_t st) {
  d.alloc((size_t)n);
  if (n) check_cuda(cudaMemcpyAsync(d.p, h, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, st));
}
template<class T>
int host_setpts(void *plan, int64_t M, const T *x, const T *y, const T *z, int64_t N, const T *s,
                const T *t, const T *u) {
  return guarded([&] {
    auto *p = as_host_plan<T>(plan);
    if (M < 0) throw Failure{ERR_NUM_NU_PTS_INVALID};
    if (M > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NUM_NU_PTS_INVALID};
    DeviceGuard guard(p->eng.opts.device);
    cudaStream_t st = p->eng.stream();
    const int dim   = p->eng.dim;
    upload<T>(p->x, x, M, st);
    if (dim > 1) upload<T>(p->y, y, M, st);
    if (dim > 2) upload<T>(p->z, z, M, st);
    if (p->eng.type == 3) {
      if (N < 0 || N > std::numeric_limits<int32_t>::max()) throw Failure{ERR_NUM_NU_PTS_INVALID};
      upload<T>(p->s, s, N, st);
      if (dim > 1) upload<T>(p->t, t, N, st);
      if (dim > 2) upload<T>(p->u, u, N, st);
    }
    p->M = M;
    p->N = N;
    p->eng.set_point_groups(host_point_groups<T>(M, p->eng.ntr, p->eng.type, p->eng.grid_cells()));
    p->eng.setpts(M, p->x.p, p->y.p, p->z.p, N, p->s.p, p->t.p, p->u.p);
  });
}
// Host-pointer execute.  Types 1 and 2 are pipelined: all uploads are queued on one copy stream
// in the order the engine consumes them, all downloads on another, and the engine's hooks
// (ExecHooks) tie them to the compute stream with events, so that PCIe transfers and kernels
// overlap: vector by vector for ntransf > 1, point group by point group for one large vector.
template<class T> int host_execute(void *plan, void *c, void *fk, bool adjoint) {
  using C = typename CxOf<T>::type;
  return guarded([&] {
    auto *p = as_host_plan<T>(plan);
    DeviceGuard guard(p->eng.opts.device);
    cudaStream_t st     = p->eng.stream();
    const int ntr       = p->eng.ntr;
    const int64_t M     = p->M;
    const int64_t nout  = p->eng.type == 3 ? p->N : p->eng.mode_count();
    const size_t nc_tot = (size_t)M * ntr, nk_tot = (size_t)nout * ntr;
    p->c.alloc(nc_tot);
    p->fk.alloc(nk_tot);
    const bool c_is_input = (p->eng.type != 2) != adjoint;
    C *hc = static_cast<C *>(c), *hfk = static_cast<C *>(fk);
    if (p->eng.type == 3 || nc_tot == 0 || nk_tot == 0) {  // plain: upload, run, download
      if (c_is_input) {
        if (nc_tot) check_cuda(cudaMemcpyAsync(p->c.p, c, sizeof(C) * nc_tot, cudaMemcpyHostToDevice, st));
      } else {
        if (nk_tot) check_cuda(cudaMemcpyAsync(p->fk.p, fk, sizeof(C) * nk_tot, cudaMemcpyHostToDevice, st));
      }
      p->eng.execute(p->c.p, p->fk.p, adjoint);
      if (c_is_input) {
        if (nk_tot) check_cuda(cudaMemcpyAsync(fk, p->fk.p, sizeof(C) * nk_tot, cudaMemcpyDeviceToHost, st));
      } else {
        if (nc_tot) check_cuda(cudaMemcpyAsync(c, p->c.p, sizeof(C) * nc_tot, cudaMemcpyDeviceToHost, st));
      }
      check_cuda(cudaStreamSynchronize(st));
      return;
    }
    if (!p->s_in) check_cuda(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    if (!p->s_out) check_cuda(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    // Whatever happens below (a CUDA error, an engine failure), no copy that reads or writes
    // the caller's host buffers may still be in flight when this call returns.
    struct Drain {
      cudaStream_t a, b, c;
      ~Drain() {
        cudaStreamSynchronize(a);
        cudaStreamSynchronize(b);
        cudaStreamSynchronize(c);
      }
    } drain{st, p->s_in, p->s_out};
    const int ngrp      = p->eng.point_groups();
    const int batch     = p->eng.batch;
    const int nbatches  = (ntr + batch - 1) / batch;
    // event slots: [0] start, then one per point unit (v, k), then one per mode batch
    auto ev_pts  = [&](int v, int k) { return p->event(1 + (size_t)v * ngrp + k); };
    auto ev_mode = [&](int b) { return p->event(1 + (size_t)ntr * ngrp + b); };
    auto pt_range = [&](int v, int k, int64_t &off, int64_t &cnt) {
      const int64_t a = ngrp > 1 ? p->eng.group_begin(k) : 0;
      const int64_t e = ngrp > 1 ? p->eng.group_begin(k + 1) : M;
      off = (int64_t)v * M + a;
      cnt = e - a;
    };
    // copies may not start before earlier work on the plan's stream (previous execute) is done
    check_cuda(cudaEventRecord(p->event(0), st));
    check_cuda(cudaStreamWaitEvent(p->s_in, p->event(0), 0));
    ExecHooks hooks;
    if (c_is_input) {
      for (int v = 0; v < ntr; ++v)
        for (int k = 0; k < ngrp; ++k) {
          int64_t off, cnt;
          pt_range(v, k, off, cnt);
          if (cnt > 0)
            check_cuda(cudaMemcpyAsync(p->c.p + off, hc + off, sizeof(C) * (size_t)cnt,
                                       cudaMemcpyHostToDevice, p->s_in));
          check_cuda(cudaEventRecord(ev_pts(v, k), p->s_in));
        }
      hooks.before_points = [&](int v, int k) {
        check_cuda(cudaStreamWaitEvent(st, ev_pts(v, k), 0));
      };
      hooks.after_modes = [&](int b0, int nb) {
        const int b = b0 / batch;
        check_cuda(cudaEventRecord(ev_mode(b), st));
        check_cuda(cudaStreamWaitEvent(p->s_out, ev_mode(b), 0));
        check_cuda(cudaMemcpyAsync(hfk + (size_t)b0 * nout, p->fk.p + (size_t)b0 * nout,
                                   sizeof(C) * (size_t)nb * nout, cudaMemcpyDeviceToHost, p->s_out));
      };
    } else {
      for (int b = 0; b < nbatches; ++b) {
        const int b0 = b * batch, nb = std::min(batch, ntr - b0);
        check_cuda(cudaMemcpyAsync(p->fk.p + (size_t)b0 * nout, hfk + (size_t)b0 * nout,
                                   sizeof(C) * (size_t)nb * nout, cudaMemcpyHostToDevice, p->s_in));
        check_cuda(cudaEventRecord(ev_mode(b), p->s_in));
      }
      hooks.before_modes = [&](int b0, int) {
        check_cuda(cudaStreamWaitEvent(st, ev_mode(b0 / batch), 0));
      };
      hooks.after_points = [&](int v, int k) {
        int64_t off, cnt;
        pt_range(v, k, off, cnt);
        check_cuda(cudaEventRecord(ev_pts(v, k), st));
        check_cuda(cudaStreamWaitEvent(p->s_out, ev_pts(v, k), 0));
        if (cnt > 0)
          check_cuda(cudaMemcpyAsync(hc + off, p->c.p + off, sizeof(C) * (size_t)cnt,
                                     cudaMemcpyDeviceToHost, p->s_out));
      };
    }
    p->eng.execute(p->c.p, p->fk.p, adjoint, &hooks);
    check_cuda(cudaStreamSynchronize(st));
    check_cuda(cudaStreamSynchronize(p->s_in));
    check_cuda(cudaStreamSynchronize(p->s_out));
  });
}
template<class T> int host_destroy(void *plan) {
  if (!plan) return 1;  // reference src/c_interface.cpp:94-95
  return guarded([&] { delete as_host_plan<T>(plan); });
}
template<class T>
int host_simple(int dim, int type, int ntr, int64_t M, const T *x, const T *y, const T *z,
                void *c, int iflag, T eps, int64_t m0, int64_t m1, int64_t m2, int64_t nk,
                const T *s, const T *t, const T *u, void *fk, const finufft_opts *o) {
  const int64_t nm[3] = {m0, m1, m2};
  void *plan          = nullptr;
  int err             = host_makeplan<T>(type, dim, nm, iflag, ntr, (double)eps, &plan, o);
  if (err) return err;
  err = host_setpts<T>(plan, M, x, y, z, nk, s, t, u);
  if (!err) err = host_execute<T>(plan, c, fk, false);
  host_destroy<T>(plan);
  return err;
}

template<class T> void fill_info(const Engine<T> &e, b200_plan_info *out) {
  out->is_float = std::is_same<T, float>::value;
  out->type  = e.type;
  out->dim   = e.dim;
  out->ntr   = e.ntr;
  out->ns    = e.ns;
  out->nc    = e.nc;
  out->batch = e.batch;
  out->sigma = e.sigma;
  out->beta  = e.beta;
  out->tol   = e.tol;
  for (int d = 0; d < 3; ++d) {
    out->nf[d]    = e.nf[d];
    out->ms[d]    = e.ms[d];
    out->nbins[d] = e.geom.nb[d];
  }
  out->M    = e.M;
  out->nsub = e.nsub;
}
}  // namespace

// =====================================================================================
extern "C" {

void cufinufft_default_opts(cufinufft_opts *o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->upsampfac          = 0.0;   // choose (2.0)
  o->gpu_method         = 0;
  o->gpu_sort           = 1;
  o->gpu_maxsubprobsize = 1024;
  o->gpu_kerevalmeth    = 1;
  o->gpu_device_id      = 0;
  o->gpu_stream         = nullptr;
}

int cufinufft_makeplan(int type, int dim, const int64_t *nm, int iflag, int ntr, double eps,
                       cufinufft_plan *plan, const cufinufft_opts *o) {
  return gpu_makeplan<double>(type, dim, nm, iflag, ntr, eps, (void **)plan, o);
}
int cufinufftf_makeplan(int type, int dim, const int64_t *nm, int iflag, int ntr, float eps,
                        cufinufftf_plan *plan, const cufinufft_opts *o) {
  return gpu_makeplan<float>(type, dim, nm, iflag, ntr, (double)eps, (void **)plan, o);
}
int cufinufft_setpts(cufinufft_plan p, int64_t M, const double *x, const double *y,
                     const double *z, int N, const double *s, const double *t, const double *u) {
  return gpu_setpts<double>(p, M, x, y, z, N, s, t, u);
}
int cufinufftf_setpts(cufinufftf_plan p, int64_t M, const float *x, const float *y,
                      const float *z, int N, const float *s, const float *t, const float *u) {
  return gpu_setpts<float>(p, M, x, y, z, N, s, t, u);
}
int cufinufft_execute(cufinufft_plan p, void *c, void *fk) { return gpu_execute<double>(p, c, fk); }
int cufinufftf_execute(cufinufftf_plan p, void *c, void *fk) { return gpu_execute<float>(p, c, fk); }
int cufinufft_destroy(cufinufft_plan p) { return gpu_destroy<double>(p); }
int cufinufftf_destroy(cufinufftf_plan p) { return gpu_destroy<float>(p); }

#define NUL nullptr
#define B200_CU_SIMPLE_DEF(P, R)                                                                \
  int cufinufft##P##1d1many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,     \
                            int64_t ms, void *fk, const cufinufft_opts *o) {                     \
    return gpu_simple<R>(1, 1, ntr, M, x, NUL, NUL, (void *)c, iflag, eps, ms, 1, 1, 0, NUL,     \
                         NUL, NUL, fk, o);                                                       \
  }                                                                                              \
  int cufinufft##P##1d1(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t ms,      \
                        void *fk, const cufinufft_opts *o) {                                     \
    return cufinufft##P##1d1many(1, M, x, c, iflag, eps, ms, fk, o);                             \
  }                                                                                              \
  int cufinufft##P##1d2many(int ntr, int64_t M, const R *x, void *c, int iflag, R eps,           \
                            int64_t ms, const void *fk, const cufinufft_opts *o) {               \
    return gpu_simple<R>(1, 2, ntr, M, x, NUL, NUL, c, iflag, eps, ms, 1, 1, 0, NUL, NUL, NUL,   \
                         (void *)fk, o);                                                         \
  }                                                                                              \
  int cufinufft##P##1d2(int64_t M, const R *x, void *c, int iflag, R eps, int64_t ms,            \
                        const void *fk, const cufinufft_opts *o) {                               \
    return cufinufft##P##1d2many(1, M, x, c, iflag, eps, ms, fk, o);                             \
  }                                                                                              \
  int cufinufft##P##1d3many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,     \
                            int64_t nk, const R *s, void *fk, const cufinufft_opts *o) {         \
    return gpu_simple<R>(1, 3, ntr, M, x, NUL, NUL, (void *)c, iflag, eps, 1, 1, 1, nk, s, NUL,  \
                         NUL, fk, o);                                                            \
  }                                                                                              \
  int cufinufft##P##1d3(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t nk,      \
                        const R *s, void *fk, const cufinufft_opts *o) {                         \
    return cufinufft##P##1d3many(1, M, x, c, iflag, eps, nk, s, fk, o);                          \
  }                                                                                              \
  int cufinufft##P##2d1many(int ntr, int64_t M, const R *x, const R *y, const void *c,           \
                            int iflag, R eps, int64_t ms, int64_t mt, void *fk,                  \
                            const cufinufft_opts *o) {                                           \
    return gpu_simple<R>(2, 1, ntr, M, x, y, NUL, (void *)c, iflag, eps, ms, mt, 1, 0, NUL, NUL, \
                         NUL, fk, o);                                                            \
  }                                                                                              \
  int cufinufft##P##2d1(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,      \
                        int64_t ms, int64_t mt, void *fk, const cufinufft_opts *o) {             \
    return cufinufft##P##2d1many(1, M, x, y, c, iflag, eps, ms, mt, fk, o);                      \
  }                                                                                              \
  int cufinufft##P##2d2many(int ntr, int64_t M, const R *x, const R *y, void *c, int iflag,      \
                            R eps, int64_t ms, int64_t mt, const void *fk,                       \
                            const cufinufft_opts *o) {                                           \
    return gpu_simple<R>(2, 2, ntr, M, x, y, NUL, c, iflag, eps, ms, mt, 1, 0, NUL, NUL, NUL,    \
                         (void *)fk, o);                                                         \
  }                                                                                              \
  int cufinufft##P##2d2(int64_t M, const R *x, const R *y, void *c, int iflag, R eps,            \
                        int64_t ms, int64_t mt, const void *fk, const cufinufft_opts *o) {       \
    return cufinufft##P##2d2many(1, M, x, y, c, iflag, eps, ms, mt, fk, o);                      \
  }                                                                                              \
  int cufinufft##P##2d3many(int ntr, int64_t M, const R *x, const R *y, const void *c,           \
                            int iflag, R eps, int64_t nk, const R *s, const R *t, void *fk,      \
                            const cufinufft_opts *o) {                                           \
    return gpu_simple<R>(2, 3, ntr, M, x, y, NUL, (void *)c, iflag, eps, 1, 1, 1, nk, s, t, NUL, \
                         fk, o);                                                                 \
  }                                                                                              \
  int cufinufft##P##2d3(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,      \
                        int64_t nk, const R *s, const R *t, void *fk,                            \
                        const cufinufft_opts *o) {                                               \
    return cufinufft##P##2d3many(1, M, x, y, c, iflag, eps, nk, s, t, fk, o);                    \
  }                                                                                              \
  int cufinufft##P##3d1many(int ntr, int64_t M, const R *x, const R *y, const R *z,              \
                            const void *c, int iflag, R eps, int64_t ms, int64_t mt,             \
                            int64_t mu, void *fk, const cufinufft_opts *o) {                     \
    return gpu_simple<R>(3, 1, ntr, M, x, y, z, (void *)c, iflag, eps, ms, mt, mu, 0, NUL, NUL,  \
                         NUL, fk, o);                                                            \
  }                                                                                              \
  int cufinufft##P##3d1(int64_t M, const R *x, const R *y, const R *z, const void *c,            \
                        int iflag, R eps, int64_t ms, int64_t mt, int64_t mu, void *fk,          \
                        const cufinufft_opts *o) {                                               \
    return cufinufft##P##3d1many(1, M, x, y, z, c, iflag, eps, ms, mt, mu, fk, o);               \
  }                                                                                              \
  int cufinufft##P##3d2many(int ntr, int64_t M, const R *x, const R *y, const R *z, void *c,     \
                            int iflag, R eps, int64_t ms, int64_t mt, int64_t mu,                \
                            const void *fk, const cufinufft_opts *o) {                           \
    return gpu_simple<R>(3, 2, ntr, M, x, y, z, c, iflag, eps, ms, mt, mu, 0, NUL, NUL, NUL,     \
                         (void *)fk, o);                                                         \
  }                                                                                              \
  int cufinufft##P##3d2(int64_t M, const R *x, const R *y, const R *z, void *c, int iflag,       \
                        R eps, int64_t ms, int64_t mt, int64_t mu, const void *fk,               \
                        const cufinufft_opts *o) {                                               \
    return cufinufft##P##3d2many(1, M, x, y, z, c, iflag, eps, ms, mt, mu, fk, o);               \
  }                                                                                              \
  int cufinufft##P##3d3many(int ntr, int64_t M, const R *x, const R *y, const R *z,              \
                            const void *c, int iflag, R eps, int64_t nk, const R *s,             \
                            const R *t, const R *u, void *fk, const cufinufft_opts *o) {         \
    return gpu_simple<R>(3, 3, ntr, M, x, y, z, (void *)c, iflag, eps, 1, 1, 1, nk, s, t, u, fk, \
                         o);                                                                     \
  }                                                                                              \
  int cufinufft##P##3d3(int64_t M, const R *x, const R *y, const R *z, const void *c,            \
                        int iflag, R eps, int64_t nk, const R *s, const R *t, const R *u,        \
                        void *fk, const cufinufft_opts *o) {                                     \
    return cufinufft##P##3d3many(1, M, x, y, z, c, iflag, eps, nk, s, t, u, fk, o);              \
  }
B200_CU_SIMPLE_DEF(, double)
B200_CU_SIMPLE_DEF(f, float)

// ------------------------------------------------------------------ host-pointer API
static void host_default_opts(finufft_opts *o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->showwarn           = 1;
  o->fftw               = 64;  // FINUFFT_FFT_DEFAULT of an FFTW build; unused here
  o->spread_sort        = 2;
  o->spread_kerevalmeth = 1;
  o->spread_kerpad      = 1;
  o->upsampfac          = 0.0;
  o->spread_nthr_atomic = -1;
}
extern const int FINUFFT_FFT_DEFAULT;
const int FINUFFT_FFT_DEFAULT = 64;  // reference include/finufft_opts.h:27
void finufft_default_opts(finufft_opts *o) { host_default_opts(o); }
void finufftf_default_opts(finufft_opts *o) { host_default_opts(o); }

int finufft_makeplan(int type, int dim, const int64_t *nm, int iflag, int ntr, double tol,
                     finufft_plan *plan, finufft_opts *o) {
  return host_makeplan<double>(type, dim, nm, iflag, ntr, tol, (void **)plan, o);
}
int finufftf_makeplan(int type, int dim, const int64_t *nm, int iflag, int ntr, float tol,
                      finufftf_plan *plan, finufft_opts *o) {
  return host_makeplan<float>(type, dim, nm, iflag, ntr, (double)tol, (void **)plan, o);
}
int finufft_setpts(finufft_plan p, int64_t M, const double *x, const double *y, const double *z,
                   int64_t N, const double *s, const double *t, const double *u) {
  return host_setpts<double>(p, M, x, y, z, N, s, t, u);
}
int finufftf_setpts(finufftf_plan p, int64_t M, const float *x, const float *y, const float *z,
                    int64_t N, const float *s, const float *t, const float *u) {
  return host_setpts<float>(p, M, x, y, z, N, s, t, u);
}
int finufft_execute(finufft_plan p, void *c, void *fk) { return host_execute<double>(p, c, fk, false); }
int finufftf_execute(finufftf_plan p, void *c, void *fk) { return host_execute<float>(p, c, fk, false); }
int finufft_execute_adjoint(finufft_plan p, void *c, void *fk) {
  return host_execute<double>(p, c, fk, true);
}
int finufftf_execute_adjoint(finufftf_plan p, void *c, void *fk) {
  return host_execute<float>(p, c, fk, true);
}
int finufft_destroy(finufft_plan p) { return host_destroy<double>(p); }
int finufftf_destroy(finufftf_plan p) { return host_destroy<float>(p); }

#define B200_HOST_SIMPLE_DEF(P, R)                                                               \
  int finufft##P##1d1many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,       \
                          int64_t ms, void *fk, finufft_opts *o) {                               \
    return host_simple<R>(1, 1, ntr, M, x, NUL, NUL, (void *)c, iflag, eps, ms, 1, 1, 0, NUL,    \
                          NUL, NUL, fk, o);                                                      \
  }                                                                                              \
  int finufft##P##1d1(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t ms,        \
                      void *fk, finufft_opts *o) {                                               \
    return finufft##P##1d1many(1, M, x, c, iflag, eps, ms, fk, o);                               \
  }                                                                                              \
  int finufft##P##1d2many(int ntr, int64_t M, const R *x, void *c, int iflag, R eps,             \
                          int64_t ms, const void *fk, finufft_opts *o) {                         \
    return host_simple<R>(1, 2, ntr, M, x, NUL, NUL, c, iflag, eps, ms, 1, 1, 0, NUL, NUL, NUL,  \
                          (void *)fk, o);                                                        \
  }                                                                                              \
  int finufft##P##1d2(int64_t M, const R *x, void *c, int iflag, R eps, int64_t ms,              \
                      const void *fk, finufft_opts *o) {                                         \
    return finufft##P##1d2many(1, M, x, c, iflag, eps, ms, fk, o);                               \
  }                                                                                              \
  int finufft##P##1d3many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,       \
                          int64_t nk, const R *s, void *fk, finufft_opts *o) {                   \
    return host_simple<R>(1, 3, ntr, M, x, NUL, NUL, (void *)c, iflag, eps, 1, 1, 1, nk, s, NUL, \
                          NUL, fk, o);                                                           \
  }                                                                                              \
  int finufft##P##1d3(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t nk,        \
                      const R *s, void *fk, finufft_opts *o) {                                   \
    return finufft##P##1d3many(1, M, x, c, iflag, eps, nk, s, fk, o);                            \
  }                                                                                              \
  int finufft##P##2d1many(int ntr, int64_t M, const R *x, const R *y, const void *c, int iflag,  \
                          R eps, int64_t ms, int64_t mt, void *fk, finufft_opts *o) {            \
    return host_simple<R>(2, 1, ntr, M, x, y, NUL, (void *)c, iflag, eps, ms, mt, 1, 0, NUL,     \
                          NUL, NUL, fk, o);                                                      \
  }                                                                                              \
  int finufft##P##2d1(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,        \
                      int64_t ms, int64_t mt, void *fk, finufft_opts *o) {                       \
    return finufft##P##2d1many(1, M, x, y, c, iflag, eps, ms, mt, fk, o);                        \
  }                                                                                              \
  int finufft##P##2d2many(int ntr, int64_t M, const R *x, const R *y, void *c, int iflag,        \
                          R eps, int64_t ms, int64_t mt, const void *fk, finufft_opts *o) {      \
    return host_simple<R>(2, 2, ntr, M, x, y, NUL, c, iflag, eps, ms, mt, 1, 0, NUL, NUL, NUL,   \
                          (void *)fk, o);                                                        \
  }                                                                                              \
  int finufft##P##2d2(int64_t M, const R *x, const R *y, void *c, int iflag, R eps, int64_t ms,  \
                      int64_t mt, const void *fk, finufft_opts *o) {                             \
    return finufft##P##2d2many(1, M, x, y, c, iflag, eps, ms, mt, fk, o);                        \
  }                                                                                              \
  int finufft##P##2d3many(int ntr, int64_t M, const R *x, const R *y, const void *c, int iflag,  \
                          R eps, int64_t nk, const R *s, const R *t, void *fk,                   \
                          finufft_opts *o) {                                                     \
    return host_simple<R>(2, 3, ntr, M, x, y, NUL, (void *)c, iflag, eps, 1, 1, 1, nk, s, t,     \
                          NUL, fk, o);                                                           \
  }                                                                                              \
  int finufft##P##2d3(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,        \
                      int64_t nk, const R *s, const R *t, void *fk, finufft_opts *o) {           \
    return finufft##P##2d3many(1, M, x, y, c, iflag, eps, nk, s, t, fk, o);                      \
  }                                                                                              \
  int finufft##P##3d1many(int ntr, int64_t M, const R *x, const R *y, const R *z,                \
                          const void *c, int iflag, R eps, int64_t ms, int64_t mt, int64_t mu,   \
                          void *fk, finufft_opts *o) {                                           \
    return host_simple<R>(3, 1, ntr, M, x, y, z, (void *)c, iflag, eps, ms, mt, mu, 0, NUL, NUL, \
                          NUL, fk, o);                                                           \
  }                                                                                              \
  int finufft##P##3d1(int64_t M, const R *x, const R *y, const R *z, const void *c, int iflag,   \
                      R eps, int64_t ms, int64_t mt, int64_t mu, void *fk, finufft_opts *o) {    \
    return finufft##P##3d1many(1, M, x, y, z, c, iflag, eps, ms, mt, mu, fk, o);                 \
  }                                                                                              \
  int finufft##P##3d2many(int ntr, int64_t M, const R *x, const R *y, const R *z, void *c,       \
                          int iflag, R eps, int64_t ms, int64_t mt, int64_t mu,                  \
                          const void *fk, finufft_opts *o) {                                     \
    return host_simple<R>(3, 2, ntr, M, x, y, z, c, iflag, eps, ms, mt, mu, 0, NUL, NUL, NUL,    \
                          (void *)fk, o);                                                        \
  }                                                                                              \
  int finufft##P##3d2(int64_t M, const R *x, const R *y, const R *z, void *c, int iflag,         \
                      R eps, int64_t ms, int64_t mt, int64_t mu, const void *fk,                 \
                      finufft_opts *o) {                                                         \
    return finufft##P##3d2many(1, M, x, y, z, c, iflag, eps, ms, mt, mu, fk, o);                 \
  }                                                                                              \
  int finufft##P##3d3many(int ntr, int64_t M, const R *x, const R *y, const R *z,                \
                          const void *c, int iflag, R eps, int64_t nk, const R *s, const R *t,   \
                          const R *u, void *fk, finufft_opts *o) {                               \
    return host_simple<R>(3, 3, ntr, M, x, y, z, (void *)c, iflag, eps, 1, 1, 1, nk, s, t, u,    \
                          fk, o);                                                                \
  }                                                                                              \
  int finufft##P##3d3(int64_t M, const R *x, const R *y, const R *z, const void *c, int iflag,   \
                      R eps, int64_t nk, const R *s, const R *t, const R *u, void *fk,           \
                      finufft_opts *o) {                                                         \
    return finufft##P##3d3many(1, M, x, y, z, c, iflag, eps, nk, s, t, u, fk, o);                \
  }
B200_HOST_SIMPLE_DEF(, double)
B200_HOST_SIMPLE_DEF(f, float)

// ------------------------------------------------------------------ introspection
int b200_get_plan_info(void *plan, b200_plan_info *out) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic || !out) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) fill_info<float>(as_plan<float>(plan)->eng, out);
    else fill_info<double>(as_plan<double>(plan)->eng, out);
  });
}
int b200_get_inner_plan_info(void *plan, b200_plan_info *out) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic || !out) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) {
      const Engine<float> *in = as_plan<float>(plan)->eng.inner();
      if (!in) throw Failure{ERR_PLAN_NOTVALID};
      fill_info<float>(*in, out);
    } else {
      const Engine<double> *in = as_plan<double>(plan)->eng.inner();
      if (!in) throw Failure{ERR_PLAN_NOTVALID};
      fill_info<double>(*in, out);
    }
  });
}
int b200_get_sort_permutation(void *plan, uint32_t *host_out) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) {
      DeviceGuard g(as_plan<float>(plan)->eng.opts.device);
      as_plan<float>(plan)->eng.copy_sort_to_host(host_out);
    } else {
      DeviceGuard g(as_plan<double>(plan)->eng.opts.device);
      as_plan<double>(plan)->eng.copy_sort_to_host(host_out);
    }
  });
}
int b200_get_raw_sort_order(void *plan, uint32_t *host_out) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) {
      DeviceGuard g(as_plan<float>(plan)->eng.opts.device);
      as_plan<float>(plan)->eng.copy_sort_to_host(host_out, true);
    } else {
      DeviceGuard g(as_plan<double>(plan)->eng.opts.device);
      as_plan<double>(plan)->eng.copy_sort_to_host(host_out, true);
    }
  });
}
int b200_get_sort_path(void *plan, int *path) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic || !path) throw Failure{ERR_PLAN_NOTVALID};
    *path = b->is_float ? as_plan<float>(plan)->eng.sort_path()
                        : as_plan<double>(plan)->eng.sort_path();
  });
}
int b200_get_window_table(void *plan, void *host_out) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) {
      auto &c = as_plan<float>(plan)->eng.coef;
      std::memcpy(host_out, c.data(), c.size() * sizeof(float));
    } else {
      auto &c = as_plan<double>(plan)->eng.coef;
      std::memcpy(host_out, c.data(), c.size() * sizeof(double));
    }
  });
}
int b200_get_phihat(void *plan, int d, void *host_out) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic || d < 0 || d > 2) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) {
      DeviceGuard g(as_plan<float>(plan)->eng.opts.device);
      as_plan<float>(plan)->eng.copy_phihat_to_host(d, (float *)host_out);
    } else {
      DeviceGuard g(as_plan<double>(plan)->eng.opts.device);
      as_plan<double>(plan)->eng.copy_phihat_to_host(d, (double *)host_out);
    }
  });
}
int b200_enable_profiling(void *plan, int on) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) as_plan<float>(plan)->eng.enable_profiling(on != 0);
    else as_plan<double>(plan)->eng.enable_profiling(on != 0);
  });
}
int b200_get_stage_ms(void *plan, float ms[5]) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic || !ms) throw Failure{ERR_PLAN_NOTVALID};
    if (b->is_float) as_plan<float>(plan)->eng.stage_ms(ms);
    else as_plan<double>(plan)->eng.stage_ms(ms);
  });
}
int b200_get_launch_count(void *plan, uint64_t *count) {
  return guarded([&] {
    auto *b = static_cast<PlanBase *>(plan);
    if (!b || b->magic != kMagic || !count) throw Failure{ERR_PLAN_NOTVALID};
    *count = b->is_float ? as_plan<float>(plan)->eng.launches : as_plan<double>(plan)->eng.launches;
  });
}
int b200_host_kernel(double tol, int dim, int type, double sigma, int is_float, int allow_small,
                     int *ns, double *beta, int *nc, void *coef) {
  return guarded([&] {
    double tol_used;
    int err = choose_kernel(tol, dim, type, sigma, is_float != 0, allow_small != 0, *ns, *beta,
                            tol_used);
    if (err) throw Failure{err};
    if (is_float) {
      std::vector<float> c;
      err = build_horner_table<float>(*ns, *beta, (float)tol_used, c, *nc);
      if (err) throw Failure{err};
      std::memcpy(coef, c.data(), c.size() * sizeof(float));
    } else {
      std::vector<double> c;
      err = build_horner_table<double>(*ns, *beta, tol_used, c, *nc);
      if (err) throw Failure{err};
      std::memcpy(coef, c.data(), c.size() * sizeof(double));
    }
  });
}
double b200_host_smallest_sigma(double tol, int dim, int type, int is_float, double maxN) {
  return smallest_feasible_sigma(tol, dim, type, is_float != 0, maxN);
}
int b200_host_sigma_feasible(double sigma, double tol, int dim, int type, int is_float,
                             double maxN) {
  return sigma_feasible(sigma, tol, dim, type, is_float != 0, maxN) ? 1 : 0;
}
double b200_host_choose_sigma(double tol, int dim, int type, int is_float, const int64_t *modes,
                              double npoints) {
  return choose_sigma(tol, dim, type, is_float != 0, modes, npoints);
}
int b200_host_sigma_candidates(double tol, int dim, int type, int is_float, double maxN,
                               double smax, double *sigma_out, int *ns_out, int cap) {
  return sigma_candidates(tol, dim, type, is_float != 0, maxN, smax, sigma_out, ns_out, cap);
}
double b200_host_choose_sigma_type3(double tol, int dim, int is_float, double nsources,
                                    double ntargets, const double *X, const double *S) {
  return choose_sigma_type3(tol, dim, is_float != 0, nsources, ntargets, X, S);
}
int64_t b200_host_fine_grid(double sigma, int64_t modes, int ns) {
  return fine_grid_size(sigma, modes, ns);
}
int b200_host_fseries(int64_t nf, int ns, int nc, int is_float, const void *coef, void *out) {
  return guarded([&] {
    if (is_float) {
      std::vector<float> ph;
      fseries_wound<float>(nf, ns, nc, (const float *)coef, ph);
      std::memcpy(out, ph.data(), ph.size() * sizeof(float));
    } else {
      std::vector<double> ph;
      fseries_wound<double>(nf, ns, nc, (const double *)coef, ph);
      std::memcpy(out, ph.data(), ph.size() * sizeof(double));
    }
  });
}
const char *b200_version(void) { return "finufft_b200 0.1 sm_100a"; }

}  // extern "C"
