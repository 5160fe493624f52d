// extern "C" boundary of the sharded (multi-GPU) plans: include/b200_sharded.h.
#include <cuda_runtime.h>

#include <new>

#include "../../include/b200_cufinufft.h"
#include "../../include/b200_sharded.h"
#include "slab.hpp"

using namespace b200;

namespace {

constexpr uint32_t kSlabMagic = 0xB20051ABu;

struct SlabBase {
  uint32_t magic = kSlabMagic;
  bool is_float  = false;
  virtual ~SlabBase() { magic = 0; }
};
template<class T> struct SlabHandle : SlabBase {
  SlabPlan<T> plan;
  SlabHandle(int type, const int64_t *nm, int iflag, double tol, int rank, int world,
             const void *uid, const EngineOpts &o)
      : plan(type, nm, iflag, tol, rank, world, uid, o) {
    is_float = std::is_same<T, float>::value;
  }
};

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

EngineOpts slab_opts(const cufinufft_opts *o) {
  cufinufft_opts d;
  if (o) d = *o;
  else cufinufft_default_opts(&d);
  if (d.upsampfac != 0.0 && d.upsampfac <= 1.0) throw Failure{ERR_UPSAMPFAC_TOO_SMALL};
  EngineOpts e;
  e.upsampfac = d.upsampfac;
  e.device    = d.gpu_device_id;
  e.stream    = (cudaStream_t)d.gpu_stream;
  e.modeord   = d.modeord;
  e.maxsub    = d.gpu_maxsubprobsize > 0 ? d.gpu_maxsubprobsize : 1024;
  e.debug     = d.debug;
  return e;
}

template<class T> SlabPlan<T> &as_slab(void *p) {
  auto *b = static_cast<SlabBase *>(p);
  if (!b || b->magic != kSlabMagic || b->is_float != std::is_same<T, float>::value)
    throw Failure{ERR_PLAN_NOTVALID};
  return static_cast<SlabHandle<T> *>(b)->plan;
}

template<class T>
int make(int type, const int64_t *nm, int iflag, double tol, int rank, int world, const void *uid,
         const cufinufft_opts *o, void **out) {
  return guarded([&] {
    if (!out) throw Failure{ERR_INVALID_ARGUMENT};
    *out = nullptr;
    *out = static_cast<SlabBase *>(new SlabHandle<T>(type, nm, iflag, tol, rank, world, uid,
                                                    slab_opts(o)));
  });
}
template<class T> int destroy(void *p) {
  return guarded([&] {
    if (!p) throw Failure{ERR_PLAN_NOTVALID};
    as_slab<T>(p);
    delete static_cast<SlabBase *>(p);
  });
}

template<class S> void fill_info(S &s, b200_slab_info *o, int is_float) {
  o->is_float = is_float;
  o->type = s.type, o->rank = s.rank, o->world = s.world, o->ns = s.ns, o->mode = s.mode;
  for (int d = 0; d < 3; ++d) o->nf[d] = s.nf[d], o->ms[d] = s.ms[d];
  o->z0 = s.z0, o->nz = s.nz, o->ylo = s.ylo, o->yhi = s.yhi;
  o->win_org = s.win_org, o->win_n = s.win_n;
  o->M = s.M, o->M_local = s.Ml;
}
}  // namespace

extern "C" {

int b200_slab_unique_id(void *uid128) { return nccl_unique_id(uid128); }

int b200_slab_makeplan(int type, const int64_t *nm, int iflag, double eps, int rank, int world,
                       const void *uid, const cufinufft_opts *o, b200_slab_plan *plan) {
  return make<double>(type, nm, iflag, eps, rank, world, uid, o, (void **)plan);
}
int b200_slabf_makeplan(int type, const int64_t *nm, int iflag, float eps, int rank, int world,
                        const void *uid, const cufinufft_opts *o, b200_slabf_plan *plan) {
  return make<float>(type, nm, iflag, (double)eps, rank, world, uid, o, (void **)plan);
}
int b200_slab_setpts(b200_slab_plan p, int64_t M, const double *x, const double *y,
                     const double *z, int routed) {
  return guarded([&] { as_slab<double>(p).setpts(M, x, y, z, routed); });
}
int b200_slabf_setpts(b200_slabf_plan p, int64_t M, const float *x, const float *y,
                      const float *z, int routed) {
  return guarded([&] { as_slab<float>(p).setpts(M, x, y, z, routed); });
}
int b200_slab_execute(b200_slab_plan p, void *c, void *fk) {
  return guarded([&] { as_slab<double>(p).execute((double2 *)c, (double2 *)fk); });
}
int b200_slabf_execute(b200_slabf_plan p, void *c, void *fk) {
  return guarded([&] { as_slab<float>(p).execute((float2 *)c, (float2 *)fk); });
}
int b200_slab_gather_modes(b200_slab_plan p, const void *blk, void *full) {
  return guarded([&] { as_slab<double>(p).gather_modes((const double2 *)blk, (double2 *)full); });
}
int b200_slabf_gather_modes(b200_slabf_plan p, const void *blk, void *full) {
  return guarded([&] { as_slab<float>(p).gather_modes((const float2 *)blk, (float2 *)full); });
}
int b200_slab_slice_modes(b200_slab_plan p, const void *full, void *blk) {
  return guarded([&] { as_slab<double>(p).slice_modes((const double2 *)full, (double2 *)blk); });
}
int b200_slabf_slice_modes(b200_slabf_plan p, const void *full, void *blk) {
  return guarded([&] { as_slab<float>(p).slice_modes((const float2 *)full, (float2 *)blk); });
}
int b200_slab_destroy(b200_slab_plan p) { return destroy<double>(p); }
int b200_slabf_destroy(b200_slabf_plan p) { return destroy<float>(p); }

#define B200_EITHER(expr_f, expr_d)                                            \
  auto *b = static_cast<SlabBase *>(plan);                                     \
  if (!b || b->magic != kSlabMagic) throw Failure{ERR_PLAN_NOTVALID};          \
  if (b->is_float) { auto &s = as_slab<float>(plan); expr_f; }                 \
  else { auto &s = as_slab<double>(plan); expr_d; }

int b200_slab_get_info(void *plan, b200_slab_info *out) {
  return guarded([&] {
    if (!out) throw Failure{ERR_INVALID_ARGUMENT};
    B200_EITHER(fill_info(s, out, 1), fill_info(s, out, 0))
  });
}
int b200_slab_get_stage_ms(void *plan, float ms[10]) {
  return guarded([&] {
    if (!ms) throw Failure{ERR_INVALID_ARGUMENT};
    B200_EITHER(s.stage_ms(ms), s.stage_ms(ms))
  });
}
int b200_slab_get_launch_count(void *plan, uint64_t *count) {
  return guarded([&] {
    if (!count) throw Failure{ERR_INVALID_ARGUMENT};
    B200_EITHER(*count = s.launches + s.engine_launches(), *count = s.launches + s.engine_launches())
  });
}

}  // extern "C"
