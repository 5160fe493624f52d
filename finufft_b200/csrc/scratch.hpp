// Stream-ordered scratch memory for setpts, from a memory pool that belongs to this library.
//
// The plan's transient buffers (sort keys, partition records, work lists: several GB at
// M = 1e8) are allocated with cudaMallocFromPoolAsync from a private pool per device, so that
// the host application's own default pool is never reconfigured.  Freed blocks stay cached in
// the pool while at least one plan is alive on the device (repeated setpts calls then cost no
// driver allocations) and are returned to the driver when the last plan is destroyed.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "errors.hpp"

namespace b200 {

// the library's pool on `device` (created on first use); nullptr if pools are unsupported
cudaMemPool_t scratch_pool(int device);
// plan life-cycle hooks: the pool is trimmed to zero when the count drops back to 0
void scratch_pool_retain(int device);
void scratch_pool_release(int device);

template<class U> struct Scratch {
  U *p            = nullptr;
  cudaStream_t st = nullptr;
  Scratch(size_t n, cudaStream_t s, int device) : st(s) {
    if (!n) return;
    cudaMemPool_t pool = scratch_pool(device);
    cudaError_t e = pool ? cudaMallocFromPoolAsync((void **)&p, n * sizeof(U), pool, s)
                         : cudaMallocAsync((void **)&p, n * sizeof(U), s);
    if (e != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      throw Failure{e == cudaErrorMemoryAllocation ? ERR_ALLOC : ERR_CUDA_FAILURE};
    }
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, st);
  }
  Scratch(const Scratch &)            = delete;
  Scratch &operator=(const Scratch &) = delete;
};

}  // namespace b200
