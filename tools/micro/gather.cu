// Micro-benchmark: cost of the random 8-byte gather c[idx[i]] (the strength read of the spread
// kernels) and of the random 8-byte scatter (the result write of the interp kernels) under the
// L2 fetch-granularity limit.   nvcc -arch=sm_100a -O3 gather.cu -o gather && ./gather
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <algorithm>
#include <random>

__global__ void k_gather(const float2* __restrict__ c, const uint32_t* __restrict__ idx, float2* out, uint32_t n, int mode) {
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t j = idx[i];
    float2 v;
    if (mode == 0) v = c[j];
    else if (mode == 1) v = __ldcs(c + j);
    else if (mode == 2) v = __ldg(c + j);
    else { asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(c + j)); }
    out[i] = v;
  }
}
__global__ void k_scatter(const float2* __restrict__ c, const uint32_t* __restrict__ idx, float2* out, uint32_t n) {
  uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[idx[i]] = c[i];
}
int main() {
  const uint32_t n = 100000000u;
  std::vector<uint32_t> h(n);
  for (uint32_t i = 0; i < n; ++i) h[i] = i;
  std::mt19937_64 rng(1);
  std::shuffle(h.begin(), h.end(), rng);
  float2 *c, *out; uint32_t* idx;
  cudaMalloc(&c, n * 8ull); cudaMalloc(&out, n * 8ull); cudaMalloc(&idx, n * 4ull);
  cudaMemset(c, 0, n * 8ull);
  cudaMemcpy(idx, h.data(), n * 4ull, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
  printf("default cudaLimitMaxL2FetchGranularity = %zu\n", lim);
  for (size_t g : {(size_t)0, (size_t)32, (size_t)64, (size_t)128}) {
    if (g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("set %zu -> %s, now %zu\n", g, cudaGetErrorString(e), lim); }
    for (int mode = 0; mode < 4; ++mode) {
      k_gather<<<148 * 16, 256>>>(c, idx, out, n, mode);
      cudaEventRecord(e0);
      for (int r = 0; r < 3; ++r) k_gather<<<148 * 16, 256>>>(c, idx, out, n, mode);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("  gather mode %d: %.3f ms\n", mode, ms / 3);
    }
    k_scatter<<<148 * 16, 256>>>(c, idx, out, n);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) k_scatter<<<148 * 16, 256>>>(c, idx, out, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("  scatter: %.3f ms\n", ms / 3);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
