// Microbenchmarks that size the spread kernel design: FFMA2 issue rate, global RED throughput.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ float2 ffma2s(float a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%2};\n mov.b64 rb, {%3,%4};\n mov.b64 rc, {%5,%6};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y) : "f"(a), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template<int MODE> __global__ void k_fma(float2 *out, const float *in, int iters) {
  float2 acc[16];
  float w[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(in[i], in[i + 16]);
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = in[32 + i + threadIdx.x % 3];
  float2 b = make_float2(in[40], in[41]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) acc[i] = ffma2s(w[r], b, acc[i]);
        else { acc[i].x = fmaf(w[r], b.x, acc[i].x); acc[i].y = fmaf(w[r], b.y, acc[i].y); }
      }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 16; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// RED patterns on a big grid. MODE 0: v2 all lanes contiguous; 1: v4 all lanes contiguous;
// 2: v2, 4 active lanes each on a different row (stride 4096 B); 3: v4 rows of 24 cells (12 lanes) x 2 rows
template<int MODE> __global__ void k_red(float *grid, size_t ncell, int iters) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarp = (gridDim.x * (size_t)blockDim.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    size_t base = ((warp + (size_t)it * nwarp) * 2654435761ull) % (ncell / 64) * 64;  // cell index, 512B aligned
    if (MODE == 0) {
      float2 *p = reinterpret_cast<float2 *>(grid) + base + lane;
      atomicAdd(p, make_float2(1.f, 2.f));
    } else if (MODE == 1) {
      float4 *p = reinterpret_cast<float4 *>(grid) + base / 2 + lane;
      atomicAdd(p, make_float4(1.f, 2.f, 3.f, 4.f));
    } else if (MODE == 2) {
      if (lane < 4) {
        float2 *p = reinterpret_cast<float2 *>(grid) + (base + (size_t)lane * 512) % ncell;
        atomicAdd(p, make_float2(1.f, 2.f));
      }
    } else {
      if (lane < 24) {
        float4 *p = reinterpret_cast<float4 *>(grid) + ((base + (size_t)(lane / 12) * 512) % ncell) / 2 + lane % 12;
        atomicAdd(p, make_float4(1.f, 2.f, 3.f, 4.f));
      }
    }
  }
}
template<class F> float timeit(F f, int rep = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < rep; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%s SMs=%d clk=%d kHz\n", pr.name, pr.multiProcessorCount, clk_khz);
  float *in; float2 *out; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, 148 * 1024 * 8 * 8);
  for (int wps = 1; wps <= 16; wps *= 2) {  // warps per SM = 4*wps
    int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = 148 * (128 * wps / threads);
    int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
      float ms = timeit([&] { if (mode == 0) k_fma<0><<<blocks, threads>>>(out, in, iters); else k_fma<1><<<blocks, threads>>>(out, in, iters); });
      double warp_instr = (double)iters * 64 * (mode == 0 ? 1 : 2) * (4.0 * wps) * 148;
      double cells = (double)iters * 64 * 32 * (4.0 * wps) * 148;
      printf("fma mode=%s warps/SM=%2d  %.3f ms  %.2f Gcell/s  -> warp-instr/ns/SM=%.3f\n", mode == 0 ? "FFMA2" : "FFMA ", 4 * wps, ms,
             cells / ms / 1e6, warp_instr / (ms * 1e6) / 148);
    }
  }
  size_t ncell = (size_t)512 * 512 * 512; float *grid; cudaMalloc(&grid, ncell * 8); cudaMemset(grid, 0, ncell * 8);
  int iters = 2000; int blocks = 148 * 8, threads = 256;
  double nw = (double)blocks * threads / 32 * iters;
  float ms;
  ms = timeit([&] { k_red<0><<<blocks, threads>>>(grid, ncell, iters); });
  printf("RED.v2 coalesced 32 lanes: %.3f ms, %.1f Gcell/s, payload %.1f GB/s\n", ms, nw * 32 / ms / 1e6, nw * 32 * 8 / ms / 1e6);
  ms = timeit([&] { k_red<1><<<blocks, threads>>>(grid, ncell, iters); });
  printf("RED.v4 coalesced 32 lanes: %.3f ms, %.1f Gcell/s, payload %.1f GB/s\n", ms, nw * 64 / ms / 1e6, nw * 64 * 8 / ms / 1e6);
  ms = timeit([&] { k_red<2><<<blocks, threads>>>(grid, ncell, iters); });
  printf("RED.v2 4 lanes 4 rows: %.3f ms, %.1f Gcell/s  (%.2f G warp-instr/s)\n", ms, nw * 4 / ms / 1e6, nw / ms / 1e6);
  ms = timeit([&] { k_red<3><<<blocks, threads>>>(grid, ncell, iters); });
  printf("RED.v4 24 lanes 2 rows: %.3f ms, %.1f Gcell/s\n", ms, nw * 48 / ms / 1e6);
  // L2-resident variant: small grid (32 MB)
  size_t small = (size_t)4 << 20;
  ms = timeit([&] { k_red<0><<<blocks, threads>>>(grid, small, iters); });
  printf("RED.v2 coalesced, 32MB footprint: %.3f ms, %.1f Gcell/s\n", ms, nw * 32 / ms / 1e6);
  ms = timeit([&] { k_red<1><<<blocks, threads>>>(grid, small, iters); });
  printf("RED.v4 coalesced, 32MB footprint: %.3f ms, %.1f Gcell/s\n", ms, nw * 64 / ms / 1e6);
  return 0;
}
