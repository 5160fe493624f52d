/* The sharded guru sequence of include/b200_sharded.h from plain C: 3D type 1, single precision,
 * device arrays.  One process per GPU; this example is the single-rank case (world = 1, no NCCL id
 * needed) so that it runs anywhere; with more ranks every process makes the same calls after
 * sharing the 128-byte id of b200_slab_unique_id.
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/sharded3d1f.c -Lfinufft_b200 \
 *       -lfinufft_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/finufft_b200 -o sharded3d1f
 *
 * Prints the relative error of one mode against the direct sum; exit code 0 on success, the
 * library's error code otherwise (15 = no CUDA device: there is no CPU fallback).
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200_cufinufft.h"
#include "b200_sharded.h"

int main(void) {
  const int64_t M = 100000, N[3] = {32, 28, 24};
  const float tol = 1e-5f;
  float *x = malloc(sizeof(float) * M), *y = malloc(sizeof(float) * M),
        *z = malloc(sizeof(float) * M), *c = malloc(sizeof(float) * 2 * M);
  unsigned s = 4321u;
  for (int64_t j = 0; j < M; ++j) {
    s = s * 1664525u + 1013904223u; x[j] = (float)(M_PI * (2.0 * (s >> 8) / 16777216.0 - 1.0));
    s = s * 1664525u + 1013904223u; y[j] = (float)(M_PI * (2.0 * (s >> 8) / 16777216.0 - 1.0));
    s = s * 1664525u + 1013904223u; z[j] = (float)(M_PI * (2.0 * (s >> 8) / 16777216.0 - 1.0));
    s = s * 1664525u + 1013904223u; c[2 * j] = (float)(2.0 * (s >> 8) / 16777216.0 - 1.0);
    s = s * 1664525u + 1013904223u; c[2 * j + 1] = (float)(2.0 * (s >> 8) / 16777216.0 - 1.0);
  }
  cufinufft_opts opts;
  cufinufft_default_opts(&opts);
  b200_slabf_plan plan;
  int ier = b200_slabf_makeplan(1, N, +1, tol, /*rank*/ 0, /*world*/ 1, NULL, &opts, &plan);
  if (ier) { fprintf(stderr, "makeplan: error %d\n", ier); return ier; }
  b200_slab_info inf;
  b200_slab_get_info(plan, &inf);
  const int64_t nblock = N[0] * (inf.yhi - inf.ylo) * N[2];   /* this rank's fk[:, ylo:yhi, :] */
  float *dx, *dy, *dz, *dc, *dfk, *fk = malloc(sizeof(float) * 2 * nblock);
  if (cudaMalloc((void **)&dx, sizeof(float) * M) || cudaMalloc((void **)&dy, sizeof(float) * M) ||
      cudaMalloc((void **)&dz, sizeof(float) * M) || cudaMalloc((void **)&dc, sizeof(float) * 2 * M) ||
      cudaMalloc((void **)&dfk, sizeof(float) * 2 * nblock))
    return 15;
  cudaMemcpy(dx, x, sizeof(float) * M, cudaMemcpyHostToDevice);
  cudaMemcpy(dy, y, sizeof(float) * M, cudaMemcpyHostToDevice);
  cudaMemcpy(dz, z, sizeof(float) * M, cudaMemcpyHostToDevice);
  cudaMemcpy(dc, c, sizeof(float) * 2 * M, cudaMemcpyHostToDevice);
  ier = b200_slabf_setpts(plan, M, dx, dy, dz, /*routed*/ 0);
  if (ier) { fprintf(stderr, "setpts: error %d\n", ier); return ier; }
  ier = b200_slabf_execute(plan, dc, dfk);
  if (ier) { fprintf(stderr, "execute: error %d\n", ier); return ier; }
  cudaDeviceSynchronize();
  cudaMemcpy(fk, dfk, sizeof(float) * 2 * nblock, cudaMemcpyDeviceToHost);
  b200_slabf_destroy(plan);

  const int k1 = -5, k2 = 4, k3 = 9;   /* world = 1: the block is the whole mode array */
  double re = 0, im = 0;
  for (int64_t j = 0; j < M; ++j) {
    const double ph = k1 * (double)x[j] + k2 * (double)y[j] + k3 * (double)z[j];
    re += c[2 * j] * cos(ph) - c[2 * j + 1] * sin(ph);
    im += c[2 * j] * sin(ph) + c[2 * j + 1] * cos(ph);
  }
  const int64_t idx = (k1 + N[0] / 2) + N[0] * ((k2 + N[1] / 2) + N[1] * (int64_t)(k3 + N[2] / 2));
  const double er = fk[2 * idx] - re, ei = fk[2 * idx + 1] - im;
  const double rel = sqrt(er * er + ei * ei) / sqrt(re * re + im * im);
  printf("sharded (world 1) mode (%d,%d,%d): rel err %.2e\n", k1, k2, k3, rel);
  return rel < 1e-3 ? 0 : 100;
}
