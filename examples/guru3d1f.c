/* The reference's guru call sequence (makeplan / setpts / execute / destroy) from plain C99,
 * bound to libfinufft_b200.so: 3D type 1, single precision, host arrays.  It is the program a
 * FINUFFT user already has (cf. reference examples/guru1d1.cpp, test/finufft3d_test.cpp); only the
 * library it links against changes.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/guru3d1f.c -Lfinufft_b200 -lfinufft_b200 -lm \
 *       -Wl,-rpath,$PWD/finufft_b200 -o guru3d1f && ./guru3d1f
 *
 * Prints the relative error of one mode against the direct sum; exit code 0 on success, the
 * library's error code otherwise (15 = no CUDA device: there is no CPU fallback).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200_finufft.h"

int main(void) {
  const int64_t M = 200000, N[3] = {40, 36, 30};
  const float tol = 1e-5f;
  float *x = malloc(sizeof(float) * M), *y = malloc(sizeof(float) * M),
        *z = malloc(sizeof(float) * M);
  float *c  = malloc(sizeof(float) * 2 * M);                      /* interleaved re,im */
  float *fk = malloc(sizeof(float) * 2 * N[0] * N[1] * N[2]);
  unsigned s = 12345u;
  for (int64_t j = 0; j < M; ++j) {
    s = s * 1664525u + 1013904223u; x[j] = (float)(M_PI * (2.0 * (s >> 8) / 16777216.0 - 1.0));
    s = s * 1664525u + 1013904223u; y[j] = (float)(M_PI * (2.0 * (s >> 8) / 16777216.0 - 1.0));
    s = s * 1664525u + 1013904223u; z[j] = (float)(M_PI * (2.0 * (s >> 8) / 16777216.0 - 1.0));
    s = s * 1664525u + 1013904223u; c[2 * j] = (float)(2.0 * (s >> 8) / 16777216.0 - 1.0);
    s = s * 1664525u + 1013904223u; c[2 * j + 1] = (float)(2.0 * (s >> 8) / 16777216.0 - 1.0);
  }
  finufft_opts opts;
  finufftf_default_opts(&opts);
  opts.upsampfac = 2.0;
  finufftf_plan plan;
  int ier = finufftf_makeplan(1, 3, N, +1, 1, tol, &plan, &opts);
  if (ier > 1) { fprintf(stderr, "makeplan: error %d\n", ier); return ier; }
  ier = finufftf_setpts(plan, M, x, y, z, 0, NULL, NULL, NULL);
  if (ier > 1) { fprintf(stderr, "setpts: error %d\n", ier); return ier; }
  ier = finufftf_execute(plan, c, fk);
  if (ier > 1) { fprintf(stderr, "execute: error %d\n", ier); return ier; }
  finufftf_destroy(plan);

  /* mode (k1,k2,k3) = (3,-7,5) against the direct sum; modes are stored k ascending, x fastest */
  const int k1 = 3, k2 = -7, k3 = 5;
  double re = 0, im = 0;
  for (int64_t j = 0; j < M; ++j) {
    const double ph = k1 * (double)x[j] + k2 * (double)y[j] + k3 * (double)z[j];
    re += c[2 * j] * cos(ph) - c[2 * j + 1] * sin(ph);
    im += c[2 * j] * sin(ph) + c[2 * j + 1] * cos(ph);
  }
  const int64_t idx = (k1 + N[0] / 2) + N[0] * ((k2 + N[1] / 2) + N[1] * (int64_t)(k3 + N[2] / 2));
  const double er = fk[2 * idx] - re, ei = fk[2 * idx + 1] - im;
  const double rel = sqrt(er * er + ei * ei) / sqrt(re * re + im * im);
  printf("mode (%d,%d,%d): %.6f%+.6fi, direct %.6f%+.6fi, rel err %.2e\n", k1, k2, k3,
         fk[2 * idx], fk[2 * idx + 1], re, im, rel);
  free(x); free(y); free(z); free(c); free(fk);
  return rel < 1e-3 ? 0 : 100;
}
