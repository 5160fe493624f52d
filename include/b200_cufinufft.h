/* Device-pointer guru API of finufft_b200: the entry points a cuFINUFFT caller binds.
 *
 * Replaces, symbol for symbol, reference include/cufinufft.h:16-38 (guru calls) and :45-186
 * (36 one-shot wrappers), implemented there by src/cuda/c_interface.cpp:33-174,188-470.
 * Semantics kept: type 1/2/3, dim 1..3, iflag>=0 means +i, n_modes x-fastest, arrays for
 * ntr>1 stacked transform-slowest, all pointers are DEVICE pointers on opts.gpu_device_id,
 * work is issued on opts.gpu_stream, return value 0 or a FINUFFT_ERR_* code
 * (dim -> 12, bad n_modes / M > INT32_MAX -> 14, type -> 10, ntr<1 -> 9, sigma<=1 -> 7,
 * CUDA failure / bad device -> 15, destroy(NULL) -> 16, type-3 NULL s/t/u -> 21).
 * Complex arrays are interleaved (re,im) pairs: cuFloatComplex / cuDoubleComplex layout.
 */
#ifndef B200_CUFINUFFT_H
#define B200_CUFINUFFT_H
#include <stdint.h>

#include "b200_nufft_opts.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cufinufft_plan_s *cufinufft_plan;   /* double precision plan handle */
typedef struct cufinufft_fplan_s *cufinufftf_plan; /* single precision plan handle */

/* reference cufinufft.h:16, defaults src/cuda/c_interface.cpp:134-174 */
void cufinufft_default_opts(cufinufft_opts *opts);

/* reference cufinufft.h:18-23 */
int cufinufft_makeplan(int type, int dim, const int64_t *n_modes, int iflag, int ntr, double eps,
                       cufinufft_plan *plan, const cufinufft_opts *opts);
int cufinufftf_makeplan(int type, int dim, const int64_t *n_modes, int iflag, int ntr, float eps,
                        cufinufftf_plan *plan, const cufinufft_opts *opts);
/* reference cufinufft.h:25-30 */
int cufinufft_setpts(cufinufft_plan plan, int64_t M, const double *d_x, const double *d_y,
                     const double *d_z, int N, const double *d_s, const double *d_t,
                     const double *d_u);
int cufinufftf_setpts(cufinufftf_plan plan, int64_t M, const float *d_x, const float *d_y,
                      const float *d_z, int N, const float *d_s, const float *d_t,
                      const float *d_u);
/* reference cufinufft.h:32-35; type 1,3: d_c in, d_fk out; type 2: d_fk in, d_c out */
int cufinufft_execute(cufinufft_plan plan, void *d_c, void *d_fk);
int cufinufftf_execute(cufinufftf_plan plan, void *d_c, void *d_fk);
/* reference cufinufft.h:37-38 */
int cufinufft_destroy(cufinufft_plan plan);
int cufinufftf_destroy(cufinufftf_plan plan);

/* One-shot wrappers, reference cufinufft.h:45-186: makeplan + setpts + execute + destroy.
 * Naming cufinufft[f]<dim>d<type>[many]; pointers are device pointers. */
#define B200_CU_SIMPLE(P, R)                                                                    \
  int cufinufft##P##1d1many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,     \
                            int64_t ms, void *fk, const cufinufft_opts *o);                      \
  int cufinufft##P##1d1(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t ms,      \
                        void *fk, const cufinufft_opts *o);                                      \
  int cufinufft##P##1d2many(int ntr, int64_t M, const R *x, void *c, int iflag, R eps,           \
                            int64_t ms, const void *fk, const cufinufft_opts *o);                \
  int cufinufft##P##1d2(int64_t M, const R *x, void *c, int iflag, R eps, int64_t ms,            \
                        const void *fk, const cufinufft_opts *o);                                \
  int cufinufft##P##1d3many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,     \
                            int64_t nk, const R *s, void *fk, const cufinufft_opts *o);          \
  int cufinufft##P##1d3(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t nk,      \
                        const R *s, void *fk, const cufinufft_opts *o);                          \
  int cufinufft##P##2d1many(int ntr, int64_t M, const R *x, const R *y, const void *c,           \
                            int iflag, R eps, int64_t ms, int64_t mt, void *fk,                  \
                            const cufinufft_opts *o);                                            \
  int cufinufft##P##2d1(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,      \
                        int64_t ms, int64_t mt, void *fk, const cufinufft_opts *o);              \
  int cufinufft##P##2d2many(int ntr, int64_t M, const R *x, const R *y, void *c, int iflag,      \
                            R eps, int64_t ms, int64_t mt, const void *fk,                       \
                            const cufinufft_opts *o);                                            \
  int cufinufft##P##2d2(int64_t M, const R *x, const R *y, void *c, int iflag, R eps,            \
                        int64_t ms, int64_t mt, const void *fk, const cufinufft_opts *o);        \
  int cufinufft##P##2d3many(int ntr, int64_t M, const R *x, const R *y, const void *c,           \
                            int iflag, R eps, int64_t nk, const R *s, const R *t, void *fk,      \
                            const cufinufft_opts *o);                                            \
  int cufinufft##P##2d3(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,      \
                        int64_t nk, const R *s, const R *t, void *fk, const cufinufft_opts *o);  \
  int cufinufft##P##3d1many(int ntr, int64_t M, const R *x, const R *y, const R *z,              \
                            const void *c, int iflag, R eps, int64_t ms, int64_t mt,             \
                            int64_t mu, void *fk, const cufinufft_opts *o);                      \
  int cufinufft##P##3d1(int64_t M, const R *x, const R *y, const R *z, const void *c,            \
                        int iflag, R eps, int64_t ms, int64_t mt, int64_t mu, void *fk,          \
                        const cufinufft_opts *o);                                                \
  int cufinufft##P##3d2many(int ntr, int64_t M, const R *x, const R *y, const R *z, void *c,     \
                            int iflag, R eps, int64_t ms, int64_t mt, int64_t mu,                \
                            const void *fk, const cufinufft_opts *o);                            \
  int cufinufft##P##3d2(int64_t M, const R *x, const R *y, const R *z, void *c, int iflag,       \
                        R eps, int64_t ms, int64_t mt, int64_t mu, const void *fk,               \
                        const cufinufft_opts *o);                                                \
  int cufinufft##P##3d3many(int ntr, int64_t M, const R *x, const R *y, const R *z,              \
                            const void *c, int iflag, R eps, int64_t nk, const R *s,             \
                            const R *t, const R *u, void *fk, const cufinufft_opts *o);          \
  int cufinufft##P##3d3(int64_t M, const R *x, const R *y, const R *z, const void *c,            \
                        int iflag, R eps, int64_t nk, const R *s, const R *t, const R *u,        \
                        void *fk, const cufinufft_opts *o);
B200_CU_SIMPLE(, double)
B200_CU_SIMPLE(f, float)
#undef B200_CU_SIMPLE

#ifdef __cplusplus
}
#endif
#endif /* B200_CUFINUFFT_H */
