/* Host-pointer guru API of finufft_b200: the entry points a FINUFFT (CPU library) caller
 * binds, served by the GPU engine.  All pointers are HOST pointers; the library stages
 * host<->device copies itself on the plan's stream (pinned buffers make them asynchronous).
 *
 * Replaces reference include/finufft/finufft_eitherprec.h:55-67 (guru) and :72-151 (simple
 * interfaces), implemented there by src/c_interface.cpp.  Complex arrays are interleaved
 * (re,im) pairs (C99 complex / std::complex layout).  Return 0 or a FINUFFT_ERR_* code;
 * tol below what the precision/grid can deliver returns 26 unless
 * opts.allow_eps_too_small (reference include/finufft/setpts.hpp:29-53);
 * destroy(NULL) returns 1 (reference src/c_interface.cpp:94-95).
 */
#ifndef B200_FINUFFT_H
#define B200_FINUFFT_H
#include <stdint.h>

#include "b200_nufft_opts.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct finufft_plan_s *finufft_plan;   /* double */
typedef struct finufftf_plan_s *finufftf_plan; /* single */

/* reference finufft_eitherprec.h:55, defaults include/finufft/plan.hpp:294-335 */
void finufft_default_opts(finufft_opts *o);
void finufftf_default_opts(finufft_opts *o);
/* reference finufft_eitherprec.h:56-58 */
int finufft_makeplan(int type, int dim, const int64_t *n_modes, int iflag, int ntr, double tol,
                     finufft_plan *plan, finufft_opts *o);
int finufftf_makeplan(int type, int dim, const int64_t *n_modes, int iflag, int ntr, float tol,
                      finufftf_plan *plan, finufft_opts *o);
/* reference finufft_eitherprec.h:59-61 */
int finufft_setpts(finufft_plan plan, int64_t M, const double *x, const double *y,
                   const double *z, int64_t N, const double *s, const double *t,
                   const double *u);
int finufftf_setpts(finufftf_plan plan, int64_t M, const float *x, const float *y,
                    const float *z, int64_t N, const float *s, const float *t, const float *u);
/* reference finufft_eitherprec.h:62-66 */
int finufft_execute(finufft_plan plan, void *c, void *fk);
int finufftf_execute(finufftf_plan plan, void *c, void *fk);
int finufft_execute_adjoint(finufft_plan plan, void *c, void *fk);
int finufftf_execute_adjoint(finufftf_plan plan, void *c, void *fk);
/* reference finufft_eitherprec.h:67 */
int finufft_destroy(finufft_plan plan);
int finufftf_destroy(finufftf_plan plan);

/* Simple interfaces, reference finufft_eitherprec.h:72-151 (host pointers). */
#define B200_HOST_SIMPLE(P, R)                                                                   \
  int finufft##P##1d1(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t ms,        \
                      void *fk, finufft_opts *o);                                                \
  int finufft##P##1d1many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,       \
                          int64_t ms, void *fk, finufft_opts *o);                                \
  int finufft##P##1d2(int64_t M, const R *x, void *c, int iflag, R eps, int64_t ms,              \
                      const void *fk, finufft_opts *o);                                          \
  int finufft##P##1d2many(int ntr, int64_t M, const R *x, void *c, int iflag, R eps,             \
                          int64_t ms, const void *fk, finufft_opts *o);                          \
  int finufft##P##1d3(int64_t M, const R *x, const void *c, int iflag, R eps, int64_t nk,        \
                      const R *s, void *fk, finufft_opts *o);                                    \
  int finufft##P##1d3many(int ntr, int64_t M, const R *x, const void *c, int iflag, R eps,       \
                          int64_t nk, const R *s, void *fk, finufft_opts *o);                    \
  int finufft##P##2d1(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,        \
                      int64_t ms, int64_t mt, void *fk, finufft_opts *o);                        \
  int finufft##P##2d1many(int ntr, int64_t M, const R *x, const R *y, const void *c, int iflag,  \
                          R eps, int64_t ms, int64_t mt, void *fk, finufft_opts *o);             \
  int finufft##P##2d2(int64_t M, const R *x, const R *y, void *c, int iflag, R eps, int64_t ms,  \
                      int64_t mt, const void *fk, finufft_opts *o);                              \
  int finufft##P##2d2many(int ntr, int64_t M, const R *x, const R *y, void *c, int iflag,        \
                          R eps, int64_t ms, int64_t mt, const void *fk, finufft_opts *o);       \
  int finufft##P##2d3(int64_t M, const R *x, const R *y, const void *c, int iflag, R eps,        \
                      int64_t nk, const R *s, const R *t, void *fk, finufft_opts *o);            \
  int finufft##P##2d3many(int ntr, int64_t M, const R *x, const R *y, const void *c, int iflag,  \
                          R eps, int64_t nk, const R *s, const R *t, void *fk,                   \
                          finufft_opts *o);                                                      \
  int finufft##P##3d1(int64_t M, const R *x, const R *y, const R *z, const void *c, int iflag,   \
                      R eps, int64_t ms, int64_t mt, int64_t mu, void *fk, finufft_opts *o);     \
  int finufft##P##3d1many(int ntr, int64_t M, const R *x, const R *y, const R *z,                \
                          const void *c, int iflag, R eps, int64_t ms, int64_t mt, int64_t mu,   \
                          void *fk, finufft_opts *o);                                            \
  int finufft##P##3d2(int64_t M, const R *x, const R *y, const R *z, void *c, int iflag,         \
                      R eps, int64_t ms, int64_t mt, int64_t mu, const void *fk,                 \
                      finufft_opts *o);                                                          \
  int finufft##P##3d2many(int ntr, int64_t M, const R *x, const R *y, const R *z, void *c,       \
                          int iflag, R eps, int64_t ms, int64_t mt, int64_t mu,                  \
                          const void *fk, finufft_opts *o);                                      \
  int finufft##P##3d3(int64_t M, const R *x, const R *y, const R *z, const void *c, int iflag,   \
                      R eps, int64_t nk, const R *s, const R *t, const R *u, void *fk,           \
                      finufft_opts *o);                                                          \
  int finufft##P##3d3many(int ntr, int64_t M, const R *x, const R *y, const R *z,                \
                          const void *c, int iflag, R eps, int64_t nk, const R *s, const R *t,   \
                          const R *u, void *fk, finufft_opts *o);
B200_HOST_SIMPLE(, double)
B200_HOST_SIMPLE(f, float)
#undef B200_HOST_SIMPLE

#ifdef __cplusplus
}
#endif
#endif /* B200_FINUFFT_H */
