/* Introspection entry points of finufft_b200 (not part of the reference ABI): let tests and
 * benches read back what a plan decided and what setpts produced, so the CUDA path can be
 * checked stage by stage against the oracle.  `plan` is any handle returned by
 * cufinufft[f]_makeplan or finufft[f]_makeplan.
 */
#ifndef B200_INTROSPECT_H
#define B200_INTROSPECT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_plan_info {
  int is_float, type, dim, ntr, ns, nc, batch;
  double sigma, beta, tol;
  int64_t nf[3], ms[3], nbins[3];
  int64_t M, nsub;
} b200_plan_info;

int b200_get_plan_info(void *plan, b200_plan_info *out);
/* sort permutation (sorted position -> user index), M entries, to HOST memory */
int b200_get_sort_permutation(void *plan, uint32_t *host_out);
/* polynomial table, nc*ns entries of the plan's real type, row k = k-th highest degree */
int b200_get_window_table(void *plan, void *host_out);
/* window Fourier series of dimension d, nf[d]/2+1 entries of the plan's real type */
int b200_get_phihat(void *plan, int d, void *host_out);
/* library build tag, e.g. "finufft_b200 0.1 sm_100a" */
const char *b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
