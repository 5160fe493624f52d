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
/* type-3 plans after setpts: the inner type-2 plan (its sigma, width, grid); error 13 otherwise */
int b200_get_inner_plan_info(void *plan, b200_plan_info *out);
/* sort permutation (sorted position -> user index), M entries, to HOST memory */
int b200_get_sort_permutation(void *plan, uint32_t *host_out);
/* the order the kernels work in, as setpts left it on the device: bins ascending, inside a bin
 * (window class of the sweep kernels, user index) */
int b200_get_raw_sort_order(void *plan, uint32_t *host_out);
/* which sort the last setpts ran: 0 counting sort, 1 partition sort, 2 stable radix sort */
int b200_get_sort_path(void *plan, int *path);
/* polynomial table, nc*ns entries of the plan's real type, row k = k-th highest degree */
int b200_get_window_table(void *plan, void *host_out);
/* window Fourier series of dimension d, nf[d]/2+1 entries of the plan's real type */
int b200_get_phihat(void *plan, int d, void *host_out);
/* stage timing: CUDA events on the plan's stream around the stages of execute and setpts.
 * ms[0] spread|interp, ms[1] FFT (cuFFT), ms[2] deconvolve|amplify, ms[3] their sum, all for
 * the LAST execute (last batch); ms[4] the last setpts. */
int b200_enable_profiling(void *plan, int on);
int b200_get_stage_ms(void *plan, float ms[5]);
/* number of this library's own kernels launched by the plan so far (cuFFT, memset excluded) */
int b200_get_launch_count(void *plan, uint64_t *count);
/* Host-only plan mathematics (no GPU needed): what makeplan would decide.
 * b200_host_kernel: width ns, shape beta, polynomial table (nc*ns values of float or double
 * written to coef, which must hold 19*16 entries), returns 0 or a FINUFFT_ERR_* code.
 * b200_host_fine_grid: fine-grid length for `modes` modes (or -1 if above 1e12).
 * b200_host_fseries: window Fourier series k=0..nf/2 from a table, in the table's precision. */
int b200_host_kernel(double tol, int dim, int type, double sigma, int is_float, int allow_small,
                     int *ns, double *beta, int *nc, void *coef);
int64_t b200_host_fine_grid(double sigma, int64_t modes, int ns);
/* automatic upsampfac (host API, upsampfac = 0): smallest sigma the plan pipeline accepts
 * (reference src/common/kernel.cpp:231-257), the feasibility test itself (:203-228), and the
 * value setpts picks for npoints points on this device's cost model */
double b200_host_smallest_sigma(double tol, int dim, int type, int is_float, double maxN);
int b200_host_sigma_feasible(double sigma, double tol, int dim, int type, int is_float, double maxN);
double b200_host_choose_sigma(double tol, int dim, int type, int is_float, const int64_t *modes,
                              double npoints);
/* the candidate (sigma, ns) pairs the sigma search scores, in the reference's order
 * (include/finufft/heuristics.hpp:82-107; smax = 2.5 gives the reference's own set), and the
 * type-3 pick for nsources sources / ntargets targets with half-widths X[dim], S[dim]
 * (heuristics.hpp:130-150, applied at setpts when finufft_opts.upsampfac = 0) */
int b200_host_sigma_candidates(double tol, int dim, int type, int is_float, double maxN,
                               double smax, double *sigma_out, int *ns_out, int cap);
double b200_host_choose_sigma_type3(double tol, int dim, int is_float, double nsources,
                                    double ntargets, const double *X, const double *S);
int b200_host_fseries(int64_t nf, int ns, int nc, int is_float, const void *coef, void *out);
/* library build tag, e.g. "finufft_b200 0.1 sm_100a" */
const char *b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
