/* Multi-GPU entry points of finufft_b200: one large 3D type-1 / type-2 transform sharded over
 * the GPUs of one box by z-slabs of the fine grid, one process (or thread) per GPU, NCCL over
 * NVLink for the exchanges (csrc/slab.hpp describes the pipeline).
 *
 * No reference interface exists for this: the reference only runs independent plans per device
 * (cufinufft_opts.gpu_device_id, include/cufinufft_opts.h:7-40; test/cuda/
 * cufinufft_multigpu_test.cu:29-132).  The calls below keep the guru shape of
 * include/cufinufft.h:16-38 (makeplan / setpts / execute / destroy, same n_modes order, iflag,
 * tol, cufinufft_opts, device pointers, error codes of include/finufft_errors.h:9-44) and add
 * what sharding needs: a rank, a world size and the 128-byte NCCL unique id shared by the ranks.
 *
 * Sharded data layout (x fastest, like the reference):
 *   points   every rank passes its own M points (any M, also 0); strengths / values in the same
 *            order.  routed = 1 promises that every point already folds into the rank's slab
 *            (error 21 otherwise); routed = 0 lets setpts route the points (or, for clustered
 *            input, replicate a window of the grid and reduce it: b200_slab_info.mode = 1).
 *   modes    rank r holds fk[:, ylo:yhi, :], i.e. ms3 x (yhi-ylo) x ms1 values, the y range
 *            being an even split of ms2 over the ranks (b200_slab_info).  b200_slab_gather_modes
 *            assembles the full ms3 x ms2 x ms1 array on every rank, b200_slab_slice_modes cuts
 *            a rank's block out of a full array.
 * Exchanges: when the ranks are separate processes with peer access (one NVLink / NVSwitch box),
 * every rank maps the others' buffers with CUDA IPC and the kernels load / store peer memory
 * directly (ghost planes added from the neighbours' windows, cropped modes stored into the
 * owners' pencils, strengths stored into the owners' arrays), ordered by one-word all-reduces;
 * otherwise grouped ncclSend/ncclRecv.  B200_NUFFT_SLAB_P2P=0 forces the NCCL path.
 * makeplan, setpts, execute, gather_modes and destroy are collective over the ranks.
 * ntransf = 1.  Batched transforms shard by vectors instead (no communication): give each rank
 * an ordinary plan for its slice of the vectors.
 */
#ifndef B200_SHARDED_H
#define B200_SHARDED_H
#include <stdint.h>

#include "b200_nufft_opts.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_slab_plan_s *b200_slab_plan;   /* double precision */
typedef struct b200_slab_fplan_s *b200_slabf_plan; /* single precision */

typedef struct b200_slab_info {
  int is_float, type, rank, world, ns, mode; /* mode 0: z-slabs, 1: replicated window */
  int64_t nf[3], ms[3];
  int64_t z0, nz;        /* fine-grid planes this rank owns */
  int64_t ylo, yhi;      /* y range of the modes this rank holds */
  int64_t win_org, win_n; /* planes of the spread / interp window (global origin, count) */
  int64_t M, M_local;    /* points given by the caller, points this rank spreads */
} b200_slab_info;

/* 128 bytes for rank 0 to create and hand to every rank (ncclGetUniqueId) */
int b200_slab_unique_id(void *uid128);

/* collective over the `world` ranks; uid128 may be NULL when world = 1 */
int b200_slab_makeplan(int type, const int64_t n_modes[3], int iflag, double eps, int rank,
                       int world, const void *uid128, const cufinufft_opts *opts,
                       b200_slab_plan *plan);
int b200_slabf_makeplan(int type, const int64_t n_modes[3], int iflag, float eps, int rank,
                        int world, const void *uid128, const cufinufft_opts *opts,
                        b200_slabf_plan *plan);
/* collective; device pointers */
int b200_slab_setpts(b200_slab_plan plan, int64_t M, const double *d_x, const double *d_y,
                     const double *d_z, int routed);
int b200_slabf_setpts(b200_slabf_plan plan, int64_t M, const float *d_x, const float *d_y,
                      const float *d_z, int routed);
/* collective; type 1: d_c in, d_fk_block out; type 2: d_fk_block in, d_c out */
int b200_slab_execute(b200_slab_plan plan, void *d_c, void *d_fk_block);
int b200_slabf_execute(b200_slabf_plan plan, void *d_c, void *d_fk_block);
/* collective */
int b200_slab_gather_modes(b200_slab_plan plan, const void *d_fk_block, void *d_fk_full);
int b200_slabf_gather_modes(b200_slabf_plan plan, const void *d_fk_block, void *d_fk_full);
/* local */
int b200_slab_slice_modes(b200_slab_plan plan, const void *d_fk_full, void *d_fk_block);
int b200_slabf_slice_modes(b200_slabf_plan plan, const void *d_fk_full, void *d_fk_block);
int b200_slab_destroy(b200_slab_plan plan);
int b200_slabf_destroy(b200_slabf_plan plan);

/* either precision */
int b200_slab_get_info(void *plan, b200_slab_info *out);
/* CUDA-event times (ms) of the stages of the last execute: [0] spread|interp, [1] ghost planes
 * (or window reduce / broadcast), [2] 2D FFT, [3] pack|unpack, [4] slab<->pencil transpose,
 * [5] 1D FFT, [6] deconvolve|amplify, [7] routing of strengths / values, [8] their sum; [9]
 * the last setpts including the routing of the coordinates */
int b200_slab_get_stage_ms(void *plan, float ms[10]);
/* kernels of this library launched by the plan so far (cuFFT, NCCL, copies excluded) */
int b200_slab_get_launch_count(void *plan, uint64_t *count);

#ifdef __cplusplus
}
#endif
#endif
