/* Option structs and return codes of the finufft_b200 C ABI.
 *
 * Both structs are byte-for-byte layout compatible with the reference library's public
 * option structs, so a caller compiled against the reference headers (or the ctypes /
 * MATLAB / Julia mirrors of them) can pass its struct to this library unchanged:
 *   cufinufft_opts  <->  reference include/cufinufft_opts.h:7-40
 *                        (mirrored by python/cufinufft/cufinufft/_cufinufft.py:90-110)
 *   finufft_opts    <->  reference include/finufft_opts.h:32-71
 * The include guards are the reference's own, so whichever header is seen first wins and
 * the two can never collide in one translation unit.
 *
 * Fields this engine honours are marked [used]; the rest are accepted and ignored (they
 * tune the reference's own kernels; this engine has exactly one method).
 */
#ifndef B200_NUFFT_OPTS_H
#define B200_NUFFT_OPTS_H

#ifndef __CUFINUFFT_OPTS_H__
#define __CUFINUFFT_OPTS_H__
typedef struct cufinufft_opts {
  double upsampfac;         /* [used] sigma; 0 = choose (2.0) */
  int gpu_method;           /* ignored: one method only */
  int gpu_sort;             /* 1 bin-sort at setpts, 0 keep the user's order (point-driven kernels) */
  int gpu_binsizex;         /* ignored: bins are the CPU library's 16 x 4 x 4 */
  int gpu_binsizey;
  int gpu_binsizez;
  int gpu_obinsizex;        /* ignored */
  int gpu_obinsizey;
  int gpu_obinsizez;
  int gpu_maxsubprobsize;   /* [used] most points of one bin handled by one warp */
  int gpu_kerevalmeth;      /* ignored: piecewise-polynomial evaluation always */
  int gpu_spreadinterponly; /* [used] 1 = spread (type 1) / interpolate (type 2) only */
  int gpu_maxbatchsize;     /* [used] transforms per FFT batch; 0 = min(ntransf, 8) */
  int gpu_device_id;        /* [used] device the plan lives on */
  void *gpu_stream;         /* [used] cudaStream_t all work is issued on */
  int modeord;              /* [used] 0 ascending modes, 1 FFT order */
  int gpu_np;               /* ignored */
  int debug;                /* [used] 1 prints plan parameters */
} cufinufft_opts;
#endif

#ifndef FINUFFT_OPTS_H
#define FINUFFT_OPTS_H
typedef struct finufft_opts {
  int modeord;              /* [used] */
  int spreadinterponly;     /* [used] */
  int debug;                /* [used] */
  int spread_debug;         /* ignored */
  int showwarn;             /* [used] warnings on stderr */
  int nthreads;             /* ignored: no host threading */
  int fftw;                 /* ignored: cuFFT */
  int spread_sort;          /* 0 keep the user's order, 1 sort, 2 library's choice (= sort) */
  int spread_kerevalmeth;   /* ignored */
  int spread_kerpad;        /* ignored */
  double upsampfac;         /* [used] 0 = choose (2.0) */
  int spread_thread;        /* ignored */
  int maxbatchsize;         /* [used] */
  int spread_nthr_atomic;   /* ignored */
  int spread_max_sp_size;   /* [used] if >0: most points of one bin handled by one warp */
  int spread_kerformula;    /* 0 only */
  int allow_eps_too_small;  /* [used] 0: error 26 when tol is unreachable, 1: proceed */
  void (*fftw_lock_fun)(void *);   /* ignored */
  void (*fftw_unlock_fun)(void *); /* ignored */
  void *fftw_lock_data;            /* ignored */
} finufft_opts;
#endif

/* Return codes: numerically identical to reference include/finufft_errors.h:9-44. */
#ifndef FINUFFT_ERRORS_H
#define FINUFFT_ERRORS_H
enum {
  FINUFFT_ERR_MAXNALLOC           = 2,
  FINUFFT_ERR_SPREAD_BOX_SMALL    = 3,
  FINUFFT_ERR_UPSAMPFAC_TOO_SMALL = 7,
  FINUFFT_ERR_NTRANS_NOTVALID     = 9,
  FINUFFT_ERR_TYPE_NOTVALID       = 10,
  FINUFFT_ERR_ALLOC               = 11,
  FINUFFT_ERR_DIM_NOTVALID        = 12,
  FINUFFT_ERR_NDATA_NOTVALID      = 14,
  FINUFFT_ERR_CUDA_FAILURE        = 15,
  FINUFFT_ERR_PLAN_NOTVALID       = 16,
  FINUFFT_ERR_INSUFFICIENT_SHMEM  = 19,
  FINUFFT_ERR_NUM_NU_PTS_INVALID  = 20,
  FINUFFT_ERR_INVALID_ARGUMENT    = 21,
  FINUFFT_ERR_UNKNOWN_EXCEPTION   = 25,
  FINUFFT_ERR_EPS_TOO_SMALL       = 26,
  FINUFFT_ERR_PSWF_SETUP          = 27
};
#endif

#endif /* B200_NUFFT_OPTS_H */
