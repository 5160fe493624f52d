/* ORACLE - TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for the dozen FFTW 3.3.10 entry points the reference's src/fft.cpp:40-222 calls
 * (FFTW is a third-party network fetch of the reference's build, cmake/setupFFTW.cmake, absent
 * here).  Signatures follow the published FFTW3 API; the transforms are computed by
 * oracle/fft_standin.hpp (plain mixed-radix FFT, same unnormalised definition).  Only what the
 * reference uses is supported: in-place, contiguous, batched c2c transforms of rank 1..3.
 */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#ifdef __cplusplus
extern "C" {
#endif

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

typedef double fftw_complex[2];
typedef float fftwf_complex[2];
typedef struct orc_fftw_plan_s *fftw_plan;
typedef struct orc_fftwf_plan_s *fftwf_plan;

#define ORC_FFTW_API(P, C, PLAN)                                                                 \
  PLAN P##plan_many_dft(int rank, const int *n, int howmany, C *in, const int *inembed,          \
                        int istride, int idist, C *out, const int *onembed, int ostride,         \
                        int odist, int sign, unsigned flags);                                    \
  void P##execute_dft(const PLAN p, C *in, C *out);                                              \
  void P##destroy_plan(PLAN p);                                                                  \
  int P##init_threads(void);                                                                     \
  void P##plan_with_nthreads(int nthreads);                                                      \
  void P##forget_wisdom(void);                                                                   \
  void P##cleanup(void);                                                                         \
  void P##cleanup_threads(void);
ORC_FFTW_API(fftw_, fftw_complex, fftw_plan)
ORC_FFTW_API(fftwf_, fftwf_complex, fftwf_plan)

#ifdef __cplusplus
}
#endif
#endif
