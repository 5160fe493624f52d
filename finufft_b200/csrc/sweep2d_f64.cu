// 2D row-sweep kernels, double precision (kernel widths 2..16).
#include "sweep2d_impl.cuh"

namespace b200 {

template<>
cudaError_t launch_spread2_sweep<double>(int ns, const Sweep2Points<double> &pts,
                                         const GridGeom<double> &g, int nc, const double *coef,
                                         const double2 *c_in, double2 *fw, cudaStream_t st) {
  using T             = double;
  constexpr bool spread = true;
  double2 *c_out      = nullptr;
  switch (ns) {
    B200_SWEEP2_CASE(2) B200_SWEEP2_CASE(3) B200_SWEEP2_CASE(4) B200_SWEEP2_CASE(5)
    B200_SWEEP2_CASE(6) B200_SWEEP2_CASE(7) B200_SWEEP2_CASE(8) B200_SWEEP2_CASE(9)
    B200_SWEEP2_CASE(10) B200_SWEEP2_CASE(11) B200_SWEEP2_CASE(12) B200_SWEEP2_CASE(13)
    B200_SWEEP2_CASE(14) B200_SWEEP2_CASE(15) B200_SWEEP2_CASE(16)
  default: return cudaErrorInvalidValue;
  }
}
template<>
cudaError_t launch_interp2_sweep<double>(int ns, const Sweep2Points<double> &pts,
                                         const GridGeom<double> &g, int nc, const double *coef,
                                         double2 *c_out, const double2 *fwc, cudaStream_t st) {
  using T             = double;
  constexpr bool spread = false;
  const double2 *c_in = nullptr;
  double2 *fw         = const_cast<double2 *>(fwc);
  switch (ns) {
    B200_SWEEP2_CASE(2) B200_SWEEP2_CASE(3) B200_SWEEP2_CASE(4) B200_SWEEP2_CASE(5)
    B200_SWEEP2_CASE(6) B200_SWEEP2_CASE(7) B200_SWEEP2_CASE(8) B200_SWEEP2_CASE(9)
    B200_SWEEP2_CASE(10) B200_SWEEP2_CASE(11) B200_SWEEP2_CASE(12) B200_SWEEP2_CASE(13)
    B200_SWEEP2_CASE(14) B200_SWEEP2_CASE(15) B200_SWEEP2_CASE(16)
  default: return cudaErrorInvalidValue;
  }
}

}  // namespace b200
