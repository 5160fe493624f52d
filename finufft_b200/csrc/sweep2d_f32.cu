// 2D row-sweep kernels, single precision (kernel widths 2..12).
#include "sweep2d_impl.cuh"

namespace b200 {

template<>
cudaError_t launch_spread2_sweep<float>(int ns, const Sweep2Points<float> &pts,
                                        const GridGeom<float> &g, int nc, const float *coef,
                                        const float2 *c_in, float2 *fw, cudaStream_t st) {
  using T             = float;
  constexpr bool spread = true;
  float2 *c_out       = nullptr;
  switch (ns) {
    B200_SWEEP2_CASE(2) B200_SWEEP2_CASE(3) B200_SWEEP2_CASE(4) B200_SWEEP2_CASE(5)
    B200_SWEEP2_CASE(6) B200_SWEEP2_CASE(7) B200_SWEEP2_CASE(8) B200_SWEEP2_CASE(9)
    B200_SWEEP2_CASE(10) B200_SWEEP2_CASE(11) B200_SWEEP2_CASE(12)
  default: return cudaErrorInvalidValue;
  }
}
template<>
cudaError_t launch_interp2_sweep<float>(int ns, const Sweep2Points<float> &pts,
                                        const GridGeom<float> &g, int nc, const float *coef,
                                        float2 *c_out, const float2 *fwc, cudaStream_t st) {
  using T             = float;
  constexpr bool spread = false;
  const float2 *c_in  = nullptr;
  float2 *fw          = const_cast<float2 *>(fwc);
  switch (ns) {
    B200_SWEEP2_CASE(2) B200_SWEEP2_CASE(3) B200_SWEEP2_CASE(4) B200_SWEEP2_CASE(5)
    B200_SWEEP2_CASE(6) B200_SWEEP2_CASE(7) B200_SWEEP2_CASE(8) B200_SWEEP2_CASE(9)
    B200_SWEEP2_CASE(10) B200_SWEEP2_CASE(11) B200_SWEEP2_CASE(12)
  default: return cudaErrorInvalidValue;
  }
}

}  // namespace b200
