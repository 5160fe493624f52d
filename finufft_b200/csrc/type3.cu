// Type 3 (nonuniform -> nonuniform) composition.  Placeholder until the stage kernels land.
#include "engine.hpp"

namespace b200 {
template<class T>
void Engine<T>::setpts_type3(int64_t, const T *, const T *, const T *, int64_t, const T *,
                             const T *, const T *) {
  throw Failure{ERR_TYPE_NOTVALID};
}
template<class T> void Engine<T>::exec_type3(C *, C *) { throw Failure{ERR_TYPE_NOTVALID}; }
template void Engine<float>::setpts_type3(int64_t, const float *, const float *, const float *,
                                          int64_t, const float *, const float *, const float *);
template void Engine<double>::setpts_type3(int64_t, const double *, const double *,
                                           const double *, int64_t, const double *,
                                           const double *, const double *);
template void Engine<float>::exec_type3(float2 *, float2 *);
template void Engine<double>::exec_type3(double2 *, double2 *);
}  // namespace b200
