// Point-driven spread / interp for plans whose caller asked for NO sort (gpu_sort = 0,
// spread_sort = 0): points are taken in the user's order, one thread per point, every stencil
// cell is one atomic add (spread) or one load (interp) on the fine grid.  This is the
// reference's unsorted mode: its CPU path then spreads with the identity permutation
// (include/finufft/spreadinterp.hpp:161-196), its GPU path runs the "nupts driven" method
// without bin sort (src/cuda/spread_nupts_driven_inst.cu:103-131).  Same arithmetic as every
// other kernel here: fold_rescale, stencil start ceil(X - ns/2), Horner windows (devmath.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "devmath.cuh"
#include "sort.cuh"

namespace b200 {

// coef = device copy of the plan's polynomial table, nc rows of ns, highest degree first
template<class T>
cudaError_t launch_direct(bool spread, int dim, int ns, int nc, const T *coef_dev, const T *x,
                          const T *y, const T *z, uint32_t M, const GridGeom<T> &g,
                          const typename CxOf<T>::type *c_in, typename CxOf<T>::type *c_out,
                          typename CxOf<T>::type *fw, cudaStream_t st);

}  // namespace b200
