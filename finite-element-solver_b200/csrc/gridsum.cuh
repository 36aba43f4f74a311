// Grid-wide deterministic sums for the persistent cooperative kernels (PCG, OC bisection).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace fes {

constexpr int kMaxGrid = 148 * 8;  // partial-sum slots per reduced value

// Sum NV per-thread values over the whole grid; every thread of every CTA gets the same bits.
template <int NV>
__device__ __forceinline__ void grid_sum(cooperative_groups::grid_group& grid, double (&v)[NV], double* partials, int slot0,
                                         double* sm) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = block_sum(v[k], sm);
    if (threadIdx.x == 0) partials[(slot0 + k) * kMaxGrid + blockIdx.x] = s;
  }
  grid.sync();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += partials[(slot0 + k) * kMaxGrid + i];
    v[k] = block_sum(s, sm);
  }
}


}  // namespace fes
