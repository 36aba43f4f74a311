// Grid-wide deterministic sums for the persistent cooperative kernels (PCG, OC bisection).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace fes {

constexpr int kMaxGrid = 148 * 8;  // partial-sum slots per reduced value

// NV values summed over the CTA in one pass (fixed tree); result valid in every thread.
// `sm` holds >= NV * 33 doubles.
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sm[k * 33 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nwarp ? sm[k * 33 + lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) sm[k * 33 + 32] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = sm[k * 33 + 32];
}

// Sum NV per-thread values over the whole grid; every thread of every CTA gets the same bits.
// `sm` holds >= NV * 33 doubles.
template <int NV>
__device__ __forceinline__ void grid_sum(cooperative_groups::grid_group& grid, double (&v)[NV], double* partials, int slot0,
                                         double* sm) {
  block_sum_n<NV>(v, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[(slot0 + k) * kMaxGrid + blockIdx.x] = v[k];
  }
  grid.sync();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += partials[(slot0 + k) * kMaxGrid + i];
    v[k] = s;
  }
  block_sum_n<NV>(v, sm);
}


}  // namespace fes
