// Shared plumbing for libfes_b200: status codes, thread-local error text, launch accounting,
// RAII device buffers and the deterministic warp/block reductions every kernel uses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/fes_b200.h"

namespace fes {

int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define FES_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fes::set_error(FES_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                \
  } while (0)

#define FES_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != FES_OK) return _s; \
  } while (0)

#define FES_LAUNCH_CHECK()         \
  do {                             \
    fes::count_launch();           \
    FES_CUDA(cudaGetLastError()); \
  } while (0)

// Device allocation owned by a scope or a handle.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// ---- deterministic reductions -------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole CTA in a fixed tree; result valid in every thread. `smem` holds >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nwarp ? smem[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) smem[32] = t;
  }
  __syncthreads();
  return smem[32];
}

}  // namespace fes
