// Hand-written device scans and a stable LSD radix sort for the pattern builder (plan.cu) and the filter
// build (simp.cu) -- the "radix sort plus warp-level scan" of the sparsity-pattern kernel set.  No library
// kernels: every launch below is one of ours.
//
//   device_scan<T, INCLUSIVE>   three-kernel scan: per-block scan of 2048-element tiles (a lane scans its 8
//                               items, warps combine with shuffle scans, the block with a second shuffle scan
//                               over the warp totals), a recursive scan of the block totals, an offset pass.
//                               In-place (in == out) is allowed.
//   radix_sort_pairs<K>         8-bit digits, least significant first, over the significant key bits only.
//                               Per pass: per-block digit histograms -> one exclusive scan over the
//                               digit-major [256][n_blocks] table (gives every block its stable base per
//                               digit) -> scatter, where the rank of a key inside its block is
//                               (keys of earlier rounds) + (earlier warps of this round) + (earlier lanes of
//                               this warp, __match_any_sync + popc).  Equal keys keep their input order,
//                               which is what makes the sorted contribution list follow np.bincount's
//                               summation order.
#pragma once
#include <vector>

#include "common.cuh"

namespace fes {

constexpr int kScanBlock = 256, kScanItems = 8, kScanTile = kScanBlock * kScanItems;

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T up = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += up;
  }
  return v;
}

// scan of one tile per block; block_sums[b] = the tile's total (NULL on the last level)
template <typename T, bool INCLUSIVE>
__global__ void __launch_bounds__(kScanBlock)
k_scan_tiles(const T* __restrict__ in, T* __restrict__ out, int64_t n, T* __restrict__ block_sums) {
  __shared__ T s_warp[kScanBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  T v[kScanItems];
  T run = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : (T)0;
    run += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T incl = warp_inclusive_scan(run);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < kScanBlock / 32 ? s_warp[lane] : (T)0;
    w = warp_inclusive_scan(w);
    if (lane < kScanBlock / 32) s_warp[lane] = w;
  }
  __syncthreads();
  T prefix = (incl - run) + (warp > 0 ? s_warp[warp - 1] : (T)0);  // exclusive prefix of this thread's first item
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const T x = v[k];
    if (base + k < n) out[base + k] = INCLUSIVE ? prefix + x : prefix;
    prefix += x;
  }
  if (block_sums && threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = prefix;
}

template <typename T>
__global__ void __launch_bounds__(kScanBlock)
k_scan_offset(T* __restrict__ out, int64_t n, const T* __restrict__ block_prefix) {
  const T add = block_prefix[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  for (int k = threadIdx.x; k < kScanTile; k += kScanBlock)
    if (base + k < n) out[base + k] += add;
}

template <typename T, bool INCLUSIVE>
int device_scan(const T* in, T* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return FES_OK;
  const int64_t nb = (n + kScanTile - 1) / kScanTile;
  if (nb == 1) {
    k_scan_tiles<T, INCLUSIVE><<<1, kScanBlock, 0, st>>>(in, out, n, nullptr);
    FES_LAUNCH_CHECK();
    return FES_OK;
  }
  DevBuf<T> sums;
  FES_CUDA(sums.alloc(nb));
  k_scan_tiles<T, INCLUSIVE><<<(unsigned)nb, kScanBlock, 0, st>>>(in, out, n, sums.p);
  FES_LAUNCH_CHECK();
  FES_TRY((device_scan<T, false>(sums.p, sums.p, nb, st)));  // exclusive prefix of every tile
  k_scan_offset<T><<<(unsigned)nb, kScanBlock, 0, st>>>(out, n, sums.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaStreamSynchronize(st));  // `sums` is freed on return
  return FES_OK;
}

// ---- radix sort ---------------------------------------------------------------------------------
constexpr int kSortBlock = 256, kSortRounds = 16, kSortTile = kSortBlock * kSortRounds;

template <typename K>
__global__ void __launch_bounds__(kSortBlock)
k_radix_hist(const K* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ hist, int64_t nb) {
  __shared__ uint32_t s_hist[256];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + (int64_t)r * kSortBlock + threadIdx.x;
    if (i < n) atomicAdd(&s_hist[(uint32_t)(keys[i] >> shift) & 255u], 1u);  // counts only: order is irrelevant here
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nb + blockIdx.x] = s_hist[threadIdx.x];  // digit-major: one scan orders digits, then blocks
}

template <typename K>
__global__ void __launch_bounds__(kSortBlock)
k_radix_scatter(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
                uint32_t* __restrict__ vals_out, int64_t n, int shift, const uint32_t* __restrict__ offsets, int64_t nb) {
  constexpr int kWarps = kSortBlock / 32;
  __shared__ uint32_t s_base[256];             // next free output slot of each digit for this block
  __shared__ uint32_t s_count[kWarps][256];    // this round's keys per (warp, digit)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s_base[threadIdx.x] = offsets[(int64_t)threadIdx.x * nb + blockIdx.x];
#pragma unroll
  for (int w = 0; w < kWarps; ++w) s_count[w][threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int r = 0; r < kSortRounds; ++r) {
    const int64_t i = base + (int64_t)r * kSortBlock + threadIdx.x;
    const bool valid = i < n;
    K key = 0;
    uint32_t val = 0;
    if (valid) { key = keys_in[i]; val = vals_in[i]; }
    // an invalid lane takes a digit nobody else has, so it matches only itself
    const uint32_t digit = valid ? ((uint32_t)(key >> shift) & 255u) : (256u + (uint32_t)lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) s_count[warp][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t before = 0;
      for (int w = 0; w < warp; ++w) before += s_count[w][digit];
      const uint32_t dst = s_base[digit] + before + rank_in_warp;
      keys_out[dst] = key;
      vals_out[dst] = val;
    }
    __syncthreads();
    {  // thread d closes digit d of this round
      uint32_t tot = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) { tot += s_count[w][threadIdx.x]; s_count[w][threadIdx.x] = 0; }
      s_base[threadIdx.x] += tot;
    }
    __syncthreads();
  }
}

// Sort (keys, vals) by the low `end_bit` key bits, stable.  Ping-pongs between the a and b buffers; on
// return *result_in_b says where the sorted arrays live.  n < 2^32.
template <typename K>
int radix_sort_pairs(K* keys_a, K* keys_b, uint32_t* vals_a, uint32_t* vals_b, int64_t n, int end_bit, cudaStream_t st,
                     bool* result_in_b) {
  *result_in_b = false;
  if (n <= 0) return FES_OK;
  const int64_t nb = (n + kSortTile - 1) / kSortTile;
  DevBuf<uint32_t> hist;
  FES_CUDA(hist.alloc((size_t)256 * nb));
  bool in_b = false;
  for (int shift = 0; shift < end_bit; shift += 8) {
    K* ki = in_b ? keys_b : keys_a;
    K* ko = in_b ? keys_a : keys_b;
    uint32_t* vi = in_b ? vals_b : vals_a;
    uint32_t* vo = in_b ? vals_a : vals_b;
    k_radix_hist<K><<<(unsigned)nb, kSortBlock, 0, st>>>(ki, n, shift, hist.p, nb);
    FES_LAUNCH_CHECK();
    FES_TRY((device_scan<uint32_t, false>(hist.p, hist.p, 256 * nb, st)));
    k_radix_scatter<K><<<(unsigned)nb, kSortBlock, 0, st>>>(ki, vi, ko, vo, n, shift, hist.p, nb);
    FES_LAUNCH_CHECK();
    in_b = !in_b;
  }
  FES_CUDA(cudaStreamSynchronize(st));  // `hist` is freed on return
  *result_in_b = in_b;
  return FES_OK;
}

}  // namespace fes
