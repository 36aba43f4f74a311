// Library-wide state: thread-local error text, launch counter, device info, and the small
// vector kernels (pow, dot, spmv) the Python drivers call directly.
#include "common.cuh"

namespace fes {

std::atomic<long long> g_launches{0};
static thread_local char t_error[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
  return code;
}

namespace {

__global__ void k_pow(const double* __restrict__ in, double p, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = pow(in[i], p);
}

// fixed-shape two-stage reduction: partial[b] = sum over a grid-stride slice, then one block
__global__ void k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                              double* __restrict__ partial) {
  __shared__ double sm[33];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += b ? a[i] * b[i] : a[i];
  s = block_sum(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void k_dot_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sm[33];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, sm);
  if (threadIdx.x == 0) *out = s;
}

// CSR y = A x, G lanes per row, shuffle-tree row sums
template <int G>
__global__ void k_spmv(int64_t n_rows, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const double* __restrict__ data, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / G;
  const int lane = (int)(t % G);
  double s = 0.0;
  if (row < n_rows) {
    const int32_t k1 = indptr[row + 1];
    for (int32_t k = indptr[row] + lane; k < k1; k += G) s += data[k] * __ldg(x + indices[k]);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (row < n_rows && lane == 0) y[row] = s;
}

// one thread per stored entry of B; B's (row, col) is located in A's row by binary search
__global__ void k_csr_add_sub(int64_t n_rows, const int32_t* __restrict__ ap, const int32_t* __restrict__ ai,
                              double* __restrict__ av, const int32_t* __restrict__ bp, const int32_t* __restrict__ bi,
                              const double* __restrict__ bv, double alpha, int* __restrict__ missing) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  for (int32_t k = bp[row]; k < bp[row + 1]; ++k) {
    const int32_t col = bi[k];
    int32_t lo = ap[row], hi = ap[row + 1];
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if (ai[mid] < col) lo = mid + 1; else hi = mid;
    }
    if (lo < ap[row + 1] && ai[lo] == col) av[lo] += alpha * bv[k];
    else atomicExch(missing, 1);
  }
}

}  // namespace
}  // namespace fes

using namespace fes;

extern "C" int fes_set_device(int device) {
  FES_CUDA(cudaSetDevice(device));
  return FES_OK;
}

extern "C" int fes_csr_add_subpattern(int64_t n_rows, const int32_t* ap, const int32_t* ai, double* av, const int32_t* bp,
                                      const int32_t* bi, const double* bv, double alpha, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_rows <= 0) return FES_OK;
  DevBuf<int> missing;
  FES_CUDA(missing.alloc(1));
  FES_CUDA(cudaMemsetAsync(missing.p, 0, sizeof(int), st));
  k_csr_add_sub<<<grid_for(n_rows, 256), 256, 0, st>>>(n_rows, ap, ai, av, bp, bi, bv, alpha, missing.p);
  FES_LAUNCH_CHECK();
  int h = 0;
  FES_CUDA(cudaMemcpyAsync(&h, missing.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  if (h) return set_error(FES_ERR_VALUE, "fes_csr_add_subpattern: B has an entry outside A's pattern");
  return FES_OK;
}

extern "C" int fes_version(void) { return 100; }
extern "C" const char* fes_last_error(void) { return t_error; }
extern "C" int64_t fes_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int fes_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes) {
  int dev = 0;
  FES_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FES_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
  return FES_OK;
}

extern "C" int fes_pow(const double* d_in, double p, int64_t n, double* d_out, void* stream) {
  if (n <= 0) return FES_OK;
  k_pow<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_in, p, n, d_out);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_dot(const double* d_a, const double* d_b, int64_t n, double* h_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!h_out) return set_error(FES_ERR_VALUE, "fes_dot: h_out is NULL");
  if (n <= 0) { *h_out = 0.0; return FES_OK; }
  const int blocks = 296;
  DevBuf<double> partial;
  FES_CUDA(partial.alloc(blocks + 1));
  k_dot_partial<<<blocks, 256, 0, st>>>(d_a, d_b, n, partial.p);
  FES_LAUNCH_CHECK();
  k_dot_final<<<1, 256, 0, st>>>(partial.p, blocks, partial.p + blocks);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaMemcpyAsync(h_out, partial.p + blocks, sizeof(double), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  return FES_OK;
}

extern "C" int fes_spmv(int64_t n_rows, const int32_t* d_indptr, const int32_t* d_indices, const double* d_data,
                        const double* d_x, double* d_y, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_rows <= 0) return FES_OK;
  k_spmv<8><<<grid_for(n_rows * 8, 256), 256, 0, st>>>(n_rows, d_indptr, d_indices, d_data, d_x, d_y);
  FES_LAUNCH_CHECK();
  return FES_OK;
}
