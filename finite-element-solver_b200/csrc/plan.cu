// CSR sparsity pattern + scatter plan, built on the device once per (connectivity, n_comp).
//
// Replaces _ScatterPlan.build / FunctionSpace._scatter_plan (ref fem/space.py:107-131,757-772).
// The reference argsorts n_el*k^2 DOF-level int64 keys; here the sort runs on the c^2-times
// smaller NODE-level key set and the scalar pattern is expanded afterwards, which
// yields the identical indptr/indices:
//   1. key[e,a,b] = node[e,a] * n_nodes + node[e,b],  payload = e*N^2 + a*N + b
//   2. stable LSD radix sort by key  -> equal keys keep ascending (e,a,b) order, i.e. the
//      order np.bincount adds an element's contribution to a slot
//   3. head flags + inclusive scan    -> pair ids, pair_start offsets, node adjacency
//   4. per-node degree scan           -> node_ptr;  expansion by n_comp -> indptr / indices
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "plan.cuh"

namespace fes {
namespace {

constexpr int kBlock = 256;

__global__ void k_make_keys(const int32_t* __restrict__ en, int64_t n_contrib, int N, int64_t n_nodes,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_contrib) return;
  const int N2 = N * N;
  const int64_t e = t / N2;
  const int ab = (int)(t - e * N2), a = ab / N, b = ab - a * N;
  const int64_t v = en[e * N + a], w = en[e * N + b];
  if (v < 0 || v >= n_nodes) atomicExch(bad, 1);
  keys[t] = (uint64_t)v * (uint64_t)n_nodes + (uint64_t)w;
  vals[t] = (uint32_t)t;
}

__global__ void k_head_flags(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  flags[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1u : 0u;
}

// pair id of entry t is scan[t]-1; heads record the pair's first contribution and its (v, w)
__global__ void k_fill_pairs(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ scan, int64_t n,
                             int64_t n_nodes, uint32_t* __restrict__ pair_start, int32_t* __restrict__ pair_row,
                             int32_t* __restrict__ node_adj, int32_t* __restrict__ degree) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t id = scan[t];
  const bool head = (t == 0) || (scan[t - 1] != id);
  if (!head) return;
  const uint64_t key = keys[t];
  const int32_t v = (int32_t)(key / (uint64_t)n_nodes);
  pair_start[id - 1] = (uint32_t)t;
  pair_row[id - 1] = v;
  node_adj[id - 1] = (int32_t)(key - (uint64_t)v * (uint64_t)n_nodes);
  atomicAdd(degree + v, 1);  // integer atomics: the result does not depend on arrival order
}

// row c*v+d starts at c*c*node_ptr[v] + d*c*deg(v)
__global__ void k_expand_indptr(const int32_t* __restrict__ node_ptr, int64_t n_nodes, int c,
                                int32_t* __restrict__ indptr) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > n_nodes) return;
  if (v == n_nodes) {
    indptr[c * n_nodes] = c * c * node_ptr[n_nodes];
    return;
  }
  const int32_t p0 = node_ptr[v], deg = node_ptr[v + 1] - p0;
  for (int d = 0; d < c; ++d) indptr[c * v + d] = c * c * p0 + d * c * deg;
}

__global__ void k_expand_indices(const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ node_adj,
                                 const int32_t* __restrict__ pair_row, int64_t n_pairs, int c,
                                 int32_t* __restrict__ indices) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int32_t v = pair_row[p], w = node_adj[p];
  const int32_t p0 = node_ptr[v], deg = node_ptr[v + 1] - p0, s = (int32_t)(p - p0);
  for (int d = 0; d < c; ++d) {
    const int64_t base = (int64_t)c * c * p0 + (int64_t)d * c * deg + (int64_t)c * s;
    for (int j = 0; j < c; ++j) indices[base + j] = c * w + j;
  }
}

__global__ void k_widen(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = in[t];
}

int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

}  // namespace
}  // namespace fes

using namespace fes;

extern "C" int fes_plan_create(const int32_t* d_elem_nodes, int64_t n_el, int N, int64_t n_nodes, int n_comp,
                               void* stream_, fes_plan** out) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!out) return set_error(FES_ERR_VALUE, "fes_plan_create: out is NULL");
  *out = nullptr;
  if (n_el < 0 || n_nodes <= 0 || N < 2 || N > 6 || n_comp < 1 || n_comp > 3)
    return set_error(FES_ERR_VALUE, "fes_plan_create: bad sizes n_el=%lld N=%d n_nodes=%lld n_comp=%d",
                     (long long)n_el, N, (long long)n_nodes, n_comp);
  const int64_t M = n_el * N * N;
  if (M >= (int64_t)1 << 32)
    return set_error(FES_ERR_VALUE, "fes_plan_create: %lld element-block entries exceed 2^32 per GPU; shard the mesh",
                     (long long)M);
  if (n_nodes >= (int64_t)1 << 31) return set_error(FES_ERR_VALUE, "fes_plan_create: n_nodes exceeds int32");

  fes_plan* P = new fes_plan();
  struct Guard {
    fes_plan* p;
    ~Guard() { delete p; }
  } guard{P};
  P->n_el = n_el; P->n_nodes = n_nodes; P->N = N; P->c = n_comp;
  P->n_dofs = n_nodes * n_comp; P->n_contrib = M;

  FES_CUDA(P->node_ptr.alloc(n_nodes + 1));
  FES_CUDA(cudaMemsetAsync(P->node_ptr.p, 0, P->node_ptr.bytes(), stream));
  FES_CUDA(P->indptr.alloc(P->n_dofs + 1));

  if (M == 0) {  // no elements: an empty operator (ref: np.bincount(minlength=n) of nothing)
    FES_CUDA(cudaMemsetAsync(P->indptr.p, 0, P->indptr.bytes(), stream));
    FES_CUDA(P->pair_start.alloc(1));
    FES_CUDA(cudaMemsetAsync(P->pair_start.p, 0, sizeof(uint32_t), stream));
    FES_CUDA(cudaStreamSynchronize(stream));
    guard.p = nullptr;
    *out = P;
    return FES_OK;
  }

  DevBuf<uint64_t> keys_a, keys_b;
  DevBuf<uint32_t> vals_a, scan;
  DevBuf<int> bad;
  FES_CUDA(keys_a.alloc(M)); FES_CUDA(keys_b.alloc(M));
  FES_CUDA(vals_a.alloc(M)); FES_CUDA(P->contrib.alloc(M));
  FES_CUDA(bad.alloc(1));
  FES_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));

  k_make_keys<<<grid_for(M, kBlock), kBlock, 0, stream>>>(d_elem_nodes, M, N, n_nodes, keys_a.p, vals_a.p, bad.p);
  FES_LAUNCH_CHECK();
  int h_bad = 0;
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) return set_error(FES_ERR_VALUE, "fes_plan_create: element node index out of range [0, %lld)", (long long)n_nodes);

  // stable LSD radix sort over just the significant key bits
  const int end_bit = bits_for((uint64_t)n_nodes * (uint64_t)n_nodes - 1);
  size_t tmp_bytes = 0;
  FES_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a.p, keys_b.p, vals_a.p, P->contrib.p, M, 0, end_bit, stream));
  size_t scan_tmp = 0;
  FES_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, M, stream));
  DevBuf<uint8_t> tmp;
  FES_CUDA(tmp.alloc(tmp_bytes > scan_tmp ? tmp_bytes : scan_tmp));
  FES_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_a.p, keys_b.p, vals_a.p, P->contrib.p, M, 0, end_bit, stream));
  count_launch(2 * ((end_bit + 7) / 8));
  keys_a.release(); vals_a.release();

  FES_CUDA(scan.alloc(M));
  k_head_flags<<<grid_for(M, kBlock), kBlock, 0, stream>>>(keys_b.p, M, scan.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, scan_tmp, scan.p, scan.p, M, stream));
  count_launch(2);
  uint32_t n_pairs = 0;
  FES_CUDA(cudaMemcpyAsync(&n_pairs, scan.p + (M - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  P->n_pairs = n_pairs;
  P->nnz = (int64_t)n_pairs * n_comp * n_comp;
  if (P->nnz >= (int64_t)1 << 31)
    return set_error(FES_ERR_VALUE, "fes_plan_create: nnz=%lld exceeds int32 CSR offsets; shard the mesh", (long long)P->nnz);

  FES_CUDA(P->pair_start.alloc((size_t)n_pairs + 1));
  FES_CUDA(P->pair_row.alloc(n_pairs));
  FES_CUDA(P->node_adj.alloc(n_pairs));
  DevBuf<int32_t> degree;
  FES_CUDA(degree.alloc(n_nodes + 1));
  FES_CUDA(cudaMemsetAsync(degree.p, 0, degree.bytes(), stream));
  k_fill_pairs<<<grid_for(M, kBlock), kBlock, 0, stream>>>(keys_b.p, scan.p, M, n_nodes, P->pair_start.p,
                                                           P->pair_row.p, P->node_adj.p, degree.p);
  FES_LAUNCH_CHECK();
  const uint32_t M32 = (uint32_t)M;
  FES_CUDA(cudaMemcpyAsync(P->pair_start.p + n_pairs, &M32, sizeof(uint32_t), cudaMemcpyHostToDevice, stream));

  size_t scan2 = 0;
  FES_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan2, degree.p, P->node_ptr.p, n_nodes + 1, stream));
  if (scan2 > tmp.bytes()) FES_CUDA(tmp.alloc(scan2));
  FES_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan2, degree.p, P->node_ptr.p, n_nodes + 1, stream));
  count_launch(2);

  FES_CUDA(P->indices.alloc(P->nnz));
  k_expand_indptr<<<grid_for(n_nodes + 1, kBlock), kBlock, 0, stream>>>(P->node_ptr.p, n_nodes, n_comp, P->indptr.p);
  FES_LAUNCH_CHECK();
  k_expand_indices<<<grid_for(n_pairs, kBlock), kBlock, 0, stream>>>(P->node_ptr.p, P->node_adj.p, P->pair_row.p,
                                                                     n_pairs, n_comp, P->indices.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaStreamSynchronize(stream));
  guard.p = nullptr;
  *out = P;
  return FES_OK;
}

extern "C" void fes_plan_destroy(fes_plan* plan) { delete plan; }
extern "C" int64_t fes_plan_n_dofs(const fes_plan* p) { return p ? p->n_dofs : 0; }
extern "C" int64_t fes_plan_nnz(const fes_plan* p) { return p ? p->nnz : 0; }
extern "C" int64_t fes_plan_n_pairs(const fes_plan* p) { return p ? p->n_pairs : 0; }
extern "C" const int32_t* fes_plan_indptr(const fes_plan* p) { return p ? p->indptr.p : nullptr; }
extern "C" const int32_t* fes_plan_indices(const fes_plan* p) { return p ? p->indices.p : nullptr; }

extern "C" int64_t fes_plan_bytes(const fes_plan* p) {
  if (!p) return 0;
  return (int64_t)(p->node_ptr.bytes() + p->node_adj.bytes() + p->pair_row.bytes() + p->pair_start.bytes() +
                   p->contrib.bytes() + p->indptr.bytes() + p->indices.bytes());
}

// what one fused assembly streams from the plan: pair_row + pair_start + contrib + node_ptr + indptr
extern "C" int64_t fes_plan_read_bytes(const fes_plan* p) {
  if (!p) return 0;
  return (int64_t)(p->pair_row.bytes() + p->pair_start.bytes() + p->contrib.bytes() + p->node_ptr.bytes() +
                   p->indptr.bytes());
}

extern "C" int fes_plan_export_csr(const fes_plan* p, int64_t* h_indptr, int64_t* h_indices) {
  if (!p || !h_indptr || !h_indices) return set_error(FES_ERR_VALUE, "fes_plan_export_csr: NULL argument");
  const int64_t n1 = p->n_dofs + 1, big = n1 > p->nnz ? n1 : p->nnz;
  DevBuf<int64_t> wide;
  FES_CUDA(wide.alloc(big));
  k_widen<<<grid_for(n1, kBlock), kBlock>>>(p->indptr.p, n1, wide.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaMemcpy(h_indptr, wide.p, n1 * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (p->nnz > 0) {
    k_widen<<<grid_for(p->nnz, kBlock), kBlock>>>(p->indices.p, p->nnz, wide.p);
    FES_LAUNCH_CHECK();
    FES_CUDA(cudaMemcpy(h_indices, wide.p, p->nnz * sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  return FES_OK;
}
