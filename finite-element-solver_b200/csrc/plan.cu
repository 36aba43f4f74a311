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

#include <stdlib.h>

#include <algorithm>
#include <vector>

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

// incident-element list of node v = the contributions of its diagonal pair (v, v): every element
// containing v appears there once, in ascending element id, with a == b == v's local index
__global__ void k_diag_counts(const int32_t* __restrict__ pair_row, const int32_t* __restrict__ node_adj,
                              const uint32_t* __restrict__ pair_start, int64_t n_pairs,
                              int32_t* __restrict__ ne_cnt) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  if (pair_row[p] == node_adj[p]) ne_cnt[pair_row[p]] = (int32_t)(pair_start[p + 1] - pair_start[p]);
}

__global__ void k_fill_ne(const int32_t* __restrict__ pair_row, const int32_t* __restrict__ node_adj,
                          const uint32_t* __restrict__ pair_start, const uint32_t* __restrict__ contrib,
                          int64_t n_pairs, int N, const int32_t* __restrict__ ne_ptr,
                          uint32_t* __restrict__ ne_code) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs || pair_row[p] != node_adj[p]) return;
  const int N2 = N * N;
  const int32_t base = ne_ptr[pair_row[p]];
  const uint32_t k0 = pair_start[p], k1 = pair_start[p + 1];
  for (uint32_t k = k0; k < k1; ++k) {
    const uint32_t code = contrib[k], e = code / N2, a = (code - e * N2) / N;
    ne_code[base + (k - k0)] = e * 8u + a;
  }
}

// per pair: tile-relative row and contribution offset; per contribution: the tile-local record
// index of (row node, element) packed with the local node indices
__global__ void k_tile_codes(const int32_t* __restrict__ pair_row, const uint32_t* __restrict__ pair_start,
                             const uint32_t* __restrict__ contrib, int64_t n_pairs, int N,
                             const int32_t* __restrict__ ne_ptr, const uint32_t* __restrict__ ne_code,
                             const int32_t* __restrict__ tile_node0, int64_t n_tiles,
                             uint16_t* __restrict__ contrib16, uint8_t* __restrict__ pair_lrow) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int32_t v = pair_row[p];
  int64_t lo = 0, hi = n_tiles;  // last tile with tile_node0[t] <= v
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (tile_node0[mid] <= v) lo = mid; else hi = mid;
  }
  const int32_t v0 = tile_node0[lo];
  const int32_t g0 = ne_ptr[v0], gv0 = ne_ptr[v], gv1 = ne_ptr[v + 1];
  const int N2 = N * N;
  pair_lrow[p] = (uint8_t)(v - v0);
  const uint32_t k0 = pair_start[p], k1 = pair_start[p + 1];
  for (uint32_t k = k0; k < k1; ++k) {
    const uint32_t code = contrib[k], e = code / N2, ab = code - e * N2, a = ab / N, b = ab - a * N;
    int32_t l = gv0, h = gv1;  // lower_bound of e in v's ascending incident list
    while (l < h) {
      const int32_t m = (l + h) >> 1;
      if ((ne_code[m] >> 3) < e) l = m + 1; else h = m;
    }
    contrib16[k] = (uint16_t)(((uint32_t)(l - g0) << 6) | (a << 3) | b);
  }
}

// Active pairs of a tile (w >= v): count them and their contributions; flag anything the packed
// metadata cannot hold.
__global__ void k_tile_count(const int32_t* __restrict__ tile_node0, int64_t n_tiles,
                             const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ node_adj,
                             const int32_t* __restrict__ pair_row, const uint32_t* __restrict__ pair_start,
                             int32_t* __restrict__ act_cnt, int32_t* __restrict__ con_cnt, int* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int32_t p0 = node_ptr[tile_node0[t]], p1 = node_ptr[tile_node0[t + 1]];
  int32_t na = 0, nc = 0;
  for (int32_t p = p0; p < p1; ++p) {
    if (node_adj[p] < pair_row[p]) continue;
    const uint32_t c = pair_start[p + 1] - pair_start[p];
    if (c > 255u) atomicExch(bad, 1);
    ++na;
    nc += (int32_t)c;
  }
  if (nc > 65535) atomicExch(bad, 1);
  act_cnt[t] = na;
  con_cnt[t] = nc;
}

// Emit a tile's active pairs in descending contribution count (stable), so the 32 pairs a warp
// gathers have equal trip counts, with their packed metadata and their contribution codes.
__global__ void k_tile_active(const int32_t* __restrict__ tile_node0, int64_t n_tiles,
                              const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ node_adj,
                              const int32_t* __restrict__ pair_row, const uint32_t* __restrict__ pair_start,
                              const uint16_t* __restrict__ contrib16, const int32_t* __restrict__ tile_act0,
                              const int32_t* __restrict__ tile_con0, int c, uint4* __restrict__ meta,
                              uint16_t* __restrict__ codes) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int32_t p0 = node_ptr[tile_node0[t]], p1 = node_ptr[tile_node0[t + 1]];
  uint32_t hi = 0;
  for (int32_t p = p0; p < p1; ++p)
    if (node_adj[p] >= pair_row[p]) hi = max(hi, pair_start[p + 1] - pair_start[p]);
  int32_t wa = tile_act0[t];
  const int32_t con_base = tile_con0[t];
  uint32_t off = 0;
  while (hi > 0) {
    uint32_t next = 0;
    for (int32_t p = p0; p < p1; ++p) {
      const int32_t v = pair_row[p], w = node_adj[p];
      if (w < v) continue;
      const uint32_t k0 = pair_start[p], cnt = pair_start[p + 1] - k0;
      if (cnt == hi) {
        const int32_t npv = node_ptr[v], degv = node_ptr[v + 1] - npv;
        const int32_t npw = node_ptr[w], degw = node_ptr[w + 1] - npw;
        int32_t lo_ = npw, hi_ = npw + degw;  // slot of v in row w (the pattern is symmetric)
        while (lo_ < hi_) {
          const int32_t m = (lo_ + hi_) >> 1;
          if (node_adj[m] < v) lo_ = m + 1; else hi_ = m;
        }
        uint4 mt;
        mt.x = off | (cnt << 16) | ((w == v ? 1u : 0u) << 24);
        mt.y = (uint32_t)c * (uint32_t)npv + (uint32_t)(p - npv);
        mt.z = (uint32_t)c * (uint32_t)npw + (uint32_t)(lo_ - npw);
        mt.w = (uint32_t)degv | ((uint32_t)degw << 16);
        meta[wa++] = mt;
        for (uint32_t k = 0; k < cnt; ++k) codes[con_base + off + k] = contrib16[k0 + k];
        off += cnt;
      } else if (cnt < hi) {
        next = max(next, cnt);
      }
    }
    hi = next;
  }
}

// One CTA per tile: collect the corner nodes of the tile's records, sort them (bitonic, shared
// memory), keep the distinct ones, and re-express every record's corners as indices into that list.
constexpr int kTileIdCap = 1024;  // records per tile (<= 256) x corner nodes (<= 4)
__global__ void __launch_bounds__(256)
k_tile_nodes(const int32_t* __restrict__ tile_node0, const int32_t* __restrict__ ne_ptr,
             const uint32_t* __restrict__ ne_code, const int32_t* __restrict__ en, int N, int n_corner,
             int32_t* __restrict__ tile_nodes, int32_t* __restrict__ tile_nnod, uint32_t* __restrict__ rec_local,
             int* __restrict__ bad) {
  __shared__ uint32_t ids[kTileIdCap];
  __shared__ uint32_t uniq[kTileIdCap];
  __shared__ int s_count;
  const int t = blockIdx.x;
  const int32_t g0 = ne_ptr[tile_node0[t]], g1 = ne_ptr[tile_node0[t + 1]];
  const int n_ids = (g1 - g0) * n_corner;
  for (int i = threadIdx.x; i < kTileIdCap; i += blockDim.x) {
    uint32_t v = 0xffffffffu;
    if (i < n_ids) {
      const int r = i / n_corner, k = i - r * n_corner;
      v = (uint32_t)en[(int64_t)(ne_code[g0 + r] >> 3) * N + k];
    }
    ids[i] = v;
  }
  __syncthreads();
  for (int k = 2; k <= kTileIdCap; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < kTileIdCap; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const uint32_t a = ids[i], b = ids[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { ids[i] = b; ids[l] = a; }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x == 0) {  // <= 1024 entries: a serial compaction is cheap next to the sort
    int c = 0;
    for (int i = 0; i < n_ids; ++i)
      if (i == 0 || ids[i] != ids[i - 1]) uniq[c++] = ids[i];
    s_count = c;
    tile_nnod[t] = c;
    if (c > 255) atomicExch(bad, 1);
  }
  __syncthreads();
  const int nd = s_count;
  if (nd > 255) return;
  for (int i = threadIdx.x; i < nd; i += blockDim.x) tile_nodes[(int64_t)t * fes_plan::kTileNodeStride + i] = (int32_t)uniq[i];
  for (int r = threadIdx.x; r < g1 - g0; r += blockDim.x) {
    const int64_t e = ne_code[g0 + r] >> 3;
    uint32_t packed = 0;
    for (int k = 0; k < n_corner; ++k) {
      const uint32_t node = (uint32_t)en[e * N + k];
      int lo = 0, hi = nd;
      while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (uniq[m] < node) lo = m + 1; else hi = m;
      }
      packed |= (uint32_t)lo << (8 * k);
    }
    rec_local[g0 + r] = packed;
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

// Tile the node range for the fused assembly kernel.  The greedy split runs on the host over the
// two prefix arrays (O(n_nodes), once per mesh); everything else stays on the device.
int build_tiles(fes_plan* P, const int32_t* d_elem_nodes, cudaStream_t stream, int gmax, bool* retry) {
  *retry = false;
  const int64_t nn = P->n_nodes, np_ = P->n_pairs;
  const int N = P->N;
  P->tiled = false;
  if (np_ == 0) return FES_OK;
  DevBuf<int32_t> ne_cnt;
  FES_CUDA(ne_cnt.alloc(nn + 1));
  FES_CUDA(cudaMemsetAsync(ne_cnt.p, 0, ne_cnt.bytes(), stream));
  k_diag_counts<<<grid_for(np_, kBlock), kBlock, 0, stream>>>(P->pair_row.p, P->node_adj.p, P->pair_start.p, np_, ne_cnt.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(P->ne_ptr.alloc(nn + 1));
  size_t tb = 0;
  FES_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, ne_cnt.p, P->ne_ptr.p, nn + 1, stream));
  DevBuf<uint8_t> tmp;
  FES_CUDA(tmp.alloc(tb));
  FES_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, ne_cnt.p, P->ne_ptr.p, nn + 1, stream));
  count_launch(2);
  std::vector<int32_t> h_ne(nn + 1), h_np(nn + 1);
  FES_CUDA(cudaMemcpyAsync(h_ne.data(), P->ne_ptr.p, (nn + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaMemcpyAsync(h_np.data(), P->node_ptr.p, (nn + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  P->n_geo = h_ne[nn];
  if (P->n_geo != P->n_el * N) return FES_OK;  // degenerate elements (repeated nodes): keep the untiled kernel
  if (P->n_el >= ((int64_t)1 << 29)) return FES_OK;
  P->geo_max = gmax;
  std::vector<int32_t> tiles;
  tiles.push_back(0);
  int64_t v0 = 0;
  for (int64_t v = 0; v < nn; ++v) {
    const int64_t geo = h_ne[v + 1] - h_ne[v0];
    if (h_ne[v + 1] - h_ne[v] > gmax) return FES_OK;  // one node alone overflows a tile: untiled kernel
    if (geo > gmax || v - v0 >= 255) {
      tiles.push_back((int32_t)v);
      v0 = v;
    }
  }
  tiles.push_back((int32_t)nn);
  P->n_tiles = (int64_t)tiles.size() - 1;
  FES_CUDA(P->tile_node0.alloc(tiles.size()));
  FES_CUDA(cudaMemcpyAsync(P->tile_node0.p, tiles.data(), tiles.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  FES_CUDA(P->ne_code.alloc(P->n_geo));
  k_fill_ne<<<grid_for(np_, kBlock), kBlock, 0, stream>>>(P->pair_row.p, P->node_adj.p, P->pair_start.p, P->contrib.p, np_,
                                                          N, P->ne_ptr.p, P->ne_code.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(P->contrib16.alloc(P->n_contrib + 16));
  FES_CUDA(P->pair_lrow.alloc(np_ + 32));
  k_tile_codes<<<grid_for(np_, kBlock), kBlock, 0, stream>>>(P->pair_row.p, P->pair_start.p, P->contrib.p, np_, N,
                                                             P->ne_ptr.p, P->ne_code.p, P->tile_node0.p, P->n_tiles,
                                                             P->contrib16.p, P->pair_lrow.p);
  FES_LAUNCH_CHECK();

  // active pairs (w >= v) per tile: counts -> offsets -> sorted emission with packed metadata
  const int64_t nt = P->n_tiles;
  DevBuf<int32_t> act_cnt, con_cnt;
  DevBuf<int> bad;
  FES_CUDA(act_cnt.alloc(nt + 1)); FES_CUDA(con_cnt.alloc(nt + 1)); FES_CUDA(bad.alloc(1));
  FES_CUDA(cudaMemsetAsync(act_cnt.p, 0, act_cnt.bytes(), stream));
  FES_CUDA(cudaMemsetAsync(con_cnt.p, 0, con_cnt.bytes(), stream));
  FES_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));
  k_tile_count<<<grid_for(nt, 64), 64, 0, stream>>>(P->tile_node0.p, nt, P->node_ptr.p, P->node_adj.p, P->pair_row.p,
                                                    P->pair_start.p, act_cnt.p, con_cnt.p, bad.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(P->tile_act0.alloc(nt + 1));
  FES_CUDA(P->tile_con0.alloc(nt + 1));
  size_t sb = 0;
  FES_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, act_cnt.p, P->tile_act0.p, nt + 1, stream));
  if (sb > tmp.bytes()) FES_CUDA(tmp.alloc(sb));
  FES_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, sb, act_cnt.p, P->tile_act0.p, nt + 1, stream));
  FES_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, sb, con_cnt.p, P->tile_con0.p, nt + 1, stream));
  count_launch(4);
  std::vector<int32_t> h_act(nt + 1), h_con(nt + 1);
  int h_bad = 0;
  FES_CUDA(cudaMemcpyAsync(h_act.data(), P->tile_act0.p, (nt + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaMemcpyAsync(h_con.data(), P->tile_con0.p, (nt + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) return FES_OK;  // a pair shared by > 255 elements: keep the untiled kernel
  P->n_active = h_act[nt];
  P->n_act_contrib = h_con[nt];
  int amax = 0, cmax = 0, dmax = 0;
  for (int64_t t = 0; t < nt; ++t) {
    amax = std::max(amax, h_act[t + 1] - h_act[t]);
    cmax = std::max(cmax, h_con[t + 1] - h_con[t]);
  }
  for (int64_t v = 0; v < nn; ++v) dmax = std::max(dmax, h_np[v + 1] - h_np[v]);
  if (dmax > 65535) return FES_OK;
  P->act_max = amax; P->con_max = cmax;
  FES_CUDA(P->act_meta.alloc(P->n_active + 4));
  FES_CUDA(P->act_codes.alloc(P->n_act_contrib + 16));
  k_tile_active<<<grid_for(nt, 64), 64, 0, stream>>>(P->tile_node0.p, nt, P->node_ptr.p, P->node_adj.p, P->pair_row.p,
                                                     P->pair_start.p, P->contrib16.p, P->tile_act0.p, P->tile_con0.p,
                                                     P->c, P->act_meta.p, P->act_codes.p);
  FES_LAUNCH_CHECK();

  // distinct nodes per tile + local corner indices per record
  P->n_corner = (N == 6) ? 3 : N;
  if (gmax * P->n_corner > kTileIdCap) return FES_OK;
  FES_CUDA(P->tile_nodes.alloc((size_t)nt * fes_plan::kTileNodeStride + 64));
  FES_CUDA(P->tile_nnod.alloc(nt + 1));
  FES_CUDA(P->rec_local.alloc(P->n_geo + 16));
  FES_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));
  k_tile_nodes<<<(unsigned)nt, 256, 0, stream>>>(P->tile_node0.p, P->ne_ptr.p, P->ne_code.p, d_elem_nodes, N, P->n_corner,
                                                 P->tile_nodes.p, P->tile_nnod.p, P->rec_local.p, bad.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) {  // a tile touches > 255 distinct nodes (scattered node numbering): retry with smaller tiles
    *retry = true;
    return FES_OK;
  }
  // packed tile headers: {g0, g1, a0, n_act}, {c0, n_con, n_distinct_nodes, v0}
  {
    std::vector<int32_t> h_nnod(nt);
    FES_CUDA(cudaMemcpy(h_nnod.data(), P->tile_nnod.p, nt * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<int32_t> hdr((size_t)nt * 8);
    for (int64_t t = 0; t < nt; ++t) {
      int32_t* h = hdr.data() + t * 8;
      h[0] = h_ne[tiles[t]]; h[1] = h_ne[tiles[t + 1]]; h[2] = h_act[t]; h[3] = h_act[t + 1] - h_act[t];
      h[4] = h_con[t]; h[5] = h_con[t + 1] - h_con[t]; h[6] = h_nnod[t]; h[7] = tiles[t];
    }
    FES_CUDA(P->tile_hdr.alloc(hdr.size() + 16));
    FES_CUDA(cudaMemcpy(P->tile_hdr.p, hdr.data(), hdr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  P->tiled = true;
  return FES_OK;
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
  {
    // tiles shrink until every tile references <= 255 distinct nodes; with records*corners <= 255
    // that always holds, so a mesh with scattered node numbering is still tiled (smaller tiles)
    int gmax = tile_geo_budget(N);
    const int n_corner = (N == 6) ? 3 : N;
    for (;;) {
      bool retry = false;
      FES_TRY(build_tiles(P, d_elem_nodes, stream, gmax, &retry));
      if (!retry) break;
      if (gmax * n_corner <= 255) break;  // cannot happen: such tiles always fit
      gmax = std::max(255 / n_corner, gmax / 2);
    }
  }
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
                   p->contrib.bytes() + p->indptr.bytes() + p->indices.bytes() + p->ne_ptr.bytes() +
                   p->ne_code.bytes() + p->tile_node0.bytes() + p->tile_act0.bytes() + p->tile_con0.bytes() +
                   p->act_meta.bytes() + p->act_codes.bytes() + p->contrib16.bytes() + p->pair_lrow.bytes() +
                   p->tile_nodes.bytes() + p->tile_nnod.bytes() + p->rec_local.bytes());
}

// what one fused assembly streams from the plan (overhead on top of the algorithmic bytes)
extern "C" int64_t fes_plan_read_bytes(const fes_plan* p) {
  if (!p) return 0;
  if (p->tiled)
    return (int64_t)(p->ne_ptr.bytes() + p->tile_node0.bytes() + p->tile_act0.bytes() + p->tile_con0.bytes() +
                     p->act_meta.bytes() + p->act_codes.bytes() + p->rec_local.bytes() + p->tile_nnod.bytes() +
                     (int64_t)p->n_tiles * 4 * 128 /* distinct-node lists, ~128 used per tile */);
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
