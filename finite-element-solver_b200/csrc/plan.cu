// CSR sparsity pattern + scatter plan, built on the device once per (connectivity, n_comp).
//
// Replaces _ScatterPlan.build / FunctionSpace._scatter_plan (ref fem/space.py:107-131,757-772).
// The reference argsorts n_el*k^2 DOF-level int64 keys; here the sort runs on the c^2-times
// smaller NODE-level key set and the scalar pattern is expanded afterwards, which
// yields the identical indptr/indices:
//   1. key[e,a,b] = node[e,a] * n_nodes + node[e,b],  payload = e*N^2 + a*N + b
//   2. stable LSD radix sort by key (sortscan.cuh) -> equal keys keep ascending (e,a,b) order, i.e. the
//      order np.bincount adds an element's contribution to a slot
//   3. head flags + inclusive scan    -> pair ids, pair_start offsets, node adjacency
//   4. per-node degree scan           -> node_ptr;  expansion by n_comp -> indptr / indices
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "plan.cuh"
#include "sortscan.cuh"

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

// ---- per-tile build ---------------------------------------------------------------------------
constexpr int kTileIdCap = 2048;   // ids sorted per tile: records, then distinct elements x corner nodes
constexpr int kTileRecCap = 1024;  // distinct elements per tile (10-bit record index in a code)

// Block-wide (256 threads): sort ids[0..n) ascending (bitonic over the enclosing power of two) and
// write the distinct values to out[]; returns their count.  s_cnt holds 257 ints.
__device__ int sort_unique(uint32_t* ids, uint32_t* out, int n, int* s_cnt) {
  int cap = 2;
  while (cap < n) cap <<= 1;
  for (int i = n + threadIdx.x; i < cap; i += blockDim.x) ids[i] = 0xffffffffu;
  __syncthreads();
  for (int k = 2; k <= cap; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const uint32_t a = ids[i], b = ids[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { ids[i] = b; ids[l] = a; }
        }
      }
      __syncthreads();
    }
  const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  int c = 0;
  for (int i = lo; i < hi; ++i) c += (i == 0 || ids[i] != ids[i - 1]);
  s_cnt[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) { const int t = s_cnt[i]; s_cnt[i] = run; run += t; }
    s_cnt[blockDim.x] = run;
  }
  __syncthreads();
  int w = s_cnt[threadIdx.x];
  for (int i = lo; i < hi; ++i)
    if (i == 0 || ids[i] != ids[i - 1]) out[w++] = ids[i];
  __syncthreads();
  return s_cnt[blockDim.x];
}

// One CTA per tile: the distinct elements touching the tile's nodes (ascending id), the distinct
// corner nodes of those elements (ascending), and every element's corners as indices into that list.
__global__ void __launch_bounds__(256)
k_tile_build(const int32_t* __restrict__ tile_node0, const int32_t* __restrict__ ne_ptr,
             const uint32_t* __restrict__ ne_code, const int32_t* __restrict__ en, int N, int n_corner,
             uint32_t* __restrict__ tile_elems, uint32_t* __restrict__ elems_sorted, uint16_t* __restrict__ slot_of,
             uint32_t* __restrict__ rec_local, int32_t* __restrict__ tile_nodes,
             int32_t* __restrict__ tile_cnt, int colour_groups, int slot_mode, int* __restrict__ bad) {
  __shared__ uint32_t ids[kTileIdCap];
  __shared__ uint32_t uniq[kTileIdCap];
  __shared__ uint32_t s_el[kTileRecCap];
  __shared__ int s_cnt[257];
  __shared__ int s_mptr[257];             // node -> its (wavefront group, corner) memberships
  __shared__ uint16_t s_mlist[kTileIdCap];
  __shared__ int s_used[kTileIdCap * 2];   // [group * n_corner + corner][residue]: reads already on that bank
  __shared__ uint8_t s_li[256];
  __shared__ uint16_t s_slot[kTileRecCap];  // record (rank in ascending element id) -> position in the kernel's tables
  const int t = blockIdx.x;
  const int32_t g0 = ne_ptr[tile_node0[2 * t]], g1 = ne_ptr[tile_node0[2 * t + 1]];
  const int n_rec = g1 - g0;  // <= inc_max <= kTileRecCap
  for (int i = threadIdx.x; i < n_rec; i += blockDim.x) ids[i] = ne_code[g0 + i] >> 3;
  __syncthreads();
  const int ne = sort_unique(ids, uniq, n_rec, s_cnt);
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    s_el[i] = uniq[i];
    elems_sorted[g0 + i] = uniq[i];
  }
  __syncthreads();
  // Position of every record in the kernel's tables (= its shared-memory slot, up to the swizzle).
  // Modes 0 / 1: ascending element id.  Mode 2: order of first appearance when the tile's nodes are
  // swept incidence-major -- the k-th incident element of node v0, of v0+1, ..., then the (k+1)-th --
  // so on a regularly numbered mesh the k-th elements of consecutive nodes sit in consecutive slots
  // whatever the element numbering's stride is (3 and 6 records on Kuhn tetrahedra), and the
  // quarter-warp gathers of phase 2 fall into distinct banks.
  if (slot_mode == 2) {
    const int32_t v0 = tile_node0[2 * t], v1 = tile_node0[2 * t + 1], nn = v1 - v0;
    for (int i = threadIdx.x; i < ne; i += blockDim.x) ids[i] = 0xffffffffu;
    __syncthreads();
    for (int i = threadIdx.x; i < n_rec; i += blockDim.x) {
      int lo = v0, hi = v1;  // node of incidence g0 + i: the last v with ne_ptr[v] <= g0 + i
      while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (ne_ptr[m] <= g0 + i) lo = m; else hi = m;
      }
      const uint32_t pos = (uint32_t)(g0 + i - ne_ptr[lo]) * (uint32_t)nn + (uint32_t)(lo - v0);
      const uint32_t e = ne_code[g0 + i] >> 3;
      int a = 0, b = ne;
      while (a < b) {
        const int m = (a + b) >> 1;
        if (s_el[m] < e) a = m + 1; else b = m;
      }
      atomicMin(&ids[a], pos);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ne; i += blockDim.x) ids[i] = (ids[i] << 10) | (uint32_t)i;  // pos < 2^18, i < 2^10: unique
    __syncthreads();
    sort_unique(ids, uniq, ne, s_cnt);
    for (int i = threadIdx.x; i < ne; i += blockDim.x) s_slot[uniq[i] & 1023u] = (uint16_t)i;
  } else {
    for (int i = threadIdx.x; i < ne; i += blockDim.x) s_slot[i] = (uint16_t)i;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    tile_elems[g0 + s_slot[i]] = s_el[i];
    slot_of[g0 + i] = s_slot[i];
  }
  const int n_ids = ne * n_corner;
  if (n_ids > kTileIdCap) {  // uniform across the CTA
    if (threadIdx.x == 0) { atomicExch(bad, 1); tile_cnt[2 * t] = ne; tile_cnt[2 * t + 1] = 0; }
    return;
  }
  for (int i = threadIdx.x; i < n_ids; i += blockDim.x) {
    const int r = i / n_corner, k = i - r * n_corner;
    ids[i] = (uint32_t)en[(int64_t)s_el[r] * N + k];
  }
  __syncthreads();
  const int nd = sort_unique(ids, uniq, n_ids, s_cnt);
  if (threadIdx.x == 0) {
    tile_cnt[2 * t] = ne;
    tile_cnt[2 * t + 1] = nd;
    if (nd > 255) atomicExch(bad, 1);
  }
  if (nd > 255) return;
  // corner ranks (index into the sorted distinct-node list) of every record
  for (int i = threadIdx.x; i < n_ids; i += blockDim.x) {
    const int r = i / n_corner, k = i - r * n_corner;
    const uint32_t node = (uint32_t)en[(int64_t)s_el[r] * N + k];
    int lo = 0, hi = nd;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (uniq[m] < node) lo = m + 1; else hi = m;
    }
    ids[i] = (uint32_t)lo;
  }
  __syncthreads();
  // Local index of each distinct node = its slot in the kernel's shared-memory coordinate table.
  // Phase 1 of the assembly kernel reads corner k of G consecutive records per LSU wavefront (a
  // quarter-warp of 16-byte (x, y) pairs, or a half-warp of 8-byte components in 3D), and the
  // address decides the bank: slot mod G.  Sorted ranks collide (the nodes of neighbouring mesh rows
  // sit a row length apart), so the slots come from a greedy colouring instead: nodes in ascending
  // order, each takes the residue that adds the fewest same-bank reads to the wavefronts it is part
  // of (ties: the emptiest residue), slot = residue + G * (nodes already holding that residue).
  const int G = colour_groups;                 // 8 or 16 records per wavefront; 0 = keep sorted ranks
  int n_slots = nd;
  if (G > 0) {
    const int n_groups = (ne + G - 1) / G;     // <= 128
    const int cap = min((nd + G - 1) / G + 2, 256 / G);
    for (int i = threadIdx.x; i < n_groups * n_corner * G; i += blockDim.x) s_used[i] = 0;
    for (int i = threadIdx.x; i <= nd; i += blockDim.x) s_mptr[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_ids; i += blockDim.x) atomicAdd(&s_mptr[ids[i] + 1], 1);
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < nd; ++i) s_mptr[i + 1] += s_mptr[i];
    __syncthreads();
    for (int i = threadIdx.x; i < nd; i += blockDim.x) s_cnt[i] = s_mptr[i];  // fill cursors (nd <= 255)
    __syncthreads();
    for (int i = threadIdx.x; i < n_ids; i += blockDim.x) {
      const int r = i / n_corner, k = i - r * n_corner;
      const int at = atomicAdd(&s_cnt[ids[i]], 1);
      s_mlist[at] = (uint16_t)(((int)s_slot[r] / G) * n_corner + k);  // (wavefront group, corner): order within a node is irrelevant
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      int fill = 0, max_fill = 0;  // lane c < G counts the nodes holding residue c
      for (int n = 0; n < nd; ++n) {
        const int m0 = s_mptr[n], m1 = s_mptr[n + 1];
        uint32_t cost = 0xffffffffu;
        if (lane < G && fill < cap) {
          uint32_t hits = 0;
          for (int m = m0; m < m1; ++m) hits += s_used[(int)s_mlist[m] * G + lane];
          cost = (hits << 16) | ((uint32_t)fill << 8) | (uint32_t)lane;
        }
        for (int o = 16; o > 0; o >>= 1) cost = min(cost, __shfl_xor_sync(0xffffffffu, cost, o));
        const int best = (int)(cost & 255u);
        const int before = __shfl_sync(0xffffffffu, fill, best);
        if (lane == 0) s_li[n] = (uint8_t)(best + G * before);
        if (lane == best) ++fill;
        max_fill = max(max_fill, before + 1);
        for (int m = m0 + lane; m < m1; m += 32) atomicAdd(&s_used[(int)s_mlist[m] * G + best], 1);  // a node can be two corners of one group
        __syncwarp();
      }
      if (lane == 0) s_cnt[256] = G * max_fill;
    }
    __syncthreads();
    n_slots = s_cnt[256];
  } else {
    for (int i = threadIdx.x; i < nd; i += blockDim.x) s_li[i] = (uint8_t)i;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_cnt[2 * t + 1] = n_slots;
  int32_t* my_nodes = tile_nodes + (int64_t)t * fes_tileset::kTileNodeStride;
  for (int i = threadIdx.x; i < n_slots; i += blockDim.x) my_nodes[i] = (int32_t)uniq[0];  // unused slots: any valid node
  __syncthreads();
  for (int i = threadIdx.x; i < nd; i += blockDim.x) my_nodes[s_li[i]] = (int32_t)uniq[i];
  for (int r = threadIdx.x; r < ne; r += blockDim.x) {
    uint32_t packed = 0;
    for (int k = 0; k < n_corner; ++k) packed |= (uint32_t)s_li[ids[r * n_corner + k]] << (8 * k);
    rec_local[g0 + s_slot[r]] = packed;
  }
}

// Work items of a pair with `cnt` contributions when an item takes at most L of them: a power-of-two
// group of lanes (so a fixed xor-shuffle tree sums the partial blocks) with m contributions each.
__device__ __forceinline__ void item_split(uint32_t cnt, uint32_t L, uint32_t& lg, uint32_t& m) {
  const uint32_t need = (cnt + L - 1) / L;
  lg = 0;
  while ((1u << lg) < need && lg < 5) ++lg;
  m = (cnt + (1u << lg) - 1) >> lg;
}

// packed tile headers (12 ints): {g0, ne, p0, n_pairs}, {c0, n_con, n_distinct_nodes, v0},
// {i0 (filled in by the host scan), n_items, 0, 0}
__global__ void k_tile_hdr(const int32_t* __restrict__ tile_node0, int64_t n_tiles, const int32_t* __restrict__ ne_ptr,
                           const int32_t* __restrict__ node_ptr, const uint32_t* __restrict__ pair_start,
                           const int32_t* __restrict__ tile_cnt, uint32_t L, int32_t* __restrict__ hdr,
                           int* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int32_t v0 = tile_node0[2 * t], v1 = tile_node0[2 * t + 1];
  const int32_t p0 = node_ptr[v0], p1 = node_ptr[v1];
  const uint32_t c0 = pair_start[p0], c1 = pair_start[p1];
  int32_t items = 0;
  for (int32_t p = p0; p < p1; ++p) {
    const uint32_t cnt = pair_start[p + 1] - pair_start[p];
    uint32_t lg, m;
    item_split(cnt, L, lg, m);
    if (m > L) atomicExch(bad, 1);  // a pair shared by > 32 L elements: keep the untiled kernel
    items += 1 << lg;
  }
  int32_t* h = hdr + t * fes_tileset::kHdrInts;
  h[0] = ne_ptr[v0]; h[1] = tile_cnt[2 * t]; h[2] = p0; h[3] = p1 - p0;
  h[4] = (int32_t)c0; h[5] = (int32_t)(c1 - c0); h[6] = tile_cnt[2 * t + 1]; h[7] = v0;
  h[8] = 0; h[9] = items; h[10] = 0; h[11] = 0;
}

// Emit a tile's work items: 4 bytes of metadata and exactly L contribution codes each (short items
// are padded with the all-zero record, so the kernel's gather loop has no trip count at all).
// A pair with more than L contributions is split over a power-of-two group of adjacent lanes (a
// diagonal pair of a triangle mesh: 6-8 contributions over 4 lanes).  Order: descending group size
// (keeps groups aligned), descending trips, then node degree, slot-in-row, node -- so neighbouring threads
// handle the SAME relative neighbour of consecutive nodes with the same stencil.  On a regularly numbered mesh their
// record indices then form a constant-stride sequence, which the rec_slot swizzle spreads over all
// banks (conflict-free gathers); on an irregular mesh any order is equally random.
//   meta = q (16) | deg (12) | lg (3) | writer (1);  code = byte offset of g~_a | byte offset of g~_b << 16
__global__ void k_tile_emit(const int32_t* __restrict__ hdr, int64_t n_tiles, const int32_t* __restrict__ node_ptr,
                            const int32_t* __restrict__ pair_row, const int32_t* __restrict__ node_adj,
                            const uint32_t* __restrict__ pair_start, const uint32_t* __restrict__ contrib,
                            const uint32_t* __restrict__ elems_sorted, const uint16_t* __restrict__ slot_of, int N, int c, int gs,
                            int slot_mode, int by_degree, uint32_t L,
                            uint32_t* __restrict__ meta, uint32_t* __restrict__ codes, int* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int32_t* h = hdr + t * fes_tileset::kHdrInts;
  const int32_t g0 = h[0], ne = h[1], p0 = h[2], p1 = h[2] + h[3], v0 = h[7];
  const int N2 = N * N;
  const uint32_t zero_code = ((uint32_t)(gs - 1) << 4) * 0x10001u;
  uint32_t hi = 0;  // class key: lg << 8 | m
  int32_t max_deg = 0, v1 = v0;
  for (int32_t p = p0; p < p1; ++p) {
    uint32_t lg, m;
    item_split(pair_start[p + 1] - pair_start[p], L, lg, m);
    hi = max(hi, (lg << 8) | m);
  }
  if (p1 > p0) v1 = pair_row[p1 - 1] + 1;
  for (int32_t v = v0; v < v1; ++v) max_deg = max(max_deg, node_ptr[v + 1] - node_ptr[v]);
  if (max_deg > 4095 || h[9] > 65535) { atomicExch(bad, 1); return; }
  int64_t wa = h[8];
  // by_degree: within a class, nodes of equal degree go together (ascending degree), so neighbouring
  // lanes hold the same relative neighbour of nodes with the SAME stencil (a criss-cross triangulation
  // alternates 4- and 8-valent nodes).  Degrees above 63 share one bucket.
  uint64_t deg_mask = by_degree ? 0ull : 1ull;
  if (by_degree)
    for (int32_t v = v0; v < v1; ++v) deg_mask |= 1ull << min(63, node_ptr[v + 1] - node_ptr[v]);
  while (hi > 0) {
    uint32_t next = 0;
    for (uint64_t dm = deg_mask; dm; dm &= dm - 1) {
    const int32_t dsel = __ffsll((long long)dm) - 1;
    // by_degree == 2: neighbouring lanes alternate between slots sl and sl + 2 of the same node, then move to the
    // next node.  Same-degree nodes are a constant number of 16-byte units apart in the value image (4 mod 8 on a
    // criss-cross triangulation), so lanes holding ONE slot of consecutive nodes hit two bank groups; the pair
    // (sl, sl + 2) doubles that to four and the row swap to all eight.  Bank model: image stores 1.77x -> 1.25x of
    // conflict-free, gathers 1.19x -> 1.41x; measured neutral (config 2: 0.1627 vs 0.1617 ms), so a knob only.
    const int32_t n_outer = (by_degree == 2) ? 2 * ((max_deg + 3) / 4) : max_deg;  // (parity, block of 4) or slot
    const int32_t n_inner = (by_degree == 2) ? 2 : 1;
    for (int32_t so = 0; so < n_outer; ++so)
      for (int32_t v = v0; v < v1; ++v)
      for (int32_t si = 0; si < n_inner; ++si) {
        const int32_t sl = (by_degree == 2) ? 4 * (so >> 1) + 2 * si + (so & 1) : so;
        const int32_t npv = node_ptr[v], degv = node_ptr[v + 1] - npv;
        if (sl >= degv) continue;
        if (by_degree && min(63, degv) != dsel) continue;
        const int32_t p = npv + sl;
        const uint32_t k0 = pair_start[p], cnt = pair_start[p + 1] - k0;
        uint32_t lg, m;
        item_split(cnt, L, lg, m);
        const uint32_t key = (lg << 8) | m;
        if (key == hi) {
          const uint32_t q = (uint32_t)c * (uint32_t)(npv - p0) + (uint32_t)sl;
          if (q > 65535u) atomicExch(bad, 1);
          for (uint32_t j = 0; j < (1u << lg); ++j, ++wa) {
            meta[wa] = (q & 0xffffu) | ((uint32_t)degv << 16) | (lg << 28) | ((j == 0 ? 1u : 0u) << 31);
            for (uint32_t x = 0; x < L; ++x) {
              const uint32_t k = j * m + x;
              uint32_t out = zero_code;
              if (x < m && k < cnt) {
                const uint32_t code = contrib[k0 + k], e = code / N2, ab = code - e * N2, a = ab / N, b = ab - a * N;
                int lo = 0, hi_ = ne;  // record of element e: its rank among the tile's distinct elements
                while (lo < hi_) {
                  const int mid = (lo + hi_) >> 1;
                  if (elems_sorted[g0 + mid] < e) lo = mid + 1; else hi_ = mid;
                }
                const uint32_t slot = rec_slot((uint32_t)slot_of[g0 + lo], slot_mode);
                out = ((a * (uint32_t)gs + slot) << 4) | ((b * (uint32_t)gs + slot) << 20);
              }
              codes[wa * L + x] = out;
            }
          }
        } else if (key < hi) {
          next = max(next, key);
        }
      }
    }
    hi = next;
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

// Incident-element lists per node (ne_ptr / ne_code), shared by the tilings and the stencil-run detection.
// Leaves P->n_geo = 0 when the mesh has degenerate elements (repeated nodes): nothing is tiled then.
int build_incidence(fes_plan* P, cudaStream_t stream, std::vector<int32_t>& h_ne, std::vector<int32_t>& h_np) {
  const int64_t nn = P->n_nodes, np_ = P->n_pairs;
  const int N = P->N;
  P->n_geo = 0;
  if (np_ == 0) return FES_OK;
  DevBuf<int32_t> ne_cnt;
  FES_CUDA(ne_cnt.alloc(nn + 1));
  FES_CUDA(cudaMemsetAsync(ne_cnt.p, 0, ne_cnt.bytes(), stream));
  k_diag_counts<<<grid_for(np_, kBlock), kBlock, 0, stream>>>(P->pair_row.p, P->node_adj.p, P->pair_start.p, np_, ne_cnt.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(P->ne_ptr.alloc(nn + 1));
  FES_TRY((device_scan<int32_t, false>(ne_cnt.p, P->ne_ptr.p, nn + 1, stream)));
  h_ne.resize(nn + 1); h_np.resize(nn + 1);
  FES_CUDA(cudaMemcpyAsync(h_ne.data(), P->ne_ptr.p, (nn + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaMemcpyAsync(h_np.data(), P->node_ptr.p, (nn + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  if (h_ne[nn] != P->n_el * N) return FES_OK;  // degenerate elements (repeated nodes): keep the untiled kernel
  if (P->n_el >= ((int64_t)1 << 29)) return FES_OK;
  P->n_geo = h_ne[nn];
  FES_CUDA(P->ne_code.alloc(P->n_geo));
  k_fill_ne<<<grid_for(np_, kBlock), kBlock, 0, stream>>>(P->pair_row.p, P->node_adj.p, P->pair_start.p, P->contrib.p, np_,
                                                          N, P->ne_ptr.p, P->ne_code.p);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

// Tile the node ranges `ranges` = {begin0, end0, begin1, end1, ...} (ascending, disjoint) for the fused assembly
// kernel.  The greedy split runs on the host over the two prefix arrays (O(n_nodes), once per mesh); everything else
// stays on the device.  Tiles never straddle a range boundary.
int build_tiles(fes_plan* P, fes_tileset& T, const std::vector<int32_t>& ranges, const std::vector<int32_t>& h_ne,
                const std::vector<int32_t>& h_np, const int32_t* d_elem_nodes, cudaStream_t stream, int inc_max,
                bool* retry) {
  *retry = false;
  const int64_t np_ = P->n_pairs;
  const int N = P->N, c = P->c;
  T.tiled = false;
  T.n_tiles = 0;
  if (np_ == 0 || P->n_geo == 0 || ranges.empty()) return FES_OK;
  T.inc_max = inc_max;
  // a tile's value image stays under 32 KB of shared memory, its packed offsets under 16 bits
  const int pair_cap = std::min(32768 / (8 * c * c), 65535 / c);
  // Boundary nodes carry more pairs per incident element than interior ones, so a run of them within
  // the incidence budget would hold far more pairs than a typical tile -- and the kernel sizes its
  // value image (shared memory, hence CTAs per SM) for the LARGEST tile.  Keep every tile within
  // 1.25x the mesh-average pairs of a full tile (a single node always fits).
  const int pair_soft = (int)(1.25 * (double)np_ / (double)P->n_geo * inc_max) + 8;
  std::vector<int32_t> tiles;
  // Runs whose length is a multiple of `node_mult` keep the lane groups of phase 2 (same relative neighbour of
  // consecutive nodes) aligned to quarter- and half-warps.  Tets: 16-node runs gather at 1.18x / 1.31x of their
  // conflict-free wavefronts, 17-node runs at 1.43x / 1.79x (tools/bank_sim.py model).
  const int64_t node_mult = getenv("FES_TILE_NODE_MULT") ? atoi(getenv("FES_TILE_NODE_MULT")) : tile_node_multiple(N);
  for (size_t r = 0; r + 1 < ranges.size(); r += 2) {
    int64_t v0 = ranges[r];
    const int64_t vend = ranges[r + 1];
    for (int64_t v = v0; v < vend; ++v) {
      if (h_ne[v + 1] - h_ne[v] > inc_max || h_np[v + 1] - h_np[v] > pair_cap) return FES_OK;  // one node overflows a tile
      if (h_ne[v + 1] - h_ne[v0] > inc_max || h_np[v + 1] - h_np[v0] > pair_cap || v - v0 >= 255 ||
          (h_np[v + 1] - h_np[v0] > pair_soft && v > v0)) {
        int64_t cut = v;
        if (node_mult > 1 && v - v0 > node_mult) cut = v0 + (v - v0) / node_mult * node_mult;
        tiles.push_back((int32_t)v0); tiles.push_back((int32_t)cut);
        v0 = cut;
        v = cut - 1;  // re-scan from the cut
      }
    }
    if (vend > v0) { tiles.push_back((int32_t)v0); tiles.push_back((int32_t)vend); }
  }
  T.n_tiles = (int64_t)tiles.size() / 2;
  const int64_t nt = T.n_tiles;
  if (nt == 0) return FES_OK;
  // per-tile tables live at the tile's offset into the global incidence list: sized to the last tile's end
  const int64_t g_end = h_ne[tiles.back()];
  FES_CUDA(T.tile_node0.alloc(tiles.size()));
  FES_CUDA(cudaMemcpyAsync(T.tile_node0.p, tiles.data(), tiles.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));

  // distinct elements and distinct corner nodes per tile
  T.n_corner = (N == 6) ? 3 : N;
  DevBuf<int> bad;
  FES_CUDA(bad.alloc(1));
  FES_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));
  FES_CUDA(T.tile_elems.alloc(g_end + 16));
  FES_CUDA(T.elems_sorted.alloc(g_end + 16));
  FES_CUDA(T.slot_of.alloc(g_end + 16));
  T.slot_mode = getenv("FES_SLOT_MODE") ? atoi(getenv("FES_SLOT_MODE")) : tile_slot_mode(N);  // env: tuning knob
  FES_CUDA(T.rec_local.alloc(g_end + 16));
  FES_CUDA(T.tile_nodes.alloc((size_t)nt * fes_tileset::kTileNodeStride + 64));
  FES_CUDA(T.tile_cnt.alloc(2 * nt));
  // coordinate-table slots coloured for 8 records per wavefront (16-byte xy pairs) or 16 (tets: 8-byte components)
  static const bool no_colour = getenv("FES_PLAN_NO_COLOUR") != nullptr;  // tuning knob: sorted ranks
  const int colour_groups = no_colour ? 0 : (N == 4 ? 16 : 8);
  k_tile_build<<<(unsigned)nt, 256, 0, stream>>>(T.tile_node0.p, P->ne_ptr.p, P->ne_code.p, d_elem_nodes, N, T.n_corner,
                                                 T.tile_elems.p, T.elems_sorted.p, T.slot_of.p, T.rec_local.p,
                                                 T.tile_nodes.p, T.tile_cnt.p, colour_groups, T.slot_mode, bad.p);
  FES_LAUNCH_CHECK();
  int h_bad = 0;
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) {  // a tile touches > 255 distinct nodes (scattered node numbering): retry with smaller tiles
    if (getenv("FES_PLAN_DEBUG")) {
      std::vector<int32_t> cnts(2 * nt);
      cudaMemcpy(cnts.data(), T.tile_cnt.p, cnts.size() * sizeof(int32_t), cudaMemcpyDeviceToHost);
      int mr = 0, mn = 0;
      for (int64_t t = 0; t < nt; ++t) { mr = std::max(mr, cnts[2 * t]); mn = std::max(mn, cnts[2 * t + 1]); }
      fprintf(stderr, "[fes plan] retry: inc_max=%d tiles=%lld max distinct elements=%d max distinct nodes=%d\n", inc_max,
              (long long)nt, mr, mn);
    }
    *retry = true;
    return FES_OK;
  }
  // an item gathers at most L contributions.  L depends on the element type only, never on the mesh,
  // so the summation tree of a pair is a function of its contribution count alone: a rank that
  // assembles a slab with ghost elements produces the same bits as the single-GPU assembly.
  const uint32_t L = (uint32_t)tile_item_len(N);
  constexpr int HI = fes_tileset::kHdrInts;
  FES_CUDA(T.tile_hdr.alloc((size_t)nt * HI + 16));
  FES_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));
  k_tile_hdr<<<grid_for(nt, kBlock), kBlock, 0, stream>>>(T.tile_node0.p, nt, P->ne_ptr.p, P->node_ptr.p, P->pair_start.p,
                                                          T.tile_cnt.p, L, T.tile_hdr.p, bad.p);
  FES_LAUNCH_CHECK();
  std::vector<int32_t> hdr((size_t)nt * HI);
  FES_CUDA(cudaMemcpyAsync(hdr.data(), T.tile_hdr.p, hdr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) return FES_OK;
  int rmax = 0, pmax = 0, cmax = 0, imax = 0, nmax = 0;
  int64_t rd = 0, items = 0;
  for (int64_t t = 0; t < nt; ++t) {
    int32_t* h = hdr.data() + t * HI;
    rmax = std::max(rmax, h[1]); pmax = std::max(pmax, h[3]); cmax = std::max(cmax, h[5]); imax = std::max(imax, h[9]); nmax = std::max(nmax, h[6]);
    h[8] = (int32_t)items;
    items += h[9];
    rd += 4 * HI + (int64_t)h[9] * 4 * (1 + L) + (int64_t)h[1] * 4 + (int64_t)h[6] * 4;
  }
  if (imax > 65535 || rmax > kTileRecCap || items * L >= ((int64_t)1 << 32)) return FES_OK;
  T.rec_max = rmax; T.pair_max = pmax; T.con_max = cmax; T.item_max = imax; T.n_items = items;
  T.nod_max = (nmax + 3) & ~3;
  // slots are swizzled inside aligned groups of 8 (mode 0)
  {
    // Smallest stride above every (swizzled) slot with the wanted residue; the last slot holds the zero record.
    // Stride = 0 (mod 8): the bank of a gathered gradient depends on its slot only, not on the local node
    // index, and the slot sequences of neighbouring lanes are what the item order and the slot modes make
    // conflict-free (bank model: gathers 1.34x -> 1.19x of conflict-free on config 2; measured 0.165 -> 0.162 ms).
    // The earlier 1 (mod 8) stride protected lanes reading different nodes of ONE record, which the
    // (class, degree, slot, node) item order no longer produces.
    static const int gs_align = getenv("FES_GS_ALIGN") ? atoi(getenv("FES_GS_ALIGN")) : 8;   // tuning knobs
    static const int gs_res = getenv("FES_GS_RES") ? atoi(getenv("FES_GS_RES")) : 0;
    int gs = ((rmax + 7) & ~7) + 1;
    while (gs % gs_align != gs_res % gs_align) ++gs;
    T.geo_stride = gs;
  }
  if ((int64_t)N * T.geo_stride * 16 > 65535) return FES_OK;  // 16-bit byte offsets in the codes
  T.tile_read_bytes = rd;
  if (getenv("FES_PLAN_DEBUG"))
    fprintf(stderr, "[fes plan] N=%d c=%d nodes=%lld pairs=%lld tiles=%lld inc_max=%d rec_max=%d pair_max=%d item_max=%d con_max=%d nod_max=%d items=%lld gs=%d\n",
            N, c, (long long)P->n_nodes, (long long)np_, (long long)nt, inc_max, rmax, pmax, imax, cmax, T.nod_max, (long long)items, T.geo_stride);
  FES_CUDA(cudaMemcpyAsync(T.tile_hdr.p, hdr.data(), hdr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  T.item_len = (int)L;
  FES_CUDA(T.pair_meta.alloc(items + 16));
  FES_CUDA(T.pair_codes.alloc(items * L + 32));
  k_tile_emit<<<grid_for(nt, 64), 64, 0, stream>>>(T.tile_hdr.p, nt, P->node_ptr.p, P->pair_row.p, P->node_adj.p,
                                                   P->pair_start.p, P->contrib.p, T.elems_sorted.p, T.slot_of.p, N, c, T.geo_stride, T.slot_mode,
                                                   getenv("FES_ITEM_BY_DEGREE") ? atoi(getenv("FES_ITEM_BY_DEGREE")) : 1, L,
                                                   T.pair_meta.p, T.pair_codes.p, bad.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  FES_CUDA(cudaStreamSynchronize(stream));  // also keeps `hdr` alive until the upload is done
  if (h_bad) return FES_OK;  // offsets beyond 16 bits: keep the untiled kernel
  T.elems_sorted.release();  // plan-time lookups only
  T.slot_of.release();
  T.tiled = true;
  return FES_OK;
}

// tiles shrink until every tile references <= 255 distinct nodes; with elements*corners <= 255
// that always holds, so a mesh with scattered node numbering is still tiled (smaller tiles)
int tile_with_retries(fes_plan* P, fes_tileset& T, const std::vector<int32_t>& ranges, const std::vector<int32_t>& h_ne,
                      const std::vector<int32_t>& h_np, const int32_t* d_elem_nodes, cudaStream_t stream) {
  const int N = P->N;
  int inc = tile_inc_budget(N);
  if (const char* env = getenv("FES_TILE_INC")) inc = std::max(16, std::min(kTileRecCap, atoi(env)));  // tuning knob
  const int n_corner = (N == 6) ? 3 : N;
  for (;;) {
    bool retry = false;
    FES_TRY(build_tiles(P, T, ranges, h_ne, h_np, d_elem_nodes, stream, inc, &retry));
    if (!retry) break;
    if (inc * n_corner <= 255) break;  // cannot happen: such tiles always fit
    inc = std::max(255 / n_corner, inc - std::max(1, inc / 8));  // shrink gently: usually a few tiles overflow
  }
  return FES_OK;
}

}  // namespace

// The tiling of every node, on first use by a form the run kernel does not cover (see fes_plan_create).
int ensure_full_tiling(fes_plan* P, const int32_t* d_elem_nodes, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(P->full_mutex);
  if (P->full_built) return FES_OK;
  FES_TRY(tile_with_retries(P, P->full, {0, (int32_t)P->n_nodes}, P->h_ne, P->h_np, d_elem_nodes, stream));
  P->full_built = true;
  return FES_OK;
}

int build_runs(fes_plan* P, const int32_t* d_elem_nodes, const std::vector<int32_t>& h_ne, const std::vector<int32_t>& h_np,
               cudaStream_t stream, std::vector<int32_t>& rest_ranges);  // plan_runs.cu
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
  bool in_b = false;
  FES_TRY(radix_sort_pairs<uint64_t>(keys_a.p, keys_b.p, vals_a.p, P->contrib.p, M, end_bit, stream, &in_b));
  if (!in_b) {  // an even number of passes left the result in the a buffers
    FES_CUDA(cudaMemcpyAsync(keys_b.p, keys_a.p, keys_a.bytes(), cudaMemcpyDeviceToDevice, stream));
    FES_CUDA(cudaMemcpyAsync(P->contrib.p, vals_a.p, vals_a.bytes(), cudaMemcpyDeviceToDevice, stream));
    FES_CUDA(cudaStreamSynchronize(stream));
  }
  keys_a.release(); vals_a.release();

  FES_CUDA(scan.alloc(M));
  k_head_flags<<<grid_for(M, kBlock), kBlock, 0, stream>>>(keys_b.p, M, scan.p);
  FES_LAUNCH_CHECK();
  FES_TRY((device_scan<uint32_t, true>(scan.p, scan.p, M, stream)));
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

  FES_TRY((device_scan<int32_t, false>(degree.p, P->node_ptr.p, n_nodes + 1, stream)));

  FES_CUDA(P->indices.alloc(P->nnz));
  k_expand_indptr<<<grid_for(n_nodes + 1, kBlock), kBlock, 0, stream>>>(P->node_ptr.p, n_nodes, n_comp, P->indptr.p);
  FES_LAUNCH_CHECK();
  k_expand_indices<<<grid_for(n_pairs, kBlock), kBlock, 0, stream>>>(P->node_ptr.p, P->node_adj.p, P->pair_row.p,
                                                                     n_pairs, n_comp, P->indices.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaStreamSynchronize(stream));
  {
    std::vector<int32_t>& h_ne = P->h_ne;
    std::vector<int32_t>& h_np = P->h_np;
    FES_TRY(build_incidence(P, stream, h_ne, h_np));
    // Stencil runs first (plan_runs.cu) + the tiling of whatever they leave uncovered.  When they take the mesh, the
    // tiling of EVERY node -- needed only by the forms the run kernel does not cover (mass, P2) -- is built on its
    // first use (ensure_full_tiling): a structured mesh pays neither its build time nor its 0.3 GB of index streams
    // unless it asks for such a form.
    if (P->n_geo > 0 && !getenv("FES_NO_RUNS")) {
      std::vector<int32_t> rest;
      FES_TRY(build_runs(P, d_elem_nodes, h_ne, h_np, stream, rest));
      if (P->runs.active) {
        FES_TRY(tile_with_retries(P, P->rest, rest, h_ne, h_np, d_elem_nodes, stream));
        if (!rest.empty() && !P->rest.tiled) P->runs.active = false;  // cannot tile the remainder: keep the full tiling only
      }
    }
    if (!P->runs.active) {
      FES_TRY(tile_with_retries(P, P->full, {0, (int32_t)n_nodes}, h_ne, h_np, d_elem_nodes, stream));
      P->full_built = true;
      std::vector<int32_t>().swap(P->h_ne);
      std::vector<int32_t>().swap(P->h_np);
    }
  }
  guard.p = nullptr;
  *out = P;
  return FES_OK;
}

// The builder's two primitives, exposed so the tests can check them against numpy directly
// (np.argsort(kind='stable') / np.cumsum) at awkward sizes and key distributions.
extern "C" int fes_sort_pairs_u64(uint64_t* d_keys, uint32_t* d_vals, int64_t n, int end_bit, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n < 0 || n >= ((int64_t)1 << 32) || end_bit < 1 || end_bit > 64)
    return set_error(FES_ERR_VALUE, "fes_sort_pairs_u64: bad arguments n=%lld end_bit=%d", (long long)n, end_bit);
  if (n == 0) return FES_OK;
  DevBuf<uint64_t> kb;
  DevBuf<uint32_t> vb;
  FES_CUDA(kb.alloc(n));
  FES_CUDA(vb.alloc(n));
  bool in_b = false;
  FES_TRY(radix_sort_pairs<uint64_t>(d_keys, kb.p, d_vals, vb.p, n, end_bit, st, &in_b));
  if (in_b) {
    FES_CUDA(cudaMemcpyAsync(d_keys, kb.p, kb.bytes(), cudaMemcpyDeviceToDevice, st));
    FES_CUDA(cudaMemcpyAsync(d_vals, vb.p, vb.bytes(), cudaMemcpyDeviceToDevice, st));
  }
  FES_CUDA(cudaStreamSynchronize(st));
  return FES_OK;
}

extern "C" int fes_exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, void* stream_) {
  if (n < 0) return set_error(FES_ERR_VALUE, "fes_exclusive_scan_i32: n < 0");
  return device_scan<int32_t, false>(d_in, d_out, n, (cudaStream_t)stream_);
}

extern "C" void fes_plan_destroy(fes_plan* plan) { delete plan; }
extern "C" int64_t fes_plan_n_dofs(const fes_plan* p) { return p ? p->n_dofs : 0; }
extern "C" int64_t fes_plan_nnz(const fes_plan* p) { return p ? p->nnz : 0; }
extern "C" int64_t fes_plan_n_pairs(const fes_plan* p) { return p ? p->n_pairs : 0; }
extern "C" int64_t fes_plan_n_tiles(const fes_plan* p) {
  if (!p) return 0;
  if (p->runs.active) return p->runs.n_tiles + p->rest.n_tiles;
  return p->full.tiled ? p->full.n_tiles : 0;
}
extern "C" int64_t fes_plan_n_run_tiles(const fes_plan* p) { return (p && p->runs.active) ? p->runs.n_tiles : 0; }
extern "C" int64_t fes_plan_run_nodes(const fes_plan* p) { return (p && p->runs.active) ? p->runs.n_nodes_covered : 0; }
extern "C" const int32_t* fes_plan_indptr(const fes_plan* p) { return p ? p->indptr.p : nullptr; }
extern "C" const int32_t* fes_plan_indices(const fes_plan* p) { return p ? p->indices.p : nullptr; }

extern "C" int64_t fes_plan_bytes(const fes_plan* p) {
  if (!p) return 0;
  return (int64_t)(p->node_ptr.bytes() + p->node_adj.bytes() + p->pair_row.bytes() + p->pair_start.bytes() +
                   p->contrib.bytes() + p->indptr.bytes() + p->indices.bytes() + p->ne_ptr.bytes() +
                   p->ne_code.bytes() + p->full.bytes() + p->rest.bytes() + p->runs.bytes());
}

// what one fused assembly streams from the plan (overhead on top of the algorithmic bytes): the stencil-run
// path reads 32 bytes per tile plus the index streams of the remainder tiling
extern "C" int64_t fes_plan_read_bytes(const fes_plan* p) {
  if (!p) return 0;
  if (p->runs.active) return (int64_t)p->runs.n_tiles * fes_runs::kTileInts * 4 + p->rest.tile_read_bytes;
  if (p->full.tiled) return p->full.tile_read_bytes;
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
