// The CSR pattern + scatter plan handle shared between plan.cu (builder) and assemble.cu.
#pragma once
#include "common.cuh"

struct fes_plan {
  int64_t n_el = 0, n_nodes = 0, n_dofs = 0, n_pairs = 0, nnz = 0, n_contrib = 0;
  int N = 0, c = 0;
  // node-level (block) CSR: row v holds its neighbour nodes in ascending order
  fes::DevBuf<int32_t> node_ptr;    // n_nodes + 1
  fes::DevBuf<int32_t> node_adj;    // n_pairs
  fes::DevBuf<int32_t> pair_row;    // n_pairs : row node of each pair
  // contributions: for pair p, contrib[pair_start[p] .. pair_start[p+1]) = e*N*N + a*N + b in
  // ascending element id -- the order np.bincount sums them in
  fes::DevBuf<uint32_t> pair_start;  // n_pairs + 1
  fes::DevBuf<uint32_t> contrib;     // n_el * N * N
  // scalar CSR, bit-identical to the reference's _ScatterPlan.indptr / .indices
  fes::DevBuf<int32_t> indptr;   // n_dofs + 1
  fes::DevBuf<int32_t> indices;  // nnz

  // ---- tiled layout for the fused assembly kernel ------------------------------------------
  // A tile is a run of consecutive nodes = a contiguous range of stored pairs = a contiguous range
  // of CSR values.  Phase 1 of the kernel evaluates one geometry record per DISTINCT element
  // touching the tile; phase 2 gathers the contributions of every pair of the tile into a
  // shared-memory image of the tile's values, which one TMA bulk store writes out.
  bool tiled = false;
  int64_t n_tiles = 0, n_geo = 0, tile_read_bytes = 0;
  int inc_max = 0;                     // (node, incident element) records per tile: the tiling budget
  int geo_stride = 0;                  // shared-memory plane stride of the records: 0 (mod 8), > every slot
  int slot_mode = 0;                   // rec_slot mode the codes were emitted with
  int rec_max = 0, pair_max = 0, con_max = 0, item_max = 0, nod_max = 0;  // distinct elements / pairs / contributions / work items / distinct nodes per tile, upper bounds
  int64_t n_items = 0;
  int item_len = 0;                    // contribution codes per work item (= tile_item_len(N))
  fes::DevBuf<int32_t> ne_ptr;         // n_nodes + 1 : incident-element ranges per node
  fes::DevBuf<uint32_t> ne_code;       // n_geo : e*8 + a, ascending e per node
  fes::DevBuf<int32_t> tile_node0;     // n_tiles + 1
  // distinct elements of a tile (ascending id) at [g0, g0 + ne), g0 = ne_ptr[first node]; each
  // element's corner nodes as indices into the tile's sorted distinct-node list
  static constexpr int kTileNodeStride = 256;
  int n_corner = 0;                    // corner nodes stored per record (N, or 3 for the 6-node triangle)
  fes::DevBuf<uint32_t> tile_elems;    // n_geo (+pad) : element ids in table order (see rec_slot)
  fes::DevBuf<uint32_t> elems_sorted;  // n_geo (+pad) : the same ids ascending   } plan-time lookups, released once the
  fes::DevBuf<uint16_t> slot_of;       // n_geo (+pad) : table position by rank   } work items are emitted
  fes::DevBuf<uint32_t> rec_local;     // n_geo (+pad) : corner k's index in the tile list at bits [8k, 8k+8)
  fes::DevBuf<int32_t> tile_nodes;     // n_tiles * 256 : sorted distinct node ids (first nd valid)
  fes::DevBuf<int32_t> tile_cnt;       // n_tiles * 2 : {distinct elements, distinct nodes}
  static constexpr int kHdrInts = 12;
  fes::DevBuf<int32_t> tile_hdr;       // n_tiles * 12 : {g0, ne, p0, n_pairs}, {c0, n_con, n_distinct_nodes, v0}, {i0, n_items, 0, 0}
  // work items of a tile (a pair, or one of the 2^lg lanes that share a long pair), ordered by
  // descending group size, descending trips, slot-in-row, node:
  //   meta = q (16) | deg (12) | lg (3) | writer (1);  row i of the pair's C x C block starts at double
  //   C*(q + i*deg) of the tile's value image.  Item t owns codes [t*item_len, (t+1)*item_len).
  fes::DevBuf<uint32_t> pair_meta;     // n_items (+pad)
  fes::DevBuf<uint32_t> pair_codes;    // n_items * item_len (+pad) : 16*(a*geo_stride + slot) | 16*(b*geo_stride + slot) << 16
                                       //   (byte offsets), slot = rec_slot(record index in tile); short items padded with
                                       //   the zero record (slot geo_stride-1); ascending element id within an item

  // staging for the host-buffer entry point (fes_assemble_host), allocated on first use
  mutable fes::DevBuf<double> stage_coords, stage_values, stage_E, stage_scale;
};

namespace fes {
// Shared-memory slot of the record at position r of a tile's element tables.
//   mode 0: tables in ascending element id, slot = an XOR swizzle of the low three bits, so strided
//           record sequences (the k-th element of consecutive nodes: stride 2 or 4 records on triangle
//           meshes) spread over the banks;
//   mode 1: ascending element id, slot = position;
//   mode 2: tables in incidence-major first-appearance order (k_tile_build), slot = position -- the
//           k-th elements of consecutive nodes sit in consecutive slots whatever the stride of the
//           element numbering.  Tetrahedral meshes (strides 3 and 6 records): gathers at 2.36x / 3.05x
//           of their conflict-free wavefronts in mode 0, 2.14x / 2.27x in mode 1, 1.31x / 1.56x in
//           mode 2 (tools/bank_sim.py model; the first two confirmed by ncu).
// Baked into the contribution codes at plan time; phase 1 of the kernel applies it when it writes.
__host__ __device__ inline uint32_t rec_slot(uint32_t r, int mode) { return mode ? r : r ^ ((r >> 3) & 7u); }
inline int tile_slot_mode(int N) { return N == 4 ? 2 : 0; }
// preferred run length multiple of a tile (see build_tiles); measured: no effect on the B200 (tets 0.5695 ms with 16,
// 0.5699 without -- the tet kernel is bound by fp64 / LDS latency at 12 warps per SM, not by the conflicts), so off
inline int tile_node_multiple(int N) { (void)N; return 1; }

// contributions one lane gathers at most (longer pairs are split over 2^lg lanes): about the count of
// an interior off-diagonal pair of that element type
inline int tile_item_len(int N) {
  switch (N) {
    case 3: return 2;   // triangle edge: 2 elements (P2 line: 1-2)
    case 4: return 6;   // tet edge: 4-6 elements
    default: return 2;  // lines, P2 triangle (most pairs: 1-2)
  }
}

// (node, incident element) records per tile for an element with N nodes: bounds the distinct
// elements a tile can reference (sized for the largest record, elasticity)
inline int tile_inc_budget(int N) {
  switch (N) {
    case 2: return 256;
    case 3: return 336;    // P1 triangle (6 doubles per element record) or P2 line: 5 CTAs per SM (measured; 384 -> 4 CTAs)
    case 4: return 384;    // P1 tet (12 doubles per element record): 16 interior nodes per tile, 3 CTAs per SM (measured: 400 +1 %, 416 -> 2 CTAs +30 %)
    default: return 192;   // P2 triangle (row records: 4 doubles per element); measured 96..256, flat above 128
  }
}
}  // namespace fes
