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
  // A tile is a run of consecutive nodes whose (node, incident element) records fit in shared
  // memory.  Phase 1 of the kernel evaluates one geometry record per (node, element); phase 2
  // gathers the contributions of the tile's ACTIVE pairs -- pairs (v, w) with w >= v; the block of
  // (w, v) is the transpose and is written by the same thread.
  bool tiled = false;
  int64_t n_tiles = 0, n_geo = 0, n_active = 0, n_act_contrib = 0;
  int geo_max = 0;                     // records per tile upper bound
  int act_max = 0, con_max = 0;        // active pairs / their contributions per tile, upper bounds
  fes::DevBuf<int32_t> ne_ptr;         // n_nodes + 1 : incident-element ranges per node
  fes::DevBuf<uint32_t> ne_code;       // n_geo : e*8 + a, ascending e per node
  fes::DevBuf<int32_t> tile_node0;     // n_tiles + 1
  fes::DevBuf<int32_t> tile_act0;      // n_tiles + 1 : active-pair ranges
  fes::DevBuf<int32_t> tile_con0;      // n_tiles + 1 : ranges in act_codes
  fes::DevBuf<uint4> act_meta;         // n_active : per active pair, in descending contribution count per tile:
                                       //   x = off16 | cnt8<<16 | diag<<24, y = q0 own, z = q0 mirror,
                                       //   w = deg own | deg mirror<<16;  row i of a block starts at C*(q0 + i*deg)
  fes::DevBuf<uint16_t> act_codes;     // n_act_contrib : (record index in tile)<<6 | a<<3 | b, ascending element id
  // distinct nodes referenced by a tile's records (so coordinates are gathered once per tile) and
  // each record's corner nodes as indices into that list
  static constexpr int kTileNodeStride = 256;
  int n_corner = 0;                    // corner nodes stored per record (N, or 3 for the 6-node triangle)
  fes::DevBuf<int32_t> tile_nodes;     // n_tiles * 256 : sorted distinct node ids (first tile_nnod[t] valid)
  fes::DevBuf<int32_t> tile_nnod;      // n_tiles
  fes::DevBuf<int32_t> tile_hdr;       // n_tiles * 8 : {g0, g1, a0, n_act, c0, n_con, n_distinct, v0}
  fes::DevBuf<uint32_t> rec_local;     // n_geo : corner k's index in the tile list at bits [8k, 8k+8)
  // build-time only (kept for k_assemble_precomputed and the untiled fallback): see above
  fes::DevBuf<uint16_t> contrib16;     // n_contrib : tile-relative codes in pair order
  fes::DevBuf<uint8_t> pair_lrow;      // n_pairs : row node relative to the tile's first node

  // staging for the host-buffer entry point (fes_assemble_host), allocated on first use
  mutable fes::DevBuf<double> stage_coords, stage_values, stage_E, stage_scale;
};

namespace fes {
// records per tile for an element with N nodes (sized for the largest record, elasticity)
inline int tile_geo_budget(int N) {
  switch (N) {
    case 2: return 256;
    case 3: return 256;    // 3-node: P1 triangle (6 doubles -> 12 KB) or P2 line
    case 4: return 256;    // P1 tet (12 doubles -> 24 KB)
    default: return 128;   // P2 triangle (36 doubles -> 36 KB)
  }
}
}  // namespace fes
