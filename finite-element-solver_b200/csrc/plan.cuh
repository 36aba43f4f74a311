// The CSR pattern + scatter plan handle shared between plan.cu (builder) and assemble.cu.
#pragma once
#include <mutex>
#include <vector>

#include "common.cuh"

// One tiling of (a subset of) the nodes for the fused assembly kernel k_assemble_tiles.
// A tile is a run of consecutive nodes = a contiguous range of stored pairs = a contiguous range
// of CSR values.  Phase 1 of the kernel evaluates one geometry record per DISTINCT element
// touching the tile; phase 2 gathers the contributions of every pair of the tile into a
// shared-memory image of the tile's values, which one TMA bulk store writes out.
// A plan holds two: `full` covers every node; `rest` covers the nodes no stencil run covers (see fes_runs).
struct fes_tileset {
  bool tiled = false;
  int64_t n_tiles = 0, tile_read_bytes = 0;
  int inc_max = 0;                     // (node, incident element) records per tile: the tiling budget
  int geo_stride = 0;                  // shared-memory plane stride of the records: 0 (mod 8), > every slot
  int slot_mode = 0;                   // rec_slot mode the codes were emitted with
  int rec_max = 0, pair_max = 0, con_max = 0, item_max = 0, nod_max = 0;  // distinct elements / pairs / contributions / work items / distinct nodes per tile, upper bounds
  int64_t n_items = 0;
  int item_len = 0;                    // contribution codes per work item (= tile_item_len(N))
  fes::DevBuf<int32_t> tile_node0;     // 2 * n_tiles : tile t covers nodes [tile_node0[2t], tile_node0[2t+1])
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
  size_t bytes() const {
    return tile_node0.bytes() + tile_elems.bytes() + rec_local.bytes() + tile_nodes.bytes() + tile_cnt.bytes() +
           tile_hdr.bytes() + pair_meta.bytes() + pair_codes.bytes();
  }
};

// Stencil runs: maximal ranges of consecutive nodes whose RELATIVE stencil repeats with period P (1 or 2)
// -- same neighbour offsets, same contributing elements up to a constant element-id shift per period, same
// local indices.  Structured meshes are made of them (a mesh row of create_rect_mesh is one run of period 2,
// a z-line of create_box_mesh one run of period 1); the detection reads the connectivity only, so a
// structured mesh with arbitrary (jittered) coordinates qualifies.  Run tiles are assembled by
// k_assemble_runs from a per-template program instead of per-tile index streams: lane t of a warp owns
// lane-step t of the tile, every shared-memory address is affine in t (unit stride: no bank conflicts), and
// nothing but the tile header, the coordinates and the per-element coefficients is read from HBM.
struct fes_runs {
  bool active = false;
  int64_t n_tiles = 0, n_nodes_covered = 0, n_templates = 0;
  // shared-memory sizing of k_assemble_runs: maxima over the templates
  int max_rec_entries = 0;             // 32 zero entries + n_lines * N * TS
  int max_coord_nodes = 0, max_img_units = 0, max_lines = 0;
  static constexpr int kTileInts = 8;  // {v0, T, template word offset, E0, first pair, raw word offset, 0, 0}
  // Every template program of the plan (layout: runs_template.hpp) sits in one slot of the kernel's __constant__
  // memory: the warp-uniform program reads of phase 2 and the line / span descriptors are constant-bank loads and
  // cost no shared-memory or L1 bandwidth.  The slot is owned from plan creation to destruction.
  static constexpr int kMaxTmplWords = 5120, kSlots = 3;
  int slot = -1;
  int n_warps = 4;                     // warps per CTA of the run kernel (= pair groups of the templates): 4 or 8
  // The remainder tiles (fes_plan::rest) run concurrently with the run kernel on this side stream, forked from and
  // joined back into the caller's stream with the two events (back to back on one stream the small kernel's drain
  // and, with programmatic dependent launch, its early-resident CTAs cost 8-40 us per assembly).
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  fes::DevBuf<int32_t> tiles;
  size_t bytes() const { return tiles.bytes(); }
  ~fes_runs();
};
namespace fes {
int runs_slot_acquire(const int32_t* words, size_t n_words, cudaStream_t stream, int* slot);  // assemble.cu; *slot = -1: none free
void runs_slot_release(int slot);
int ensure_full_tiling(fes_plan* P, const int32_t* d_elem_nodes, cudaStream_t stream);  // plan.cu
}
inline fes_runs::~fes_runs() {
  if (slot >= 0) fes::runs_slot_release(slot);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (side) cudaStreamDestroy(side);
}

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

  // incident-element lists per node (shared by both tilings and the stencil-run detection)
  int64_t n_geo = 0;
  fes::DevBuf<int32_t> ne_ptr;         // n_nodes + 1 : incident-element ranges per node
  fes::DevBuf<uint32_t> ne_code;       // n_geo : e*8 + a, ascending e per node
  fes_tileset full;                    // every node: forms / element types the run kernel does not cover; built at plan
                                       // creation when the mesh has no runs, else on first use (ensure_full_tiling)
  bool full_built = false;
  std::mutex full_mutex;
  std::vector<int32_t> h_ne, h_np;     // host copies of ne_ptr / node_ptr, kept for that deferred build
  fes_tileset rest;                    // the nodes outside stencil runs (used together with `runs`)
  fes_runs runs;

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
