// Element matrices and fused CSR assembly.
//
// fes_element_matrices replaces Form.element_matrices (ref fem/forms.py:202-209,245-247,254-258,
// 323-333) + Element._affine_geometry (ref fem/elements.py:174-212).
// fes_assemble replaces FunctionSpace.assemble = element_matrices + _ScatterPlan.scatter
// (ref fem/space.py:133-145,679-696,791-803) as ONE kernel: a thread owns one stored node pair
// (a C x C block of the CSR), walks that pair's contributions in ascending element id -- the
// order np.bincount(group, weights=Ke.ravel()[order]) adds them -- evaluates each
// element block in registers and writes every CSR value exactly once.  No atomics, so the
// result is deterministic; element matrices never exist in HBM.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "elements.cuh"
#include "plan.cuh"
#include "runs_template.hpp"
#include "tma.cuh"

namespace fes {
namespace {

constexpr int kBlock = 128;

// Lame parameters per unit modulus are computed once on the host (no fp64 divide per element):
// mu = E * mu1, lam = E * lam1  (ref fem/materials.py:24-28)
struct Coeffs {
  double E, mu1, lam1, factor;
  const double* d_E;
  const double* d_scale;
  double lam_over_mu;  // = 2 nu / (1 - 2 nu)
  double sign;         // sign of the global factor: the tiled kernel works with |factor|, a negate pass follows
};

__device__ __forceinline__ void element_coeffs(const Coeffs& c, int64_t e, Material& m, double& w) {
  const double E = c.d_E ? __ldg(c.d_E + e) : c.E;
  m = Material{E * c.mu1, E * c.lam1};
  w = c.d_scale ? c.factor * __ldg(c.d_scale + e) : c.factor;
}

// ---- unfused element matrices: one thread per (element, a, b) block -------------------------
template <int ELEM, int FORM, int C, int SD>
__global__ void k_element_matrices(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords,
                                   Coeffs cf, double* __restrict__ Ke) {
  constexpr int N = ET<ELEM>::N, K = N * C;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_el * N * N) return;
  const int64_t e = t / (N * N);
  const int ab = (int)(t - e * N * N), a = ab / N, b = ab - a * N;
  Geom<ET<ELEM>::RD, SD> g;
  load_geom<ELEM, SD>(en + e * N, coords, g);
  Material m; double w;
  element_coeffs(cf, e, m, w);
  double blk[C][C];
  eval_block<ELEM, FORM, C, SD>(g, a, b, m, w, blk);
  double* out = Ke + e * K * K;
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) out[(a * C + i) * K + (b * C + j)] = blk[i][j];
}

template <int ELEM, int SD>
__global__ void k_volumes(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords,
                          double* __restrict__ vol) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  Geom<ET<ELEM>::RD, SD> g;
  load_geom<ELEM, SD>(en + e * ET<ELEM>::N, coords, g);
  vol[e] = g.scale * ref_measure(ET<ELEM>::RD);
}

// ---- fused row-gather assembly: one thread per stored node pair ------------------------------
template <int ELEM, int FORM, int C, int SD>
__global__ void __launch_bounds__(kBlock)
k_assemble_pairs(int64_t n_pairs, const uint32_t* __restrict__ pair_start, const int32_t* __restrict__ pair_row,
                 const int32_t* __restrict__ node_ptr, const uint32_t* __restrict__ contrib,
                 const int32_t* __restrict__ indptr, const int32_t* __restrict__ en,
                 const double* __restrict__ coords, Coeffs cf, double* __restrict__ values) {
  constexpr int N = ET<ELEM>::N, N2 = N * N;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  double acc[C][C];
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) acc[i][j] = 0.0;
  const uint32_t k0 = pair_start[p], k1 = pair_start[p + 1];
  for (uint32_t k = k0; k < k1; ++k) {
    const uint32_t code = __ldg(contrib + k);
    const uint32_t e = code / N2, ab = code - e * N2, a = ab / N, b = ab - a * N;
    Geom<ET<ELEM>::RD, SD> g;
    load_geom<ELEM, SD>(en + (int64_t)e * N, coords, g);
    Material m; double w;
    element_coeffs(cf, e, m, w);
    double blk[C][C];
    eval_block<ELEM, FORM, C, SD>(g, (int)a, (int)b, m, w, blk);
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) acc[i][j] = __dadd_rn(acc[i][j], blk[i][j]);  // plain +=, never an fma
  }
  const int32_t v = pair_row[p];
  const int32_t s = (int32_t)(p - node_ptr[v]);
#pragma unroll
  for (int i = 0; i < C; ++i) {
    double* row = values + indptr[C * v + i] + C * s;
#pragma unroll
    for (int j = 0; j < C; ++j) row[j] = acc[i][j];
  }
}


__global__ void k_negate(double* __restrict__ v, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) v[t] = -v[t];
}

// ---- fused assembly, tiled (the production kernel) ------------------------------------------
// Persistent CTAs; each takes tiles of consecutive nodes (= a contiguous range of CSR values).
//   staging: one producer lane drives three rings of 1-D TMA bulk copies (cp.async.bulk, SASS
//            UBLKCP) completing on mbarriers: tile headers (3 tiles ahead), the distinct-node list
//            (2 ahead), the per-element corner indices / element ids (1 ahead, two buffers) and the
//            work-item words / contribution codes (one buffer, refilled right after phase 2 so the
//            copy lands under the next tile's phase 1).  Nothing the compute threads touch comes from a dependent
//            global load except the node coordinates, and those are fetched one tile ahead into
//            registers so they fly under phase 2.
//   phase 1: one thread per DISTINCT element of the tile evaluates the geometry once -- cofactors,
//            one rsqrt (no divide, no sqrt), every node's physical gradient at every quadrature
//            point pre-multiplied by sqrt(w_q mu_e vol_e) -- into shared-memory planes
//            (element-minor, swizzled slots; the plane stride is 0 mod 8, so a gradient's bank depends
//            on its slot only and the slot sequences of neighbouring lanes decide the conflicts).
//   phase 2: one thread per work item = a stored pair (v, w), or one of the 2^lg adjacent lanes that
//            share a pair with many contributions.  A lane walks its contributions in ascending
//            element id, accumulates the raw outer products S += g~_a (x) g~_b in registers; a fixed
//            xor-shuffle tree adds the lanes of a group; the writer lane applies
//            K = (lam/mu) S + S^T + tr(S) I once and writes the block into the shared-memory image
//            of the tile's CSR values.
//   store:   one TMA bulk store (cp.async.bulk.global.shared::cta) writes the image to HBM in full
//            lines; it drains while the next tile is evaluated.
// Every CSR value is written exactly once; no atomics and a fixed summation tree, so results are
// bit-reproducible, and the blocks of (v, w) and (w, v) are bitwise transposes of each other.
template <int ELEM, int FORM, int C, int SD>
struct RecLayout {
  static constexpr int N = ET<ELEM>::N, NQ = ET<ELEM>::NQ;
  // P2 gradient forms: the element is affine, so every shape gradient at every quadrature point is a
  // fixed combination of the RD rows of the inverse Jacobian -- the record holds just those rows
  // (4 doubles for the 6-node triangle instead of 36) and phase 2 contracts them with the constant
  // per-node-pair table c_p2tri_pair
  static constexpr bool kRows = ET<ELEM>::P2 && FORM != FES_FORM_MASS;
  static constexpr int NG = kRows ? ET<ELEM>::RD : NQ * N;  // gradient vectors per record
  // per gradient vector: one double2 plane (x, y) + a double plane (z) in 3D; mass: one weight
  static constexpr int NV2 = (FORM == FES_FORM_MASS) ? 0 : NG * (SD / 2);
  static constexpr int NV1 = (FORM == FES_FORM_MASS) ? 1 : NG * (SD & 1);
  static constexpr int DOUBLES = 2 * NV2 + NV1;
};

// CTA size, measured (tools/exp_sweep.sh): 4 warps per CTA and more CTAs per SM win for the gradient
// forms; the mass form (hardly any arithmetic) prefers 8 warps.  The environment variable
// FES_TILE_THREADS (128 / 256) overrides both at run time.
template <int FORM>
constexpr int tile_threads() {
  return FORM == FES_FORM_MASS ? 256 : 128;
}
// records a thread evaluates per pass of phase 1 (memory-level parallelism); one for the register-heavy tet geometry
// (measured: tets 0.5625 -> 0.5503 ms) and for 8-warp CTAs
template <int T, int SD> constexpr int rec_batch() { return (T >= 256 || SD == 3) ? 1 : 2; }
#ifndef FES_TET_CH
#define FES_TET_CH 2  // contributions gathered per chunk on tets (loads issued before their FMAs)
#endif
constexpr int kHdrRing = 4, kNodRing = 3;
__host__ __device__ constexpr int item_len_of(int N) { return (N == 3) ? 2 : (N == 4) ? 6 : 2; }  // = tile_item_len(N)
constexpr int kHdrBytes = fes_tileset::kHdrInts * 4;

__host__ __device__ constexpr size_t up16(size_t b) { return (b + 15) & ~(size_t)15; }

// shared-memory carve-up: geometry records, node coordinates, the value image, the work-item streams
// (one copy each), two copies of the per-element streams, the node-list ring and the header ring
template <int ELEM, int FORM, int C, int SD>
struct TileSmem {
  using R = RecLayout<ELEM, FORM, C, SD>;
  __host__ __device__ static size_t rec_bytes(int gs) { return up16((size_t)gs * R::DOUBLES * 8); }
  __host__ __device__ static size_t xyz_bytes(int nmax) { return up16((size_t)nmax * SD * 8); }   // distinct-node coordinates
  __host__ __device__ static size_t out_bytes(int pmax) { return up16(((size_t)pmax * C * C + 2) * 8); }
  static constexpr int L = item_len_of(ET<ELEM>::N);
  __host__ __device__ static size_t meta_bytes(int imax) { return up16((size_t)imax * 4 + 16); }
  __host__ __device__ static size_t code_bytes(int imax) { return up16((size_t)imax * L * 4 + 32); }
  __host__ __device__ static size_t loc_bytes(int rmax) { return up16((size_t)rmax * 4 + 32); }
  __host__ __device__ static size_t stage_bytes(int rmax) { return 2 * loc_bytes(rmax); }   // phase-1 streams (x2)
  __host__ __device__ static size_t item_bytes(int imax) { return meta_bytes(imax) + code_bytes(imax); }  // phase-2 streams (x1)
  __host__ __device__ static size_t nod_bytes(int nmax) { return up16((size_t)nmax * 4); }       // distinct-node ids
  // The mass form's phase 1 is a handful of instructions, so a copy issued after phase 2 does not land under
  // it: mass double-buffers the work-item streams (P2 vector mass 0.190 -> 0.178 ms).  The P2 gradient forms
  // show the same wait (ncu: 10 % of the samples) but lose more to the CTA the second buffer costs
  // (0.243 vs 0.238 ms), so they keep one buffer like the P1 forms.
  static constexpr int kItemBufs = (FORM == FES_FORM_MASS) ? 2 : 1;
  __host__ __device__ static size_t total(int gs, int rmax, int pmax, int imax, int nmax) {
    return rec_bytes(gs) + xyz_bytes(nmax) + out_bytes(pmax) + 2 * stage_bytes(rmax) + kItemBufs * item_bytes(imax) + kNodRing * nod_bytes(nmax) +
           kHdrRing * up16(kHdrBytes);
  }
};

struct TileDims {
  int geo_stride, rec_max, pair_max, item_max, nod_max, slot_mode;  // nod_max: distinct nodes per tile, upper bound
};

template <int ELEM, int FORM, int C, int SD, int kTileThreads>
__global__ void __launch_bounds__(kTileThreads, kTileThreads >= 256 ? 3 : 1)
k_assemble_tiles(const int32_t* __restrict__ tile_hdr, int n_tiles, const uint32_t* __restrict__ tile_elems,
                 const uint32_t* __restrict__ rec_local, const int32_t* __restrict__ tile_nodes,
                 const uint32_t* __restrict__ pair_meta, const uint32_t* __restrict__ pair_codes,
                 const double* __restrict__ coords, Coeffs cf, TileDims dims, int use_tma_store,
                 double* __restrict__ values) {
  using R = RecLayout<ELEM, FORM, C, SD>;
  using TS = TileSmem<ELEM, FORM, C, SD>;
  constexpr int N = ET<ELEM>::N, NQ = ET<ELEM>::NQ, RD = ET<ELEM>::RD;
  constexpr int kRecBatch = rec_batch<kTileThreads, SD>();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_h[kHdrRing], bar_n[kNodRing], bar_s[2], bar_c[2];
  __shared__ uint32_t s_skew[2], s_skew_c[2][2];
  const int GS = dims.geo_stride;
  [[maybe_unused]] const uint32_t inv_gs = (uint32_t)(((1u << 24) + (uint32_t)GS - 1u) / (uint32_t)GS);  // a = (i * inv_gs) >> 24 for i = a * GS + slot
  // planes: v2[plane * GS + slot] (double2), then v1[plane * GS + slot] (double)
  double2* v2 = reinterpret_cast<double2*>(smem_raw);
  double* v1 = reinterpret_cast<double*>(smem_raw) + (size_t)2 * R::NV2 * GS;
  unsigned char* sp = smem_raw + TS::rec_bytes(GS);
  double* s_xyz = reinterpret_cast<double*>(sp);                      sp += TS::xyz_bytes(dims.nod_max);
  double* s_out = reinterpret_cast<double*>(sp);                      sp += TS::out_bytes(dims.pair_max);
  unsigned char* stage0 = sp;
  const size_t stage_sz = TS::stage_bytes(dims.rec_max);
  sp += 2 * stage_sz;
  constexpr int kItemBufs = TS::kItemBufs;
  unsigned char* items0 = sp;                                         sp += kItemBufs * TS::item_bytes(dims.item_max);
  auto meta_of = [&](int k) { return items0 + (kItemBufs == 2 ? (k & 1) : 0) * TS::item_bytes(dims.item_max); };
  auto code_of = [&](int k) { return meta_of(k) + TS::meta_bytes(dims.item_max); };
  unsigned char* nod0 = sp;                                           sp += kNodRing * TS::nod_bytes(dims.nod_max);
  unsigned char* hdr0 = sp;
  const bool per_elem = cf.d_E != nullptr || cf.d_scale != nullptr;
  double sq_uniform = 0.0;  // sqrt(w mu ref_measure / NQ) when no per-element coefficient varies it
  if constexpr (FORM != FES_FORM_MASS) {
    double c2 = cf.factor * (ref_measure(RD) / (double)NQ);
    if constexpr (FORM == FES_FORM_ELASTIC) c2 *= cf.E * cf.mu1;
    sq_uniform = sqrt(c2);
  }

  const uint32_t v2_b = smem_u32(v2), v1_b = smem_u32(v1);  // shared addresses of the record planes
  // slot GS-1 of plane 0 is never a record (slots are < GS-1): the all-zero gradient that pads short items
  const uint32_t zero_code = ((uint32_t)(GS - 1) << 4) * 0x10001u;
  for (int pl = threadIdx.x; pl < R::NV2 + R::NV1; pl += kTileThreads) {  // in every plane (all quadrature points)
    if (pl < R::NV2) v2[(size_t)pl * GS + GS - 1] = make_double2(0.0, 0.0);
    else v1[(size_t)(pl - R::NV2) * GS + GS - 1] = 0.0;
  }
  asm volatile("griddepcontrol.launch_dependents;");  // the next grid of the stream may be scheduled as CTAs of this one retire
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
  if (n_my <= 0) return;
  auto tile_of = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };
  auto hdr_of = [&](int k) { return reinterpret_cast<const int4*>(hdr0 + (k & (kHdrRing - 1)) * up16(kHdrBytes)); };
  auto nod_of = [&](int k) { return reinterpret_cast<const int32_t*>(nod0 + (k % kNodRing) * TS::nod_bytes(dims.nod_max)); };
  // Producer duties are spread over the first lanes of different warps, so no warp carries the whole
  // serial issue sequence into the barriers: the LAST warp owns the image store (and the wait for
  // its reads; warp 0 gets the longest work items in phase 2), the others fill the rings.
  constexpr int kWarps = kTileThreads / 32;
  const bool lane0 = (threadIdx.x & 31) == 0;
  const int warp_id = threadIdx.x >> 5;
  const bool producer = threadIdx.x == kTileThreads - 32;                 // image store
  const bool p_ring = lane0 && warp_id == (kWarps > 2 ? kWarps - 2 : 0);  // headers + node lists
  const bool p_loc = lane0 && warp_id == (kWarps > 3 ? kWarps - 3 : 0);   // per-element streams
  const bool p_items = lane0 && warp_id == (kWarps > 2 ? 1 : 0);          // work-item streams

  // ---- producer side: ring fills
  auto issue_hdr = [&](int k) {
    uint64_t* bar = &bar_h[k & (kHdrRing - 1)];
    mbar_arrive_expect_tx(bar, kHdrBytes);
    bulk_g2s(const_cast<int4*>(hdr_of(k)), tile_hdr + (int64_t)tile_of(k) * fes_tileset::kHdrInts, kHdrBytes, bar);
  };
  auto issue_nod = [&](int k) {
    mbar_wait(&bar_h[k & (kHdrRing - 1)], (uint32_t)(k >> 2) & 1u);
    const uint32_t b_nod = ((uint32_t)hdr_of(k)[1].z * 4u + 15u) & ~15u;
    uint64_t* bar = &bar_n[k % kNodRing];
    mbar_arrive_expect_tx(bar, b_nod);
    if (b_nod) bulk_g2s(const_cast<int32_t*>(nod_of(k)), tile_nodes + (int64_t)tile_of(k) * fes_tileset::kTileNodeStride, b_nod, bar);
  };
  auto win = [](uint32_t i0, uint32_t count, uint32_t es) {
    return (uint32_t)(((((uint64_t)i0 * es) & 15) + (uint64_t)count * es + 15) & ~(uint64_t)15);
  };
  auto issue_loc = [&](int k) {  // phase-1 streams of tile k (double-buffered); its header landed before issue_nod(k) returned
    const int buf = k & 1;
    const int4 h0 = hdr_of(k)[0];
    unsigned char* s_loc = stage0 + buf * stage_sz;
    unsigned char* s_el = s_loc + TS::loc_bytes(dims.rec_max);
    const uint32_t g0 = (uint32_t)h0.x, ne = (uint32_t)h0.y;
    const uint32_t b_loc = win(g0, ne, 4);
    uint64_t* bar = &bar_s[buf];
    mbar_arrive_expect_tx(bar, b_loc * (per_elem ? 2u : 1u));
    uint32_t t;
    s_skew[buf] = stage_window<4>(s_loc, rec_local, g0, ne, bar, t);
    if (per_elem) stage_window<4>(s_el, tile_elems, g0, ne, bar, t);  // same skew as the corner indices
  };
  // phase-2 streams of tile k (item words + contribution codes): ONE buffer, refilled as soon as
  // phase 2 of tile k-1 is over, so the copy lands under phase 1 of tile k
  auto issue_items = [&](int k) {
    const int4 h2 = hdr_of(k)[2];
    constexpr uint32_t L = (uint32_t)TS::L;
    const uint32_t i0 = (uint32_t)h2.x, ni = (uint32_t)h2.y;
    const uint32_t b_meta = win(i0, ni, 4), b_code = win(i0 * L, ni * L, 4);
    const int ib = kItemBufs == 2 ? (k & 1) : 0;
    uint64_t* bar = &bar_c[ib];
    mbar_arrive_expect_tx(bar, b_meta + b_code);
    uint32_t t;
    s_skew_c[ib][1] = stage_window<4>(meta_of(k), pair_meta, i0, ni, bar, t);
    s_skew_c[ib][0] = stage_window<4>(code_of(k), pair_codes, i0 * L, ni * L, bar, t);
  };

  // coordinates of a tile's distinct nodes: global -> registers (<= 255 nodes: a node or two per
  // thread), stored to shared memory later so the loads fly under phase 2 of the previous tile
  constexpr int kNPT = (255 + kTileThreads - 1) / kTileThreads;
  double cx[kNPT][SD];
  auto load_xyz = [&](const int32_t* s_nod, int nd) {
#pragma unroll
    for (int k = 0; k < kNPT; ++k) {
      const int i = threadIdx.x + k * kTileThreads;
      if (i < nd) {
        const int32_t node = s_nod[i];
        if constexpr (SD == 2) {
          const double2 xy = __ldg(reinterpret_cast<const double2*>(coords) + node);
          cx[k][0] = xy.x; cx[k][1] = xy.y;
        } else {
#pragma unroll
          for (int d = 0; d < SD; ++d) cx[k][d] = __ldg(coords + (int64_t)node * SD + d);
        }
      }
    }
  };
  auto store_xyz = [&](int nd) {
#pragma unroll
    for (int k = 0; k < kNPT; ++k) {
      const int i = threadIdx.x + k * kTileThreads;
      if (i < nd) {
        if constexpr (SD == 2) {
          reinterpret_cast<double2*>(s_xyz)[i] = make_double2(cx[k][0], cx[k][1]);
        } else {
#pragma unroll
          for (int d = 0; d < SD; ++d) s_xyz[i * SD + d] = cx[k][d];
        }
      }
    }
  };

  if (producer) {
    for (int i = 0; i < kHdrRing; ++i) mbar_init(&bar_h[i], 1);
    for (int i = 0; i < kNodRing; ++i) mbar_init(&bar_n[i], 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(&bar_c[0], 1);
    mbar_init(&bar_c[1], 1);
    for (int k = 0; k < 3 && k < n_my; ++k) issue_hdr(k);
    issue_nod(0);
    if (n_my > 1) issue_nod(1);
    issue_loc(0);
    issue_items(0);
  }
  // Programmatic dependent launch: the next kernel of the stream may be scheduled as this grid's CTAs
  // retire (its prologue -- barrier init and the first plan-stream copies above, which read only the
  // immutable plan -- then overlaps this grid's tail); nothing a previous kernel may have written
  // (coordinates, per-element coefficients, the CSR values) is touched before the wait below.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  mbar_wait(&bar_h[0], 0);
  mbar_wait(&bar_n[0], 0);
  load_xyz(nod_of(0), hdr_of(0)[1].z);
  store_xyz(hdr_of(0)[1].z);
  __syncthreads();

  // Steady state, tile k:  [header(k), coordinates(k) in shared memory; streams(k) in flight or landed]
  //   producer: header(k+3), node list(k+2), streams(k+1)  |  phase 1(k)  |  barrier  |
  //   coordinates(k+1) -> registers  |  phase 2(k)  |  coordinates(k+1) -> shared memory  |  barrier  |  bulk store(k)
  for (int k = 0; k < n_my; ++k) {
    const int buf = k & 1;
    const bool has_next = k + 1 < n_my;
    if (p_ring) {
      if (k + 3 < n_my) issue_hdr(k + 3);
      if (k + 2 < n_my) issue_nod(k + 2);
    }
    if (p_loc && has_next) {
      mbar_wait(&bar_h[(k + 1) & (kHdrRing - 1)], (uint32_t)((k + 1) >> 2) & 1u);  // its header: issued three tiles ago
      issue_loc(k + 1);
    }
    if constexpr (kItemBufs == 2) {  // the other buffer was last read in phase 2 of tile k-1
      if (p_items && has_next) {
        mbar_wait(&bar_h[(k + 1) & (kHdrRing - 1)], (uint32_t)((k + 1) >> 2) & 1u);
        issue_items(k + 1);
      }
    }
    const int4 h0 = hdr_of(k)[0], h2 = hdr_of(k)[2];
    const int32_t ne = h0.y, n_pairs = h0.w, n_items = h2.y;
    const int64_t val0 = (int64_t)C * C * h0.z;          // first CSR value of the tile
    const int skew_d = (int)(val0 & 1);                   // image slot i <-> values[val0 + i]; equal 16-byte phase
    unsigned char* s_loc = stage0 + buf * stage_sz;
    unsigned char* s_el = s_loc + TS::loc_bytes(dims.rec_max);
    mbar_wait(&bar_s[buf], (uint32_t)(k >> 1) & 1u);  // issued a whole tile ago
    const uint32_t* locs = reinterpret_cast<const uint32_t*>(s_loc + s_skew[buf]);
    const uint32_t* elems = reinterpret_cast<const uint32_t*>(s_el + s_skew[buf]);

    // ---- phase 1: one geometry record per distinct element (corner coordinates from shared memory).
    // Pass b starts at thread b*T/2, so a short last pass does not land on warp 0 every time.
    for (int32_t blk = 0; blk * kTileThreads < ne; blk += kRecBatch) {
      bool live[kRecBatch];
      int32_t gi[kRecBatch];
      uint32_t e[kRecBatch];
      double X[kRecBatch][RD + 1][SD];
#pragma unroll
      for (int r = 0; r < kRecBatch; ++r) {
        gi[r] = (blk + r) * kTileThreads + (int)((threadIdx.x + (unsigned)(blk + r) * (kTileThreads / 2)) % kTileThreads);
        live[r] = gi[r] < ne;
        const uint32_t loc = live[r] ? locs[gi[r]] : 0u;
        e[r] = (live[r] && per_elem) ? elems[gi[r]] : 0u;
#pragma unroll
        for (int n = 0; n <= RD; ++n) {
          const int li = (loc >> (8 * n)) & 255;
          if constexpr (SD == 2) {
            const double2 xy = reinterpret_cast<const double2*>(s_xyz)[li];
            X[r][n][0] = xy.x; X[r][n][1] = xy.y;
          } else {
#pragma unroll
            for (int s = 0; s < SD; ++s) X[r][n][s] = s_xyz[li * SD + s];
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRecBatch; ++r) {
        if (!live[r]) continue;
        Material m; double w;
        element_coeffs(cf, e[r], m, w);
        const int rec = (int)rec_slot((uint32_t)gi[r], dims.slot_mode);
        if constexpr (FORM == FES_FORM_MASS) {
          Geom<RD, SD> g;
          make_geom<RD, SD>(X[r], g);
          v1[rec] = w * (g.scale * ref_measure(RD));
        } else {
          // g~ = sqrt(w_q mu_e vol) grad(phi) with grad(phi) = cof / det and vol = |det| ref_measure:
          //    = cof * sign(det) * sqrt(w mu ref_measure / NQ) * rsqrt(|det|)  -- one rsqrt, no divide.
          // Weight and modulus ride on the gradients (lam/mu is applied in phase 2), so phase 2
          // accumulates plain outer products.
          double cof[RD][SD], det;
          make_cof<RD, SD>(X[r], cof, det);
          double f = sq_uniform, d = fabs(det);
          if (per_elem) {
            double c2 = w * (ref_measure(RD) / (double)NQ);
            if constexpr (FORM == FES_FORM_ELASTIC) c2 *= m.mu;
            // sqrt(c2) / sqrt(|det|) = c2 rsqrt(c2 |det|): one rsqrt per record either way (f * rsqrt(d) below); a void
            // element (weight 0) contributes exact zeros
            f = c2 > 0.0 ? c2 : 0.0;
            d = c2 > 0.0 ? c2 * d : 1.0;
          }
          f = copysign(f * rsqrt(d), det);
#pragma unroll
          for (int i = 0; i < RD; ++i)
#pragma unroll
            for (int s = 0; s < SD; ++s) cof[i][s] *= f;
          if constexpr (R::kRows) {
            static_assert(SD == 2, "row records are implemented for the P2 triangle");
#pragma unroll
            for (int rr = 0; rr < RD; ++rr) v2[rr * GS + rec] = make_double2(cof[rr][0], cof[rr][1]);
          } else
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int n = 0; n < N; ++n) {
              double gn[SD];
              shape_grad_rows<ELEM, SD>(cof, q, n, gn);
#pragma unroll
              for (int h = 0; h < SD / 2; ++h)
                v2[((q * N + n) * (SD / 2) + h) * GS + rec] = make_double2(gn[2 * h], gn[2 * h + 1]);
              if constexpr (SD & 1) v1[(q * N + n) * GS + rec] = gn[SD - 1];
            }
        }
      }
    }
    if (producer) bulk_wait_read_all();  // the previous tile's store has finished reading the image (measured: never the critical path)
    __syncthreads();                     // records visible, image free, coordinates consumed

    int nd_next = 0;
    if (has_next) {  // the next tile's coordinates start flying now and land under phase 2
      mbar_wait(&bar_h[(k + 1) & (kHdrRing - 1)], (uint32_t)((k + 1) >> 2) & 1u);
      mbar_wait(&bar_n[(k + 1) % kNodRing], (uint32_t)((k + 1) / kNodRing) & 1u);
      nd_next = hdr_of(k + 1)[1].z;
      load_xyz(nod_of(k + 1), nd_next);
    }

    // ---- phase 2: one thread per work item (explicit shared addresses, bases in registers)
    const int ib = kItemBufs == 2 ? (k & 1) : 0;
    mbar_wait(&bar_c[ib], (uint32_t)(kItemBufs == 2 ? (k >> 1) : k) & 1u);  // issued before phase 1 of this tile (a whole tile ago with two buffers)
    const uint32_t code_b = smem_u32(code_of(k)) + s_skew_c[ib][0];
    const uint32_t meta_b = smem_u32(meta_of(k)) + s_skew_c[ib][1];
    static_assert(TS::L % 2 == 0, "item codes are fetched two at a time");
    const uint32_t img_b = smem_u32(s_out) + 8u * (uint32_t)skew_d;
    // A chunk of CH contributions: S += g~_a (x) g~_b each.  A code holds the byte offsets of the two
    // gradients in the double2 planes (the z plane of a 3D gradient sits at half that offset).  All
    // loads of a chunk are issued before its FMAs; a lane with fewer contributions reads the
    // all-zero record instead of branching (fma(0, 0, acc) == acc, bit for bit).
    constexpr int kItemLen = TS::L;
    constexpr int CH = FORM == FES_FORM_MASS ? 1 : (R::kRows ? kItemLen : (NQ > 1 ? 1 : (SD == 3 ? FES_TET_CH : kItemLen)));
    auto gather_chunk = [&](const uint32_t (&code)[CH], double (&acc)[C][C]) {
      if constexpr (FORM == FES_FORM_MASS) {
        const uint32_t ia = (code[0] & 0xffffu) >> 4, ib = code[0] >> 20;
        const uint32_t a = (uint32_t)(((uint64_t)ia * inv_gs) >> 24), b = (uint32_t)(((uint64_t)ib * inv_gs) >> 24);
        acc[0][0] = fma(lds_f64(v1_b + 8u * (ia - a * (uint32_t)GS)), mass_ref<ELEM>((int)a, (int)b), acc[0][0]);
      } else if constexpr (R::kRows) {
        // S += sum_rs M_ab[r][s] G_r (x) G_s: two gradient rows per contribution (one LDS.128 each) and
        // the pair's 2 x 2 table from constant memory (neighbouring lanes mostly share (a, b))
        double2 g0[CH], g1[CH];
        uint32_t ab[CH];
#pragma unroll
        for (int x = 0; x < CH; ++x) {
          const uint32_t oa = code[x] & 0xffffu, ia = oa >> 4, ib = code[x] >> 20;
          const uint32_t a = (uint32_t)(((uint64_t)ia * inv_gs) >> 24), b = (uint32_t)(((uint64_t)ib * inv_gs) >> 24);
          const uint32_t so = oa - 16u * a * (uint32_t)GS;  // byte offset of the record's slot in plane 0
          g0[x] = lds_f64x2(v2_b + so);
          g1[x] = lds_f64x2(v2_b + so + 16u * (uint32_t)GS);
          ab[x] = a * 6u + b;
        }
#pragma unroll
        for (int x = 0; x < CH; ++x) {
          const double* m = c_p2tri_pair.m[ab[x]];
          const double m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
          // P_s = sum_r M[r][s] G_r ;  S += P_0 (x) G_0 + P_1 (x) G_1
          const double p0x = fma(m10, g1[x].x, m00 * g0[x].x), p0y = fma(m10, g1[x].y, m00 * g0[x].y);
          const double p1x = fma(m11, g1[x].x, m01 * g0[x].x), p1y = fma(m11, g1[x].y, m01 * g0[x].y);
          if constexpr (FORM == FES_FORM_LAPLACIAN) {
            acc[0][0] = fma(p0x, g0[x].x, acc[0][0]); acc[0][0] = fma(p0y, g0[x].y, acc[0][0]);
            acc[0][0] = fma(p1x, g1[x].x, acc[0][0]); acc[0][0] = fma(p1y, g1[x].y, acc[0][0]);
          } else {
            acc[0][0] = fma(p0x, g0[x].x, acc[0][0]); acc[0][0] = fma(p1x, g1[x].x, acc[0][0]);
            acc[0][1] = fma(p0x, g0[x].y, acc[0][1]); acc[0][1] = fma(p1x, g1[x].y, acc[0][1]);
            acc[1][0] = fma(p0y, g0[x].x, acc[1][0]); acc[1][0] = fma(p1y, g1[x].x, acc[1][0]);
            acc[1][1] = fma(p0y, g0[x].y, acc[1][1]); acc[1][1] = fma(p1y, g1[x].y, acc[1][1]);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          double2 ta[CH], tb[CH];
          [[maybe_unused]] double za[CH], zb[CH];
          // a warp of diagonal pairs (a == b in every lane; items are sorted, so such warps are whole):
          // one gradient per contribution instead of two
          bool same = true;
#pragma unroll
          for (int x = 0; x < CH; ++x) same = same && ((code[x] ^ (code[x] >> 16)) & 0xffffu) == 0u;
          same = __all_sync(0xffffffffu, same);
#pragma unroll
          for (int x = 0; x < CH; ++x) {
            const uint32_t oa = code[x] & 0xffffu, ob = code[x] >> 16;
            ta[x] = lds_f64x2(v2_b + oa + (uint32_t)(q * N * 16) * (uint32_t)GS);
            if constexpr (SD & 1) za[x] = lds_f64(v1_b + (oa >> 1) + (uint32_t)(q * N * 8) * (uint32_t)GS);
            if (!same) {
              tb[x] = lds_f64x2(v2_b + ob + (uint32_t)(q * N * 16) * (uint32_t)GS);
              if constexpr (SD & 1) zb[x] = lds_f64(v1_b + (ob >> 1) + (uint32_t)(q * N * 8) * (uint32_t)GS);
            } else {
              tb[x] = ta[x];
              if constexpr (SD & 1) zb[x] = za[x];
            }
          }
#pragma unroll
          for (int x = 0; x < CH; ++x) {
            double A_[SD], B_[SD];
            A_[0] = ta[x].x; A_[1] = ta[x].y;
            B_[0] = tb[x].x; B_[1] = tb[x].y;
            if constexpr (SD & 1) { A_[SD - 1] = za[x]; B_[SD - 1] = zb[x]; }
            if constexpr (FORM == FES_FORM_LAPLACIAN) {
#pragma unroll
              for (int s = 0; s < SD; ++s) acc[0][0] = fma(A_[s], B_[s], acc[0][0]);
            } else {
#pragma unroll
              for (int i = 0; i < C; ++i)
#pragma unroll
                for (int j = 0; j < C; ++j) acc[i][j] = fma(A_[i], B_[j], acc[i][j]);
            }
          }
        }
      }
    };
    // Items with several chunks (tets: 6 contributions in chunks of 2): software pipeline -- the gathers of
    // chunk c+1 are issued before the FMAs of chunk c, so a warp's own loads overlap its arithmetic (the tet
    // kernel runs 12 warps per SM; ncu showed it stalled on the LDS -> DFMA dependency, not on bandwidth).
    constexpr bool kPipelined = FORM != FES_FORM_MASS && !R::kRows && NQ == 1 && (kItemLen / CH) > 1;
    struct Operands {
      double2 ta[CH], tb[CH];
      double za[CH], zb[CH];
    };
    [[maybe_unused]] auto load_ops = [&](const uint32_t (&code)[CH], Operands& o) {
      bool same = true;
#pragma unroll
      for (int x = 0; x < CH; ++x) same = same && ((code[x] ^ (code[x] >> 16)) & 0xffffu) == 0u;
      same = __all_sync(0xffffffffu, same);
#pragma unroll
      for (int x = 0; x < CH; ++x) {
        const uint32_t oa = code[x] & 0xffffu, ob = code[x] >> 16;
        o.ta[x] = lds_f64x2(v2_b + oa);
        if constexpr (SD & 1) o.za[x] = lds_f64(v1_b + (oa >> 1));
        if (!same) {
          o.tb[x] = lds_f64x2(v2_b + ob);
          if constexpr (SD & 1) o.zb[x] = lds_f64(v1_b + (ob >> 1));
        } else {
          o.tb[x] = o.ta[x];
          if constexpr (SD & 1) o.zb[x] = o.za[x];
        }
      }
    };
    [[maybe_unused]] auto fma_ops = [&](const Operands& o, double (&acc)[C][C]) {
#pragma unroll
      for (int x = 0; x < CH; ++x) {
        double A_[SD], B_[SD];
        A_[0] = o.ta[x].x; A_[1] = o.ta[x].y;
        B_[0] = o.tb[x].x; B_[1] = o.tb[x].y;
        if constexpr (SD & 1) { A_[SD - 1] = o.za[x]; B_[SD - 1] = o.zb[x]; }
        if constexpr (FORM == FES_FORM_LAPLACIAN) {
#pragma unroll
          for (int s_ = 0; s_ < SD; ++s_) acc[0][0] = fma(A_[s_], B_[s_], acc[0][0]);
        } else {
#pragma unroll
          for (int i = 0; i < C; ++i)
#pragma unroll
            for (int j = 0; j < C; ++j) acc[i][j] = fma(A_[i], B_[j], acc[i][j]);
        }
      }
    };
    for (int t0 = 0; t0 < n_items; t0 += kTileThreads) {  // warp-uniform trip count: the shuffles need every lane
      const int t = t0 + (int)threadIdx.x;
      const bool live = t < n_items;
      const uint32_t mt = live ? lds_u32(meta_b + 4u * (uint32_t)t) : 0u;
      const uint32_t lg = (mt >> 28) & 7u;
      uint32_t cd[kItemLen];  // the item's codes: exactly kItemLen, padded with the zero record by the plan
#pragma unroll
      for (int kk = 0; kk < kItemLen; kk += 2) {
        const uint2 two = live ? lds_u32x2(code_b + 4u * ((uint32_t)t * kItemLen + kk)) : make_uint2(zero_code, zero_code);
        cd[kk] = two.x; cd[kk + 1] = two.y;
      }
      double acc[C][C];
#pragma unroll
      for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) acc[i][j] = 0.0;
      if constexpr (kPipelined) {
        constexpr int NC = kItemLen / CH;
        Operands cur, nxt;
        uint32_t part[CH];
#pragma unroll
        for (int x = 0; x < CH; ++x) part[x] = cd[x];
        load_ops(part, cur);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          bool more = false;
          if (c + 1 < NC) {  // later chunks are padding for every lane of most warps (items are sorted by trips)
#pragma unroll
            for (int x = 0; x < CH; ++x) part[x] = cd[(c + 1) * CH + x];
            more = __any_sync(0xffffffffu, part[0] != zero_code);
            if (more) load_ops(part, nxt);
          }
          fma_ops(cur, acc);
          if (!more) break;
          cur = nxt;
        }
      } else {
#pragma unroll
        for (int c0_ = 0; c0_ < kItemLen; c0_ += CH) {
          uint32_t part[CH];
#pragma unroll
          for (int x = 0; x < CH; ++x) part[x] = cd[c0_ + x];
          // later chunks are padding for every lane of most warps (items are sorted by trips)
          if (c0_ == 0 || __any_sync(0xffffffffu, part[0] != zero_code)) gather_chunk(part, acc);
        }
      }
      // lanes that share a pair: fixed xor tree over the 2^lg adjacent lanes (both partners add, so
      // every lane of the group ends with the same bits)
      const uint32_t lg_max = __reduce_max_sync(0xffffffffu, lg);
      for (uint32_t o = 0; o < lg_max; ++o) {
        constexpr int NACC = (FORM == FES_FORM_ELASTIC) ? C * C : 1;
#pragma unroll
        for (int x = 0; x < NACC; ++x) {
          double& a_ = acc[x / C][x % C];
          const double other = __shfl_xor_sync(0xffffffffu, a_, 1 << o);
          if (o < lg) a_ += other;
        }
      }
      if (!(mt >> 31)) continue;  // not the writer lane of its pair
      double K[C][C];
      if constexpr (FORM == FES_FORM_ELASTIC) {
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < C; ++i) tr += acc[i][i];
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
          for (int j = 0; j < C; ++j) K[i][j] = fma(cf.lam_over_mu, acc[i][j], acc[j][i]) + (i == j ? tr : 0.0);
      } else if constexpr (FORM == FES_FORM_LAPLACIAN) {
        K[0][0] = acc[0][0];
      } else {
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
          for (int j = 0; j < C; ++j) K[i][j] = (i == j) ? acc[0][0] : 0.0;
      }
      // row i of the block starts at double C * (q + i * deg) of the tile's value image
      const uint32_t q0 = mt & 0xffffu, deg = (mt >> 16) & 0xfffu;
      if constexpr (C == 2) {
        // neighbouring threads hold the same slot of consecutive nodes, 2*deg 16-byte slots apart:
        // lanes 4..7 of each group of 8 write their rows in the opposite order, which makes the
        // eight addresses of a quarter-warp hit eight different bank groups when deg is odd
        const bool swap = (threadIdx.x >> 2) & 1;
        const uint32_t r0 = img_b + 16u * (q0 + (swap ? deg : 0u)), r1 = img_b + 16u * (q0 + (swap ? 0u : deg));
        sts_f64x2(r0, swap ? K[1][0] : K[0][0], swap ? K[1][1] : K[0][1]);
        sts_f64x2(r1, swap ? K[0][0] : K[1][0], swap ? K[0][1] : K[1][1]);
      } else {
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
          for (int j = 0; j < C; ++j) sts_f64(img_b + 8u * (C * (q0 + i * deg) + j), K[i][j]);
      }
    }
    if (has_next) store_xyz(nd_next);  // the coordinate buffer was last read before this tile's first barrier

    // ---- store: the image is the tile's contiguous slice of `values`
    const int n_out = C * C * n_pairs;
    double* img = s_out + skew_d;
    fence_proxy_async_smem();  // generic-proxy accesses (image writes, item reads) ordered before the async proxy's
    __syncthreads();           // phase 2 is over: the image is complete, the item buffer is free
    if (kItemBufs == 1 && p_items && has_next) issue_items(k + 1);
    if (use_tma_store) {
      if (producer) {
        const int head = skew_d;                           // an odd first slot is 16-byte misaligned on both sides
        const int body = (n_out - head) & ~1, tail = n_out - head - body;
        if (body > 0) bulk_s2g(values + val0 + head, img + head, (uint32_t)body * 8u);
        if (head) values[val0] = img[0];
        if (tail) values[val0 + n_out - 1] = img[n_out - 1];
      }
    } else {
      for (int i = threadIdx.x; i < n_out; i += kTileThreads) values[val0 + i] = img[i];
      __syncthreads();
    }
  }
  if (producer) bulk_wait_read_all();  // shared memory must outlive the last store's reads
}

// PrecomputedForm: gather caller-supplied (n_el, K, K) matrices, optionally scaled per element
template <int C>
__global__ void __launch_bounds__(kBlock)
k_assemble_precomputed(int64_t n_pairs, int N, const uint32_t* __restrict__ pair_start,
                       const int32_t* __restrict__ pair_row, const int32_t* __restrict__ node_ptr,
                       const uint32_t* __restrict__ contrib, const int32_t* __restrict__ indptr,
                       const double* __restrict__ Ke, const double* __restrict__ scale, double factor,
                       double* __restrict__ values) {
  const int N2 = N * N, K = N * C;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  double acc[C][C];
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) acc[i][j] = 0.0;
  const uint32_t k0 = pair_start[p], k1 = pair_start[p + 1];
  for (uint32_t k = k0; k < k1; ++k) {
    const uint32_t code = __ldg(contrib + k);
    const uint32_t e = code / N2, ab = code - e * N2, a = ab / N, b = ab - a * N;
    const double w = scale ? factor * __ldg(scale + e) : factor;
    const double* src = Ke + (int64_t)e * K * K;
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j)
        acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(w, __ldg(src + (a * C + i) * K + (b * C + j))));
  }
  const int32_t v = pair_row[p];
  const int32_t s = (int32_t)(p - node_ptr[v]);
#pragma unroll
  for (int i = 0; i < C; ++i) {
    double* row = values + indptr[C * v + i] + C * s;
#pragma unroll
    for (int j = 0; j < C; ++j) row[j] = acc[i][j];
  }
}

// ---- dispatch over (element, form, n_comp, spatial_dim) --------------------------------------
struct Job {
  int mode;  // 0 = element matrices, 1 = fused assembly
  const fes_plan* plan;
  const int32_t* en;
  int64_t n_el;
  const double* coords;
  Coeffs cf;
  double* out;
  cudaStream_t stream;
  bool untiled = false;  // force the reference-order (untiled) kernel
};

#include "assemble_runs.cuh"

template <int ELEM, int FORM, int C, int SD, int kTileThreads>
int launch_tiles(const Job& j, const fes_tileset& T, cudaStream_t stream, bool pdl = true) {
  const fes_tileset* P = &T;
  const TileDims dims{P->geo_stride, P->rec_max, P->pair_max, P->item_max, P->nod_max, P->slot_mode};
  const size_t smem = TileSmem<ELEM, FORM, C, SD>::total(dims.geo_stride, dims.rec_max, dims.pair_max, dims.item_max,
                                                         dims.nod_max);
  static size_t configured = 0;  // per template instantiation: largest opt-in so far
  static size_t occ_smem = 0;
  static int per_sm = 0, sms = 0;
  if (per_sm == 0 || smem != occ_smem) {
    if (smem > configured && smem + 1024 > 48 * 1024)  // static smem counts toward the 48 KB default
      FES_CUDA(cudaFuncSetAttribute(k_assemble_tiles<ELEM, FORM, C, SD, kTileThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    configured = std::max(configured, smem);
    int dev = 0;
    FES_CUDA(cudaGetDevice(&dev));
    FES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_tiles<ELEM, FORM, C, SD, kTileThreads>,
                                                           kTileThreads, smem));
    occ_smem = smem;
    if (per_sm < 1) {
      per_sm = 0;
      return set_error(FES_ERR_CUDA, "assembly kernel does not fit on an SM (%zu B smem)", smem);
    }
    if (getenv("FES_PLAN_DEBUG"))
      fprintf(stderr, "[fes assemble] threads=%d smem=%zu B CTAs/SM=%d\n", kTileThreads, smem, per_sm);
  }
  Coeffs cf = j.cf;
  if (FORM != FES_FORM_MASS && cf.factor < 0.0) { cf.factor = -cf.factor; cf.sign = -1.0; }
  const int64_t grid = std::min<int64_t>(P->n_tiles, (int64_t)sms * per_sm);  // persistent CTAs
  static const bool no_tma = getenv("FES_NO_TMA_STORE") != nullptr;  // tuning knob: plain coalesced stores
  const int use_tma = (!no_tma && (reinterpret_cast<uintptr_t>(j.out) & 15) == 0) ? 1 : 0;
  static const bool no_pdl = getenv("FES_NO_PDL") != nullptr;  // tuning knob
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(kTileThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = (no_pdl || !pdl) ? 0 : 1;
  FES_CUDA(cudaLaunchKernelEx(&lc, k_assemble_tiles<ELEM, FORM, C, SD, kTileThreads>, (const int32_t*)P->tile_hdr.p,
                              (int)P->n_tiles, (const uint32_t*)P->tile_elems.p, (const uint32_t*)P->rec_local.p,
                              (const int32_t*)P->tile_nodes.p, (const uint32_t*)P->pair_meta.p,
                              (const uint32_t*)P->pair_codes.p, (const double*)j.coords, cf, dims, use_tma, j.out));
  FES_LAUNCH_CHECK();
  return FES_OK;
}

template <int ELEM, int FORM, int C, int SD>
int run(const Job& j) {
  constexpr int N = ET<ELEM>::N;
  if (j.mode == 0) {
    const int64_t n = j.n_el * N * N;
    if (n == 0) return FES_OK;
    k_element_matrices<ELEM, FORM, C, SD><<<grid_for(n, 256), 256, 0, j.stream>>>(j.en, j.n_el, j.coords, j.cf, j.out);
  } else {
    const fes_plan* P = j.plan;
    if (P->n_pairs == 0) return FES_OK;
    const bool no_runs = getenv("FES_NO_RUNS_LAUNCH") != nullptr;  // A-B knob, read per call: the tiled kernel on every node
    const bool use_runs = runs_supported<ELEM, FORM, C, SD>() && P->runs.active && !no_runs && !j.untiled &&
                          ((reinterpret_cast<uintptr_t>(j.out) | reinterpret_cast<uintptr_t>(j.coords)) & 15) == 0;  // 16-byte units
    if (!use_runs && !j.untiled && !P->full_built)  // a form the run kernel does not cover, on a plan that deferred the full tiling
      FES_TRY(ensure_full_tiling(const_cast<fes_plan*>(P), j.en, j.stream));
    if ((use_runs || P->full.tiled) && !j.untiled) {
      static const int env_threads = getenv("FES_TILE_THREADS") ? atoi(getenv("FES_TILE_THREADS")) : 0;  // tuning knob
      const int threads = env_threads ? env_threads : tile_threads<FORM>();
      const fes_tileset& T = use_runs ? P->rest : P->full;
      static const bool skip_rest = getenv("FES_EXP_SKIP_REST") != nullptr;  // timing experiment only: leaves the remainder rows unwritten
      if (use_runs) {
        // the remainder tiles on the plan's side stream, concurrently with the run kernel
        const fes_runs& R = P->runs;
        const bool rest = T.n_tiles > 0 && !skip_rest;
        if (rest) {
          FES_CUDA(cudaEventRecord(R.ev_fork, j.stream));
          FES_CUDA(cudaStreamWaitEvent(R.side, R.ev_fork, 0));
          if (threads == 256) FES_TRY((launch_tiles<ELEM, FORM, C, SD, 256>(j, T, R.side, false)));
          else FES_TRY((launch_tiles<ELEM, FORM, C, SD, 128>(j, T, R.side, false)));
          FES_CUDA(cudaEventRecord(R.ev_join, R.side));
        }
        FES_TRY((launch_runs<ELEM, FORM, C, SD>(j, j.stream)));
        if (rest) FES_CUDA(cudaStreamWaitEvent(j.stream, R.ev_join, 0));
      } else if (T.n_tiles > 0) {
        if (threads == 256) FES_TRY((launch_tiles<ELEM, FORM, C, SD, 256>(j, T, j.stream)));
        else FES_TRY((launch_tiles<ELEM, FORM, C, SD, 128>(j, T, j.stream)));
      }
      if (FORM != FES_FORM_MASS && j.cf.factor < 0.0)  // a negative global factor (ScaledForm): the kernels ran with |factor|
        k_negate<<<grid_for(P->nnz, 256), 256, 0, j.stream>>>(j.out, P->nnz);
      else
        return FES_OK;
    } else {
      k_assemble_pairs<ELEM, FORM, C, SD><<<grid_for(P->n_pairs, kBlock), kBlock, 0, j.stream>>>(
          P->n_pairs, P->pair_start.p, P->pair_row.p, P->node_ptr.p, P->contrib.p, P->indptr.p, j.en, j.coords, j.cf,
          j.out);
    }
  }
  FES_LAUNCH_CHECK();
  return FES_OK;
}

template <int ELEM, int SD>
int run_mass(const Job& j, int c) {
  switch (c) {
    case 1: return run<ELEM, FES_FORM_MASS, 1, SD>(j);
    case 2: return run<ELEM, FES_FORM_MASS, 2, SD>(j);
    case 3: return run<ELEM, FES_FORM_MASS, 3, SD>(j);
  }
  return set_error(FES_ERR_UNSUPPORTED, "mass form supports 1..3 components, got %d", c);
}

int dispatch(const Job& j, int elem, int form, int c, int sd) {
  const int rd = elem_ref_dim(elem);
  if (elem < FES_P1_LINE || elem > FES_P2_TRI) return set_error(FES_ERR_UNSUPPORTED, "unknown element kind %d", elem);
  if (form == FES_FORM_MASS) {
    if (elem == FES_P1_LINE && sd == 1) return run_mass<FES_P1_LINE, 1>(j, c);
    if (elem == FES_P1_LINE && sd == 2) return run_mass<FES_P1_LINE, 2>(j, c);
    if (elem == FES_P2_LINE && sd == 2) return run_mass<FES_P2_LINE, 2>(j, c);
    if (elem == FES_P1_TRI && sd == 2) return run_mass<FES_P1_TRI, 2>(j, c);
    if (elem == FES_P1_TRI && sd == 3) return run_mass<FES_P1_TRI, 3>(j, c);
    if (elem == FES_P2_TRI && sd == 2) return run_mass<FES_P2_TRI, 2>(j, c);
    if (elem == FES_P1_TET && sd == 3) return run_mass<FES_P1_TET, 3>(j, c);
    return set_error(FES_ERR_UNSUPPORTED, "no mass form for element kind %d in %dD", elem, sd);
  }
  if (sd != rd)
    return set_error(FES_ERR_UNSUPPORTED, "gradient forms need a volume element (kind %d has dim %d, got spatial_dim %d)",
                     elem, rd, sd);
  if (form == FES_FORM_LAPLACIAN) {
    if (c != 1) return set_error(FES_ERR_VALUE, "LaplacianForm is scalar; the space has %d components", c);
    if (elem == FES_P1_LINE) return run<FES_P1_LINE, FES_FORM_LAPLACIAN, 1, 1>(j);
    if (elem == FES_P1_TRI) return run<FES_P1_TRI, FES_FORM_LAPLACIAN, 1, 2>(j);
    if (elem == FES_P1_TET) return run<FES_P1_TET, FES_FORM_LAPLACIAN, 1, 3>(j);
    if (elem == FES_P2_TRI) return run<FES_P2_TRI, FES_FORM_LAPLACIAN, 1, 2>(j);
    return set_error(FES_ERR_UNSUPPORTED, "no Laplacian for element kind %d", elem);
  }
  if (form == FES_FORM_ELASTIC) {
    if (c != sd) return set_error(FES_ERR_VALUE, "LinearElasticForm needs %d components, the space has %d", sd, c);
    if (elem == FES_P1_TRI) return run<FES_P1_TRI, FES_FORM_ELASTIC, 2, 2>(j);
    if (elem == FES_P1_TET) return run<FES_P1_TET, FES_FORM_ELASTIC, 3, 3>(j);
    if (elem == FES_P2_TRI) return run<FES_P2_TRI, FES_FORM_ELASTIC, 2, 2>(j);
    return set_error(FES_ERR_UNSUPPORTED, "no elastic constitutive matrix for element kind %d", elem);
  }
  return set_error(FES_ERR_UNSUPPORTED, "unknown form %d", form);
}

int check_coeffs(const fes_coeffs* c, Coeffs& out) {
  if (!c) return set_error(FES_ERR_VALUE, "coeffs is NULL");
  out = Coeffs{c->E, 1.0 / (2.0 * (1.0 + c->nu)), c->nu / ((1.0 + c->nu) * (1.0 - 2.0 * c->nu)), c->factor, c->d_E,
               c->d_scale, 2.0 * c->nu / (1.0 - 2.0 * c->nu), 1.0};
  return FES_OK;
}

}  // namespace

// ---- template slots of the stencil-run kernel (c_run_tmpl, assemble_runs.cuh) ------------------
namespace {
std::atomic<int> g_slot_used[fes_runs::kSlots];
}
int runs_slot_acquire(const int32_t* words, size_t n_words, cudaStream_t stream, int* slot) {
  *slot = -1;
  if (n_words > (size_t)fes_runs::kMaxTmplWords) return FES_OK;
  for (int s = 0; s < fes_runs::kSlots; ++s) {
    int expect = 0;
    if (!g_slot_used[s].compare_exchange_strong(expect, 1)) continue;
    // kernels of a plan that released this slot may still be running: wait before overwriting it (plan creation
    // synchronises anyway)
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess)
      e = cudaMemcpyToSymbolAsync(c_run_tmpl, words, n_words * sizeof(int32_t), (size_t)s * fes_runs::kMaxTmplWords * sizeof(int32_t),
                                  cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
      g_slot_used[s].store(0);
      return set_error(FES_ERR_CUDA, "template upload failed: %s", cudaGetErrorString(e));
    }
    *slot = s;
    return FES_OK;
  }
  return FES_OK;
}
void runs_slot_release(int slot) {
  if (slot >= 0 && slot < fes_runs::kSlots) g_slot_used[slot].store(0);
}
}  // namespace fes

using namespace fes;

extern "C" int fes_element_matrices(int elem_kind, int form, int n_comp, int spatial_dim, const int32_t* d_elem_nodes,
                                    int64_t n_el, const double* d_coords, const fes_coeffs* coeffs, double* d_Ke,
                                    void* stream) {
  Job j{0, nullptr, d_elem_nodes, n_el, d_coords, {}, d_Ke, (cudaStream_t)stream};
  FES_TRY(check_coeffs(coeffs, j.cf));
  return dispatch(j, elem_kind, form, n_comp, spatial_dim);
}

extern "C" int fes_element_volumes(int elem_kind, int sd, const int32_t* en, int64_t n_el, const double* coords,
                                   double* vol, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el, 256);
#define FES_VOL(EL, SD_) k_volumes<EL, SD_><<<g, 256, 0, st>>>(en, n_el, coords, vol)
  if (elem_kind == FES_P1_LINE && sd == 1) FES_VOL(FES_P1_LINE, 1);
  else if (elem_kind == FES_P1_LINE && sd == 2) FES_VOL(FES_P1_LINE, 2);
  else if (elem_kind == FES_P2_LINE && sd == 2) FES_VOL(FES_P2_LINE, 2);
  else if (elem_kind == FES_P1_TRI && sd == 2) FES_VOL(FES_P1_TRI, 2);
  else if (elem_kind == FES_P1_TRI && sd == 3) FES_VOL(FES_P1_TRI, 3);
  else if (elem_kind == FES_P2_TRI && sd == 2) FES_VOL(FES_P2_TRI, 2);
  else if (elem_kind == FES_P1_TET && sd == 3) FES_VOL(FES_P1_TET, 3);
  else return set_error(FES_ERR_UNSUPPORTED, "no geometry for element kind %d in %dD", elem_kind, sd);
#undef FES_VOL
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_assemble(const fes_plan* plan, int elem_kind, int form, int spatial_dim, const int32_t* d_elem_nodes,
                            const double* d_coords, const fes_coeffs* coeffs, double* d_values, void* stream) {
  if (!plan) return set_error(FES_ERR_VALUE, "fes_assemble: plan is NULL");
  if (elem_nodes_of(elem_kind) != plan->N)
    return set_error(FES_ERR_VALUE, "fes_assemble: plan was built for %d-node elements, element kind %d has %d",
                     plan->N, elem_kind, elem_nodes_of(elem_kind));
  Job j{1, plan, d_elem_nodes, plan->n_el, d_coords, {}, d_values, (cudaStream_t)stream};
  FES_TRY(check_coeffs(coeffs, j.cf));
  return dispatch(j, elem_kind, form, plan->c, spatial_dim);
}

extern "C" int fes_assemble_precomputed(const fes_plan* P, const double* d_Ke, const double* d_scale, double factor,
                                        double* d_values, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!P) return set_error(FES_ERR_VALUE, "fes_assemble_precomputed: plan is NULL");
  if (P->n_pairs == 0) return FES_OK;
  const int g = grid_for(P->n_pairs, kBlock);
#define FES_PRE(C_)                                                                                             \
  k_assemble_precomputed<C_><<<g, kBlock, 0, st>>>(P->n_pairs, P->N, P->pair_start.p, P->pair_row.p, P->node_ptr.p, \
                                                   P->contrib.p, P->indptr.p, d_Ke, d_scale, factor, d_values)
  if (P->c == 1) FES_PRE(1);
  else if (P->c == 2) FES_PRE(2);
  else FES_PRE(3);
#undef FES_PRE
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_assemble_host(const fes_plan* plan, int elem_kind, int form, int spatial_dim,
                                 const int32_t* d_elem_nodes, const double* h_coords, int64_t n_nodes,
                                 const fes_coeffs* coeffs, const double* h_E, const double* h_scale, double* h_values,
                                 void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!plan || !coeffs) return set_error(FES_ERR_VALUE, "fes_assemble_host: NULL argument");
  if (coeffs->d_E || coeffs->d_scale)
    return set_error(FES_ERR_VALUE, "fes_assemble_host: pass per-element coefficients through h_E / h_scale");
  // device staging lives with the plan: repeated assemblies (a SIMP loop, a time stepper
  // rebuilding its operator) pay the copies, not an allocation
  DevBuf<double>&coords = plan->stage_coords, &E = plan->stage_E, &scale = plan->stage_scale,
                &values = plan->stage_values;
  if (coords.n != (size_t)n_nodes * spatial_dim) FES_CUDA(coords.alloc((size_t)n_nodes * spatial_dim));
  if (values.n != (size_t)plan->nnz) FES_CUDA(values.alloc(plan->nnz));
  FES_CUDA(cudaMemcpyAsync(coords.p, h_coords, coords.bytes(), cudaMemcpyHostToDevice, st));
  fes_coeffs c = *coeffs;
  if (h_E) {
    if (E.n != (size_t)plan->n_el) FES_CUDA(E.alloc(plan->n_el));
    FES_CUDA(cudaMemcpyAsync(E.p, h_E, E.bytes(), cudaMemcpyHostToDevice, st));
    c.d_E = E.p;
  }
  if (h_scale) {
    if (scale.n != (size_t)plan->n_el) FES_CUDA(scale.alloc(plan->n_el));
    FES_CUDA(cudaMemcpyAsync(scale.p, h_scale, scale.bytes(), cudaMemcpyHostToDevice, st));
    c.d_scale = scale.p;
  }
  FES_TRY(fes_assemble(plan, elem_kind, form, spatial_dim, d_elem_nodes, coords.p, &c, values.p, st));
  FES_CUDA(cudaMemcpyAsync(h_values, values.p, values.bytes(), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  return FES_OK;
}
