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

#include "elements.cuh"
#include "plan.cuh"
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
  double sign;         // sign of the global factor; the tiled kernel works with |factor|
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


// ---- fused assembly, tiled (the production kernel) ------------------------------------------
// One CTA (128 threads) per tile of consecutive nodes.
//   staging: thread 0 arms an mbarrier and issues two 1-D TMA bulk copies (cp.async.bulk, SASS
//            UBLKCP) of the tile's contiguous plan streams -- the packed metadata of its active
//            pairs and their contribution codes -- which land while phase 1 runs.
//   phase 1: one thread per (node, incident element) record, two records per thread in flight,
//            evaluates the element's geometry ONCE -- inverse Jacobian (the only fp64 divide) and
//            every node's physical gradient at every quadrature point, pre-multiplied by
//            sqrt(w_q mu_e) -- into shared memory, double2-packed and record-minor so both the
//            writes and the LDS.128 reads of phase 2 are as wide as they can be.
//   phase 2: one thread per ACTIVE pair (v, w >= v), taken in descending contribution count so a
//            warp's trip counts agree.  It walks the pair's contributions in ascending element id
//            (np.bincount's order), accumulates the raw outer products S += g~_a (x) g~_b in
//            registers, applies K = (lam/mu) S + S^T + tr(S) I once, and writes the block of (v, w)
//            and -- K being symmetric, bit for bit -- its transpose as the block of (w, v).
// Every CSR value is written exactly once; no atomics, so results are reproducible.
template <int ELEM, int FORM, int C, int SD>
struct RecLayout {
  static constexpr int N = ET<ELEM>::N, NQ = ET<ELEM>::NQ;
  // per (quadrature point, node): SD/2 double2 planes + (SD & 1) double plane; mass: one weight
  static constexpr int NV2 = (FORM == FES_FORM_MASS) ? 0 : NQ * N * (SD / 2);
  static constexpr int NV1 = (FORM == FES_FORM_MASS) ? 1 : NQ * N * (SD & 1);
  static constexpr int DOUBLES = 2 * NV2 + NV1;
};

constexpr int kTileThreads = 128;
constexpr int kRecBatch = 2;  // records a thread evaluates per pass of phase 1 (memory-level parallelism)

// shared-memory carve-up: geometry records and node coordinates (one copy), then two copies of
// the TMA-staged plan streams (double buffering across tiles)
template <int ELEM, int FORM, int C, int SD>
struct TileSmem {
  using R = RecLayout<ELEM, FORM, C, SD>;
  __host__ __device__ static size_t rec_bytes(int gmax) { return ((size_t)gmax * R::DOUBLES * 8 + 15) & ~(size_t)15; }
  __host__ __device__ static size_t xyz_bytes() { return (size_t)256 * SD * 8; }   // distinct-node coordinates
  __host__ __device__ static size_t meta_bytes(int amax) { return (size_t)amax * 16 + 16; }
  __host__ __device__ static size_t code_bytes(int cmax) { return ((size_t)cmax * 2 + 32 + 15) & ~(size_t)15; }
  __host__ __device__ static size_t nod_bytes() { return (size_t)256 * 4; }        // distinct-node ids
  __host__ __device__ static size_t loc_bytes(int gmax) { return ((size_t)gmax * 4 + 32 + 15) & ~(size_t)15; }
  __host__ __device__ static size_t stage_bytes(int gmax, int amax, int cmax) {
    return meta_bytes(amax) + code_bytes(cmax) + nod_bytes() + loc_bytes(gmax);
  }
  __host__ __device__ static size_t total(int gmax, int amax, int cmax) {
    return rec_bytes(gmax) + xyz_bytes() + 2 * stage_bytes(gmax, amax, cmax);
  }
};

template <int ELEM, int FORM, int C, int SD>
__global__ void __launch_bounds__(kTileThreads)
k_assemble_tiles(const int4* __restrict__ tile_hdr, int n_tiles, const uint32_t* __restrict__ ne_code,
                 const uint4* __restrict__ act_meta, const uint16_t* __restrict__ act_codes,
                 const int32_t* __restrict__ tile_nodes, const uint32_t* __restrict__ rec_local,
                 const double* __restrict__ coords, Coeffs cf, int geo_stride, int act_max, int con_max,
                 double* __restrict__ values) {
  using R = RecLayout<ELEM, FORM, C, SD>;
  using TS = TileSmem<ELEM, FORM, C, SD>;
  constexpr int N = ET<ELEM>::N, NQ = ET<ELEM>::NQ, RD = ET<ELEM>::RD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_skew[2][2];
  const int GS = geo_stride;
  // planes: v2[plane * GS + record] (double2), then v1[plane * GS + record] (double)
  double2* v2 = reinterpret_cast<double2*>(smem_raw);
  double* v1 = reinterpret_cast<double*>(smem_raw) + (size_t)2 * R::NV2 * GS;
  double* s_xyz = reinterpret_cast<double*>(smem_raw + TS::rec_bytes(GS));
  unsigned char* stage0 = smem_raw + TS::rec_bytes(GS) + TS::xyz_bytes();
  const size_t stage_sz = TS::stage_bytes(GS, act_max, con_max);

  // One thread arms the stage's mbarrier and issues four 1-D TMA bulk copies of the tile's
  // contiguous plan streams; they land while the previous tile is still being computed.
  auto issue = [&](const int4& h0, const int4& h1, int tile, int buf) {
    unsigned char* st = stage0 + buf * stage_sz;
    unsigned char* s_code = st + TS::meta_bytes(act_max);
    unsigned char* s_nod = s_code + TS::code_bytes(con_max);
    unsigned char* s_loc = s_nod + TS::nod_bytes();
    const uint32_t n_rec = (uint32_t)(h0.y - h0.x);
    const uint32_t b_meta = (uint32_t)h0.w * 16u;
    const uint32_t b_code = (uint32_t)(((((uint64_t)h1.x * 2) & 15) + (uint64_t)h1.y * 2 + 15) & ~(uint64_t)15);
    const uint32_t b_nod = ((uint32_t)h1.z * 4u + 15u) & ~15u;
    const uint32_t b_loc = (uint32_t)(((((uint64_t)h0.x * 4) & 15) + (uint64_t)n_rec * 4 + 15) & ~(uint64_t)15);
    mbar_arrive_expect_tx(&s_bar[buf], b_meta + b_code + b_nod + b_loc);
    uint32_t t;
    if (b_nod) bulk_g2s(s_nod, tile_nodes + (int64_t)tile * fes_plan::kTileNodeStride, b_nod, &s_bar[buf]);
    s_skew[buf][1] = stage_window<4>(s_loc, rec_local, (uint32_t)h0.x, n_rec, &s_bar[buf], t);
    if (b_meta) bulk_g2s(st, act_meta + h0.z, b_meta, &s_bar[buf]);
    s_skew[buf][0] = stage_window<2>(s_code, act_codes, (uint32_t)h1.x, (uint32_t)h1.y, &s_bar[buf], t);
  };

  int tile = blockIdx.x;
  if (tile >= n_tiles) return;
  int4 h0 = __ldg(tile_hdr + 2 * tile), h1 = __ldg(tile_hdr + 2 * tile + 1);
  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    issue(h0, h1, tile, 0);
  }
  __syncthreads();

  for (int it = 0; tile < n_tiles; ++it) {
    const int buf = it & 1;
    const uint32_t parity = (uint32_t)(it >> 1) & 1u;
    const int next = tile + (int)gridDim.x;
    int4 n0 = h0, n1 = h1;
    if (next < n_tiles) {
      n0 = __ldg(tile_hdr + 2 * next);
      n1 = __ldg(tile_hdr + 2 * next + 1);
      if (threadIdx.x == 0) issue(n0, n1, next, buf ^ 1);  // that stage was last read before the barrier below
    }
    const int32_t g0 = h0.x, g1 = h0.y, n_act = h0.w, nd = h1.z;
    unsigned char* st = stage0 + buf * stage_sz;
    const uint4* s_meta = reinterpret_cast<const uint4*>(st);
    unsigned char* s_code = st + TS::meta_bytes(act_max);
    const int32_t* s_nod = reinterpret_cast<const int32_t*>(s_code + TS::code_bytes(con_max));
    unsigned char* s_loc = s_code + TS::code_bytes(con_max) + TS::nod_bytes();
    mbar_wait(&s_bar[buf], parity);  // this tile's plan streams have landed

    // ---- phase 0: coordinates of the tile's distinct nodes, gathered once
    for (int i = threadIdx.x; i < nd; i += kTileThreads) {
      const int32_t node = s_nod[i];
      if constexpr (SD == 2) {
        reinterpret_cast<double2*>(s_xyz)[i] = __ldg(reinterpret_cast<const double2*>(coords) + node);
      } else {
#pragma unroll
        for (int s = 0; s < SD; ++s) s_xyz[i * SD + s] = __ldg(coords + (int64_t)node * SD + s);
      }
    }
    __syncthreads();
    const uint32_t* locs = reinterpret_cast<const uint32_t*>(s_loc + s_skew[buf][1]);

  // ---- phase 1: geometry records (corner coordinates come from shared memory)
  const bool per_elem = cf.d_E != nullptr || cf.d_scale != nullptr;
  for (int32_t base = g0 + threadIdx.x; base < g1; base += kTileThreads * kRecBatch) {
    uint32_t e[kRecBatch];
    bool live[kRecBatch];
    double X[kRecBatch][RD + 1][SD];
#pragma unroll
    for (int r = 0; r < kRecBatch; ++r) {
      const int32_t gi = base + r * kTileThreads;
      live[r] = gi < g1;
      e[r] = (live[r] && per_elem) ? (__ldg(ne_code + gi) >> 3) : 0u;
      const uint32_t loc = live[r] ? locs[gi - g0] : 0u;
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
      Geom<RD, SD> g;
      make_geom<RD, SD>(X[r], g);
      Material m; double w;
      element_coeffs(cf, e[r], m, w);
      const int rec = base + r * kTileThreads - g0;
      const double vol = g.scale * ref_measure(RD);
      if constexpr (FORM == FES_FORM_MASS) {
        v1[rec] = w * vol;
      } else {
        // g~ = sqrt(w_q mu_e) g: weight and modulus ride on the gradients (mu1 is divided out by
        // using lam/mu below), so phase 2 accumulates plain outer products
        double wq = w * vol / (double)NQ;
        if constexpr (FORM == FES_FORM_ELASTIC) wq *= m.mu;
        const double sq = sqrt(wq);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
          for (int n = 0; n < N; ++n) {
            double gn[SD];
            shape_grad<ELEM, SD>(g, q, n, gn);
#pragma unroll
            for (int h = 0; h < SD / 2; ++h)
              v2[((q * N + n) * (SD / 2) + h) * GS + rec] = make_double2(sq * gn[2 * h], sq * gn[2 * h + 1]);
            if constexpr (SD & 1) v1[(q * N + n) * GS + rec] = sq * gn[SD - 1];
          }
      }
    }
  }
  __syncthreads();        // records visible

  // ---- phase 2: one thread per active pair
  const uint16_t* codes = reinterpret_cast<const uint16_t*>(s_code + s_skew[buf][0]);
  for (int t = threadIdx.x; t < n_act; t += kTileThreads) {
    const uint4 mt = s_meta[t];
    const uint32_t k0 = mt.x & 0xffffu, cnt = (mt.x >> 16) & 0xffu;
    const bool diag = (mt.x >> 24) & 1u;
    double acc[C][C];
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) acc[i][j] = 0.0;
    for (uint32_t k = k0; k < k0 + cnt; ++k) {
      const uint32_t code = codes[k];
      const int a = (code >> 3) & 7, b = code & 7;
      const int rec = code >> 6;
      if constexpr (FORM == FES_FORM_MASS) {
        acc[0][0] = fma(v1[rec], mass_ref<ELEM>(a, b), acc[0][0]);
      } else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          double A_[SD], B_[SD];
#pragma unroll
          for (int h = 0; h < SD / 2; ++h) {
            const double2 t2 = v2[((q * N + a) * (SD / 2) + h) * GS + rec];
            A_[2 * h] = t2.x; A_[2 * h + 1] = t2.y;
          }
          if constexpr (SD & 1) A_[SD - 1] = v1[(q * N + a) * GS + rec];
          if (a == b) {
#pragma unroll
            for (int s = 0; s < SD; ++s) B_[s] = A_[s];
          } else {
#pragma unroll
            for (int h = 0; h < SD / 2; ++h) {
              const double2 t2 = v2[((q * N + b) * (SD / 2) + h) * GS + rec];
              B_[2 * h] = t2.x; B_[2 * h + 1] = t2.y;
            }
            if constexpr (SD & 1) B_[SD - 1] = v1[(q * N + b) * GS + rec];
          }
          if constexpr (FORM == FES_FORM_LAPLACIAN) {
#pragma unroll
            for (int s = 0; s < SD; ++s) acc[0][0] = fma(A_[s], B_[s], acc[0][0]);
          } else {
#pragma unroll
            for (int i = 0; i < C; ++i)
#pragma unroll
              for (int j = 0; j < C; ++j) acc[i][j] = fma(A_[i], B_[j], acc[i][j]);  // S += g~_a (x) g~_b
          }
        }
      }
    }
    double K[C][C];
    if constexpr (FORM == FES_FORM_ELASTIC) {
      double tr = 0.0;
#pragma unroll
      for (int i = 0; i < C; ++i) tr += acc[i][i];
#pragma unroll
      for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j)
          K[i][j] = cf.sign * (fma(cf.lam_over_mu, acc[i][j], acc[j][i]) + (i == j ? tr : 0.0));
    } else if constexpr (FORM == FES_FORM_LAPLACIAN) {
      K[0][0] = cf.sign * acc[0][0];
    } else {
#pragma unroll
      for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) K[i][j] = (i == j) ? acc[0][0] : 0.0;
    }
    // row i of a block starts at C * (q0 + i * deg); the block of (w, v) is K^T
    const int64_t q_own = mt.y, q_mir = mt.z;
    const int64_t deg_own = mt.w & 0xffffu, deg_mir = mt.w >> 16;
#pragma unroll
    for (int i = 0; i < C; ++i) {
      double* row = values + (int64_t)C * (q_own + i * deg_own);
      if constexpr (C == 2) {
        *reinterpret_cast<double2*>(row) = make_double2(K[i][0], K[i][1]);
      } else {
#pragma unroll
        for (int j = 0; j < C; ++j) row[j] = K[i][j];
      }
    }
    if (!diag) {
#pragma unroll
      for (int i = 0; i < C; ++i) {
        double* row = values + (int64_t)C * (q_mir + i * deg_mir);
        if constexpr (C == 2) {
          *reinterpret_cast<double2*>(row) = make_double2(K[0][i], K[1][i]);
        } else {
#pragma unroll
          for (int j = 0; j < C; ++j) row[j] = K[j][i];
        }
      }
    }
  }
    __syncthreads();  // records, coordinates and this stage are free for reuse
    tile = next;
    h0 = n0;
    h1 = n1;
  }
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
    if (P->tiled && !j.untiled) {
      const size_t smem = TileSmem<ELEM, FORM, C, SD>::total(P->geo_max, P->act_max, P->con_max);
      static size_t configured = 0;  // per template instantiation: largest opt-in so far
      if (smem + 1024 > 48 * 1024 && smem > configured) {  // static smem counts toward the 48 KB default
        FES_CUDA(cudaFuncSetAttribute(k_assemble_tiles<ELEM, FORM, C, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        configured = smem;
      }
      Coeffs cf = j.cf;
      if (FORM != FES_FORM_MASS && cf.factor < 0.0) { cf.factor = -cf.factor; cf.sign = -1.0; }
      static int per_sm = 0, sms = 0;  // per template instantiation
      if (per_sm == 0 || smem > configured) {
        int dev = 0;
        FES_CUDA(cudaGetDevice(&dev));
        FES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_tiles<ELEM, FORM, C, SD>,
                                                               kTileThreads, smem));
        if (per_sm < 1) return set_error(FES_ERR_CUDA, "assembly kernel does not fit on an SM (%zu B smem)", smem);
      }
      const int64_t grid = std::min<int64_t>(P->n_tiles, (int64_t)sms * per_sm);  // persistent CTAs
      k_assemble_tiles<ELEM, FORM, C, SD><<<(unsigned)grid, kTileThreads, smem, j.stream>>>(
          reinterpret_cast<const int4*>(P->tile_hdr.p), (int)P->n_tiles, P->ne_code.p, P->act_meta.p, P->act_codes.p,
          P->tile_nodes.p, P->rec_local.p, j.coords, cf, P->geo_max, P->act_max, P->con_max, j.out);
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
