// k_assemble_runs: fused assembly of stencil-run tiles (layout and idea: runs_template.hpp; the plan side is
// plan_runs.cu).  Included by assemble.cu inside namespace fes { namespace {, after Coeffs / element_coeffs.
//
// A tile is T <= 32 lane-steps of a run; lane t of every warp owns lane-step t (the P nodes v0 + step t + phi;
// step = P for a run of consecutive nodes, the node stride of a strided run).
// The template programs of a plan live in a slot of __constant__ memory (uploaded when the plan is built): every
// descriptor and program word is a warp-uniform constant-bank load, not a shared-memory or L1 access.
//   staging : the coordinate spans of the NEXT tile (strided node sequences -> consecutive slots) and, when the launch
//             carries them, the per-element coefficients of its element lines, global -> shared memory with
//             cp.async, issued after phase 1 so they land under phase 2
//   phase 1 : warp per element line, lane = entry: cofactors, one rsqrt, the N gradients scaled by sqrt(w mu vol)
//             -- the arithmetic of k_assemble_tiles, to the bit -- into planes indexed by the entry
//   phase 2 : warp w walks its share of the period's pairs; a contribution is two gradient loads at
//             (code + t): unit stride across the warp, conflict-free, no predicates (short partial sums are
//             padded with the zero entries; on tets a padded tail is skipped by one uniform test).  The partial sums of L contributions are combined in the balanced
//             tree the tiled kernel's xor shuffles form, so a pair gets the same bits from either kernel.  The
//             block goes into a padded shared-memory image (odd row stride: the lanes' stores hit distinct banks)
//   store   : TMA bulk stores (cp.async.bulk.global.shared::cta): one per tile when the image needs no padding,
//             one per image row otherwise; they drain under the next tile's phase 1
// Two barriers per tile.  Nothing but the 32-byte tile header, the coordinates and the per-element coefficients
// comes from HBM.

struct RunDims {
  int rec_entries, coord_nodes, img_units;
  int lines;  // element lines per tile, upper bound: sizes the per-element coefficient tables (allocated only when used)
};

// Template slots: a plan with runs owns one from fes_plan_create to fes_plan_destroy; a plan that finds none free
// keeps the tiled kernel for every node.  (A by-value kernel parameter of this size made every launch ~15 us slower.)
__constant__ int32_t c_run_tmpl[fes_runs::kSlots][fes_runs::kMaxTmplWords];

template <int ELEM, int FORM, int C, int SD>
constexpr bool runs_supported() {
  return ((ELEM == FES_P1_TRI && SD == 2) || (ELEM == FES_P1_TET && SD == 3)) &&
         (FORM == FES_FORM_LAPLACIAN || FORM == FES_FORM_ELASTIC);
}

template <int C, int SD>
struct RunSmem {
  static constexpr int UB = (C % 2 == 0) ? 16 : 8;  // bytes per image unit
  __host__ __device__ static size_t xyz_bytes(int nodes) { return up16((size_t)nodes * SD * 8); }
  __host__ __device__ static size_t v2_bytes(int entries) { return up16((size_t)entries * 16); }
  __host__ __device__ static size_t v1_bytes(int entries) { return (SD & 1) ? up16((size_t)entries * 8) : 0; }
  __host__ __device__ static size_t img_bytes(int units) { return up16((size_t)units * UB + 16); }
  __host__ __device__ static size_t coef_bytes(int lines) { return up16((size_t)lines * 33 * 8); }  // one table per coefficient array
  __host__ __device__ static size_t total(const RunDims& d, int n_coef) {
    return xyz_bytes(d.coord_nodes) + v2_bytes(d.rec_entries) + v1_bytes(d.rec_entries) + img_bytes(d.img_units) +
           (size_t)n_coef * coef_bytes(d.lines);
  }
};


__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Gradient loads of phase 2 by 32-bit shared address.  FES_RUN_PLAIN_LDS=1 (A-B build): ordinary C++ loads through a
// shared-memory pointer instead of volatile asm, so the compiler may hoist the loads of the next partial sum over the
// FMAs of the current one (it still orders them against the barriers).
#ifndef FES_RUN_PLAIN_LDS
#define FES_RUN_PLAIN_LDS 0
#endif
#ifndef FES_RUN_PAIR_UNROLL
#define FES_RUN_PAIR_UNROLL 1  // pairs of phase 2 unrolled together
#endif
#ifndef FES_RUN_SHORT_PARTIALS
#define FES_RUN_SHORT_PARTIALS 1
#endif
__device__ __forceinline__ double2 run_lds2(uint32_t addr) {
#if FES_RUN_PLAIN_LDS
  extern __shared__ __align__(16) unsigned char smem_raw[];
  return *reinterpret_cast<const double2*>(smem_raw + (addr - smem_u32(smem_raw)));
#else
  return lds_f64x2(addr);
#endif
}
__device__ __forceinline__ double run_lds1(uint32_t addr) {
#if FES_RUN_PLAIN_LDS
  extern __shared__ __align__(16) unsigned char smem_raw[];
  return *reinterpret_cast<const double*>(smem_raw + (addr - smem_u32(smem_raw)));
#else
  return lds_f64(addr);
#endif
}

// One partial sum of a pair: exactly L contributions, codes at program words [cw, cw + L); every load is issued
// before the FMAs.  kDiag: the pair (v, v), both gradients of a contribution are the same vector.
template <int FORM, int C, int SD, int L, bool kDiag, int NACC>
__device__ __forceinline__ void run_partial_m(const int32_t* __restrict__ tmw, int cw, uint32_t lane2, uint32_t lane1,
                                              double (&acc)[NACC]) {
  double2 ta[L], tb[L];
  [[maybe_unused]] double za[L], zb[L];
#pragma unroll
  for (int x = 0; x < L; ++x) {
    const uint32_t code = (uint32_t)tmw[cw + x];
    const uint32_t oa = code & 0xffffu;
    ta[x] = run_lds2(lane2 + oa);
    if constexpr (SD & 1) za[x] = run_lds1(lane1 + (oa >> 1));
    if constexpr (!kDiag) {
      const uint32_t ob = code >> 16;
      tb[x] = run_lds2(lane2 + ob);
      if constexpr (SD & 1) zb[x] = run_lds1(lane1 + (ob >> 1));
    }
  }
#pragma unroll
  for (int x = 0; x < NACC; ++x) acc[x] = 0.0;
#pragma unroll
  for (int x = 0; x < L; ++x) {
    double A_[SD], B_[SD];
    A_[0] = ta[x].x; A_[1] = ta[x].y;
    if constexpr (SD & 1) A_[SD - 1] = za[x];
    if constexpr (kDiag) {
#pragma unroll
      for (int s = 0; s < SD; ++s) B_[s] = A_[s];
    } else {
      B_[0] = tb[x].x; B_[1] = tb[x].y;
      if constexpr (SD & 1) B_[SD - 1] = zb[x];
    }
    if constexpr (FORM == FES_FORM_LAPLACIAN) {
#pragma unroll
      for (int s = 0; s < SD; ++s) acc[0] = fma(A_[s], B_[s], acc[0]);
    } else {
#pragma unroll
      for (int a_ = 0; a_ < C; ++a_)
#pragma unroll
        for (int b_ = 0; b_ < C; ++b_) acc[a_ * C + b_] = fma(A_[a_], B_[b_], acc[a_ * C + b_]);
    }
  }
}

// A partial sum = exactly L codes, short ones padded with code 0 (the zero entries) at the tail.  On tets (L = 6) six
// of a node's fourteen off-diagonal pairs have 4 contributions: when the fifth code is padding only four are
// executed (one warp-uniform test per partial sum; fma(0, 0, acc) == acc, so the bits do not change).
template <int FORM, int C, int SD, int N, bool kDiag, int NACC>
__device__ __forceinline__ void run_partial(const int32_t* __restrict__ tmw, int cw, uint32_t lane2, uint32_t lane1,
                                            double (&acc)[NACC]) {
  constexpr int L = item_len_of(N);
  if constexpr (L == 6 && FES_RUN_SHORT_PARTIALS) {
    if (tmw[cw + 4] == 0) { run_partial_m<FORM, C, SD, 4, kDiag>(tmw, cw, lane2, lane1, acc); return; }
  }
  run_partial_m<FORM, C, SD, L, kDiag>(tmw, cw, lane2, lane1, acc);
}

template <int NACC>
__device__ __forceinline__ void run_add(double (&a)[NACC], const double (&b)[NACC]) {
#pragma unroll
  for (int x = 0; x < NACC; ++x) a[x] = a[x] + b[x];
}

// Balanced tree over the 2^lg partial sums of a pair: the order of the tiled kernel's xor-shuffle tree.
template <int FORM, int C, int SD, int N, bool kDiag, int NACC>
__device__ __forceinline__ void run_pair_sum(const int32_t* __restrict__ tmw, int cw, uint32_t lg, uint32_t lane2,
                                             uint32_t lane1, double (&acc)[NACC]) {
  constexpr int L = item_len_of(N);
  run_partial<FORM, C, SD, N, kDiag>(tmw, cw, lane2, lane1, acc);
  if (lg >= 1) {
    double b[NACC];
    run_partial<FORM, C, SD, N, kDiag>(tmw, cw + L, lane2, lane1, b);
    run_add(acc, b);
    if (lg >= 2) {
      double c_[NACC];
      run_partial<FORM, C, SD, N, kDiag>(tmw, cw + 2 * L, lane2, lane1, c_);
      run_partial<FORM, C, SD, N, kDiag>(tmw, cw + 3 * L, lane2, lane1, b);
      run_add(c_, b);
      if constexpr (N == 3) {
        if (lg >= 3) {  // ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7))
          run_add(acc, c_);
          double d_[NACC];
          run_partial<FORM, C, SD, N, kDiag>(tmw, cw + 4 * L, lane2, lane1, c_);
          run_partial<FORM, C, SD, N, kDiag>(tmw, cw + 5 * L, lane2, lane1, b);
          run_add(c_, b);
          run_partial<FORM, C, SD, N, kDiag>(tmw, cw + 6 * L, lane2, lane1, d_);
          run_partial<FORM, C, SD, N, kDiag>(tmw, cw + 7 * L, lane2, lane1, b);
          run_add(d_, b);
          run_add(c_, d_);
        }
      }
      run_add(acc, c_);
    }
  }
}

// resident CTAs per SM the register allocation aims at: shared memory allows 2 (tets) / 7 (triangles)
#ifndef FES_RUN_KB2D
#define FES_RUN_KB2D 1  // records per thread in flight in phase 1, triangles
#endif
#ifndef FES_RUN_KB3D
#define FES_RUN_KB3D 1  // ... tets (measured at 151^3: 1.396 ms with 1, 1.597 with 2, 1.590 with 3)
#endif
#ifndef FES_RUN_MINCTA2D
#define FES_RUN_MINCTA2D 7
#endif
template <int SD, int kRunThreads>
constexpr int run_min_ctas() { return SD == 3 ? 2 : (kRunThreads >= 256 ? 3 : (kRunThreads >= 192 ? 4 : FES_RUN_MINCTA2D)); }

template <int ELEM, int FORM, int C, int SD, int kRunThreads>
__global__ void __launch_bounds__(kRunThreads, run_min_ctas<SD, kRunThreads>())
k_assemble_runs(const int32_t* __restrict__ tiles, int n_tiles, int slot, const double* __restrict__ coords,
                int64_t n_nodes, int64_t n_el, Coeffs cf, RunDims dims, int use_bulk, double* __restrict__ values) {
  using RS = RunSmem<C, SD>;
  constexpr int N = ET<ELEM>::N, RD = ET<ELEM>::RD;
  static_assert(ET<ELEM>::NQ == 1 && RD == SD, "run tiles: P1 volume elements");
  constexpr int L = item_len_of(N);
  constexpr int UB = RS::UB;
  constexpr int NACC = (FORM == FES_FORM_ELASTIC) ? C * C : 1;
  constexpr int kZero = fes_runs_host::kZeroEntries;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  struct TmView { const int32_t* w; };
  const TmView tm{c_run_tmpl[slot]};
  unsigned char* sp = smem_raw;
  double* s_xyz = reinterpret_cast<double*>(sp);    sp += RS::xyz_bytes(dims.coord_nodes);
  double2* v2 = reinterpret_cast<double2*>(sp);     sp += RS::v2_bytes(dims.rec_entries);
  double* v1 = reinterpret_cast<double*>(sp);       sp += RS::v1_bytes(dims.rec_entries);
  unsigned char* s_img = sp;                        sp += RS::img_bytes(dims.img_units);
  // per-element coefficients of the tile's element lines, [line][entry], staged with the coordinates (the SIMP
  // reassembly: rho^p as d_scale; a modulus field as d_E); present only when the launch carries such arrays
  double* s_scale = reinterpret_cast<double*>(sp);  sp += cf.d_scale ? RS::coef_bytes(dims.lines) : 0;
  double* s_E = reinterpret_cast<double*>(sp);
  const uint32_t v2_b = smem_u32(v2), v1_b = smem_u32(v1), img_b = smem_u32(s_img);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool per_elem = cf.d_E != nullptr || cf.d_scale != nullptr;
  double sq_uniform;
  {
    double c2 = cf.factor * (ref_measure(RD) / 1.0);
    if constexpr (FORM == FES_FORM_ELASTIC) c2 *= cf.E * cf.mu1;
    sq_uniform = sqrt(c2);
  }
  // the zero entries that pad short partial sums
  if (tid < kZero) {
    v2[tid] = make_double2(0.0, 0.0);
    if constexpr (SD & 1) v1[tid] = 0.0;
  }
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (n_my <= 0) return;
  // Tile headers {v0, T, template offset, E0 | first pair, -, pair stride, -} travel global -> shared memory with
  // cp.async two tiles ahead, into a ring of four 32-byte slots, and are read from shared memory when their tile comes
  // up: no thread ever waits for a header.  (Loaded into registers -- even through volatile asm -- the compiler moves
  // them to uniform registers right behind the load: a long-scoreboard stall per tile, 12 % of the samples on
  // config 2.)  Completion rides on the cp.async wait + barrier the coordinate staging already has.
  __shared__ __align__(16) int4 s_hdr[4][2];
  const int4* hdr4 = reinterpret_cast<const int4*>(tiles);
  auto fetch_hdr = [&](int k) {
    if (tid < 2) cp_async16(&s_hdr[k & 3][tid], hdr4 + 2 * ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) + tid);
  };
  fetch_hdr(0);
  if (n_my > 1) fetch_hdr(1);
  cp_async_wait_all();
  __syncthreads();

  // staging of one tile's coordinate spans (strided node sequences v0 + start + step * i -> slots base + i): warp w
  // takes spans w, w + kWarps, ...
  auto stage = [&](int32_t v0, int32_t T, int32_t toff, int32_t E0) {
    const int step = tm.w[toff + 9], n_spans = tm.w[toff + 5], so = toff + tm.w[toff + 7];
    if (per_elem) {
      // Element of entry tau' of line l = E0 + eoff_l + tau' * delta.  Consecutive THREADS take consecutive LINES of
      // one entry: the lines are sorted by eoff and the elements of a cell column are numbered consecutively, so a
      // quarter-warp's 8-byte reads fall into one or two sectors (thread = entry would touch 32 sectors per
      // instruction: measured +45 % on config 2 with a rho^p scale).  Tables are [line][33]: the odd pitch keeps the
      // scattered 8-byte writes of one entry on distinct banks.
      const int n_lines = tm.w[toff + 1], lo = toff + tm.w[toff + 6];
      const int64_t delta = tm.w[toff + 14];
      const uint32_t inv = ((1u << 20) + (uint32_t)n_lines - 1u) / (uint32_t)n_lines;
      for (int i = tid; i < n_lines * 32; i += kRunThreads) {
        const int tp = (int)(((uint32_t)i * inv) >> 20), l = i - tp * n_lines;
        const int ld = lo + l * (2 + N);
        if (tp < T + tm.w[ld + 1]) {
          int64_t e = (int64_t)E0 + tm.w[ld] + (int64_t)tp * delta;
          e = e < 0 ? 0 : (e >= n_el ? n_el - 1 : e);  // entries no lane references may fall outside the mesh
          if (cf.d_scale) cp_async8(s_scale + l * 33 + tp, cf.d_scale + e);
          if (cf.d_E) cp_async8(s_E + l * 33 + tp, cf.d_E + e);
        }
      }
    }
    for (int s = warp; s < n_spans; s += kRunThreads / 32) {
      const int32_t start = tm.w[so + 4 * s], n_ext = tm.w[so + 4 * s + 1], base = tm.w[so + 4 * s + 2];
      const int cnt = T + n_ext;
      for (int i = lane; i < cnt; i += 32) {
        const int64_t node = (int64_t)v0 + start + (int64_t)step * i;
        if (node < 0 || node >= n_nodes) continue;  // entries no lane references may fall outside the mesh
        if constexpr (SD == 2) {
          cp_async16(reinterpret_cast<double2*>(s_xyz) + base + i, reinterpret_cast<const double2*>(coords) + node);
        } else {
#pragma unroll
          for (int comp = 0; comp < SD; ++comp) cp_async8(s_xyz + (base + i) * SD + comp, coords + node * SD + comp);
        }
      }
    }
  };
  {
    const int4 h0 = s_hdr[0][0];
    stage(h0.x, h0.y, h0.z, h0.w);
  }
  cp_async_wait_all();
  __syncthreads();

  for (int k = 0; k < n_my; ++k) {
    const int4 h0 = s_hdr[k & 3][0], g0 = s_hdr[k & 3][1];
    const int32_t T = h0.y, toff = h0.z;  // (E0 = h0.w is consumed by the staging of the tile, one iteration earlier)
    const int32_t my_pair0 = g0.x, pstride = g0.z;
    const bool has_next = k + 1 < n_my;
    if (k + 2 < n_my) fetch_hdr(k + 2);  // its slot held tile k - 2, last read two barriers ago
    const int n_lines = tm.w[toff + 1], psh = tm.w[toff + 3], LU = tm.w[toff + 4];
    // image row r starts at unit r LU + r / g, g = 2^(psh - 1) rows per pad unit (psh = 0: no padding)
    auto row_off = [&](int r) { return (uint32_t)(r * LU + (psh ? (r >> (psh - 1)) : 0)); };
    // 8-byte units and an odd row length: the image needs no padding (psh == 0), i.e. it IS the tile's contiguous
    // slice of `values` and one TMA bulk store writes it.  An odd first value index is mirrored by an 8-byte skew of
    // the image, so both sides of the copy share their 16-byte phase.
    // A strided run's rows are pstride pairs apart in `values` (contiguous runs: pstride = the pairs per lane-step)
    const int GR = C * C * pstride / (UB / 8);  // units between consecutive image rows in `values`
    const bool bulk = UB == 8 && psh == 0 && GR == LU && use_bulk;
    const bool row_bulk = UB == 16 && (use_bulk & 2);
    const uint32_t skew = bulk ? (uint32_t)(((int64_t)C * C * my_pair0) & 1) : 0u;

    // ---- phase 1: one geometry record per (line, entry).  A warp takes whole lines (lane = entry; a line has at
    // most 32 entries), kB of them at a time: the line descriptors are warp-uniform constant loads, the coordinate
    // reads and record writes unit-stride, and the long fp64 rsqrt chains of the kB records overlap
    {
      const int lo = toff + tm.w[toff + 6];
      constexpr int kB = (SD == 3) ? FES_RUN_KB3D : FES_RUN_KB2D;
      constexpr int kWarps = kRunThreads / 32;
      for (int l0 = warp; l0 < n_lines; l0 += kB * kWarps) {
        double X[kB][RD + 1][SD];
        bool live[kB];
#pragma unroll
        for (int r = 0; r < kB; ++r) {
          const int l = l0 + r * kWarps;
          const int ld = lo + l * (2 + N);
          live[r] = l < n_lines && lane < T + tm.w[ld + 1];
          if (live[r]) {
#pragma unroll
            for (int n = 0; n <= RD; ++n) {
              const int ci = tm.w[ld + 2 + n] + lane;
              if constexpr (SD == 2) {
                const double2 xy = reinterpret_cast<const double2*>(s_xyz)[ci];
                X[r][n][0] = xy.x; X[r][n][1] = xy.y;
              } else {
#pragma unroll
                for (int s = 0; s < SD; ++s) X[r][n][s] = s_xyz[ci * SD + s];
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < kB; ++r) {
          if (!live[r]) continue;
          const int l = l0 + r * kWarps;
          double cof[RD][SD], det;
          make_cof<RD, SD>(X[r], cof, det);
          double f = sq_uniform, d = fabs(det);
          if (per_elem) {  // element_coeffs' arithmetic on the staged values
            const double E = cf.d_E ? s_E[l * 33 + lane] : cf.E;
            const double w = cf.d_scale ? cf.factor * s_scale[l * 33 + lane] : cf.factor;
            double c2 = w * (ref_measure(RD) / 1.0);
            if constexpr (FORM == FES_FORM_ELASTIC) c2 *= E * cf.mu1;
            // sqrt(c2) / sqrt(|det|) = c2 rsqrt(c2 |det|): one rsqrt per record either way (f * rsqrt(d) below); a void
            // element (weight 0) contributes exact zeros
            f = c2 > 0.0 ? c2 : 0.0;
            d = c2 > 0.0 ? c2 * d : 1.0;
          }
          f = copysign(f * rsqrt(d), det);
#pragma unroll
          for (int i2 = 0; i2 < RD; ++i2)
#pragma unroll
            for (int s = 0; s < SD; ++s) cof[i2][s] *= f;
#pragma unroll
          for (int n = 0; n < N; ++n) {
            double gn[SD];
            shape_grad_rows<ELEM, SD>(cof, 0, n, gn);
            const int ent = kZero + (l * N + n) * 32 + lane;
            v2[ent] = make_double2(gn[0], gn[1]);
            if constexpr (SD & 1) v1[ent] = gn[SD - 1];
          }
        }
      }
    }
    bulk_wait_read_all();  // the previous tile's bulk stores (issued by this thread, if any) have finished reading the image
    __syncthreads();  // records visible; the coordinate table is free; the previous image has been copied out

    if (has_next) {  // lands under phase 2
      const int4 h1 = s_hdr[(k + 1) & 3][0];
      stage(h1.x, h1.y, h1.z, h1.w);
    }

    // ---- phase 2: warp = pair group, lane = lane-step
    if (lane < T) {
      const uint32_t lane2 = v2_b + 16u * (uint32_t)lane, lane1 = v1_b + 8u * (uint32_t)lane;
      const uint32_t img_lane = img_b + 8u * skew + row_off(lane) * (uint32_t)UB;
      const int pb = toff + tm.w[toff + 20 + warp];
      const int np = tm.w[pb];
      int w = pb + 1;
      constexpr int kPairUnroll = FES_RUN_PAIR_UNROLL;
#pragma unroll kPairUnroll
      for (int pi = 0; pi < np; ++pi) {
        const uint32_t head = (uint32_t)tm.w[w];
        const uint32_t lg = (head >> 27) & 3u;
        const bool diag = (head >> 29) & 1u;
        const int cw = w + 1;
        w += 1 + (L << lg);
        double acc[NACC];
        if (diag) run_pair_sum<FORM, C, SD, N, true>(tm.w, cw, lg, lane2, lane1, acc);
        else run_pair_sum<FORM, C, SD, N, false>(tm.w, cw, lg, lane2, lane1, acc);
        double K[C][C];
        if constexpr (FORM == FES_FORM_ELASTIC) {
          double tr = 0.0;
#pragma unroll
          for (int a_ = 0; a_ < C; ++a_) tr += acc[a_ * C + a_];
#pragma unroll
          for (int a_ = 0; a_ < C; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < C; ++b_)
              K[a_][b_] = fma(cf.lam_over_mu, acc[a_ * C + b_], acc[b_ * C + a_]) + (a_ == b_ ? tr : 0.0);
        } else {
          K[0][0] = acc[0];
        }
        const uint32_t r0 = img_lane + (head & 0x7fffu), row_b = (head >> 15) & 0xfffu;
        if constexpr (C == 2) {
          sts_f64x2(r0, K[0][0], K[0][1]);
          sts_f64x2(r0 + row_b, K[1][0], K[1][1]);
        } else {
#pragma unroll
          for (int a_ = 0; a_ < C; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < C; ++b_) sts_f64(r0 + (uint32_t)a_ * row_b + 8u * (uint32_t)b_, K[a_][b_]);
        }
      }
    }
    cp_async_wait_all();
    if (bulk || row_bulk) fence_proxy_async_smem();  // the image writes (generic proxy) ordered before the bulk store's reads
    __syncthreads();  // image complete; records free; the next tile's coordinates visible

    // ---- store: image rows -> the tile's contiguous slice of `values`
    if (bulk) {
      if (tid == 0) {
        const int64_t val0 = (int64_t)C * C * my_pair0;
        const int n_out = T * LU, head = (int)skew;
        const int body = (n_out - head) & ~1, tail = n_out - head - body;
        const double* img = reinterpret_cast<const double*>(s_img) + skew;
        if (body > 0) bulk_s2g(values + val0 + head, img + head, (uint32_t)body * 8u);
        if (head) values[val0] = img[0];
        if (tail) values[val0 + n_out - 1] = img[n_out - 1];
      }
    } else if (row_bulk) {
      // padded rows, 16-byte units: one bulk store per group of g adjacent rows (the rows between two pad units are
      // contiguous on both sides; g = 1 for a strided run, whose rows are GR units apart in `values`).  Every warp
      // issues its share (UBLKCP takes uniform operands, so a warp's lanes are served one after the other: spread
      // over the warps, none of them arrives late at the next tile's phase 1)
      constexpr int kWarps = kRunThreads / 32;
      const int gsh = psh ? psh - 1 : 0, g = 1 << gsh;
      const int n_groups = (T + g - 1) >> gsh, gpw = (n_groups + kWarps - 1) / kWarps, j = warp * gpw + lane;
      if (lane < gpw && j < n_groups) {
        const int r0 = j << gsh, rows = min(g, T - r0);
        bulk_s2g(reinterpret_cast<unsigned char*>(values + (int64_t)C * C * my_pair0) + (size_t)r0 * GR * UB,
                 s_img + (size_t)row_off(r0) * UB, (uint32_t)(rows * LU * UB));
      }
    } else {  // padded rows: thread x takes units x, x + threads, ...
      unsigned char* dst = reinterpret_cast<unsigned char*>(values + (int64_t)C * C * my_pair0);
      const int total = T * LU, dr = tm.w[toff + 10], du = tm.w[toff + 11];
      int r = tid / LU, u = tid - r * LU;
#pragma unroll 2
      for (int x = tid; x < total; x += kRunThreads) {
        const uint32_t src = img_b + (row_off(r) + (uint32_t)u) * (uint32_t)UB;
        const size_t o = (size_t)(uint32_t)(r * GR + u) * UB;
        if constexpr (UB == 16) {
          const double2 v = lds_f64x2(src);
          *reinterpret_cast<double2*>(dst + o) = v;
        } else {
          *reinterpret_cast<double*>(dst + o) = lds_f64(src);
        }
        r += dr; u += du;
        if (u >= LU) { u -= LU; ++r; }
      }
    }
  }
  bulk_wait_read_all();  // shared memory must outlive the last bulk stores' reads
}

template <int ELEM, int FORM, int C, int SD, int kRunThreads>
int launch_runs_t(const Job& j, cudaStream_t stream) {
  const fes_runs& R = j.plan->runs;
  const RunDims dims{R.max_rec_entries, R.max_coord_nodes, R.max_img_units, R.max_lines};
  static const int exp_coef = getenv("FES_EXP_COEF_SMEM") ? atoi(getenv("FES_EXP_COEF_SMEM")) : 0;  // occupancy experiment
  const size_t smem = RunSmem<C, SD>::total(dims, (j.cf.d_scale ? 1 : 0) + (j.cf.d_E ? 1 : 0) + exp_coef);
  static size_t configured = 0, occ_smem = 0;
  static int per_sm = 0, sms = 0;
  if (per_sm == 0 || smem != occ_smem) {
    if (smem > configured && smem + 1024 > 48 * 1024)
      FES_CUDA(cudaFuncSetAttribute(k_assemble_runs<ELEM, FORM, C, SD, kRunThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = std::max(configured, smem);
    int dev = 0;
    FES_CUDA(cudaGetDevice(&dev));
    FES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_runs<ELEM, FORM, C, SD, kRunThreads>, kRunThreads, smem));
    occ_smem = smem;
    if (per_sm < 1) {
      per_sm = 0;
      return set_error(FES_ERR_CUDA, "run assembly kernel does not fit on an SM (%zu B smem)", smem);
    }
    if (getenv("FES_PLAN_DEBUG")) fprintf(stderr, "[fes assemble runs] threads=%d smem=%zu B CTAs/SM=%d\n", kRunThreads, smem, per_sm);
  }
  Coeffs cf = j.cf;
  if (cf.factor < 0.0) { cf.factor = -cf.factor; cf.sign = -1.0; }
  const int64_t grid = std::min<int64_t>(R.n_tiles, (int64_t)sms * per_sm);
  static const bool no_bulk = getenv("FES_NO_TMA_STORE") != nullptr;  // tuning knob: the copy loop for every image
  // padded images (16-byte units) go out as one 448-byte-class bulk store per image row, issued by the lanes of one
  // warp: no shared-memory reads or store instructions on the LSU (config 2: 0.1348 -> 0.1245 ms); knob: the copy loop
  static const bool row_bulk = getenv("FES_RUN_NO_ROW_BULK") == nullptr;
  k_assemble_runs<ELEM, FORM, C, SD, kRunThreads><<<(unsigned)grid, kRunThreads, smem, stream>>>(
      (const int32_t*)R.tiles.p, (int)R.n_tiles, R.slot, (const double*)j.coords, (int64_t)j.plan->n_nodes,
      (int64_t)j.plan->n_el, cf, dims, no_bulk ? 0 : (row_bulk ? 3 : 1), j.out);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

template <int ELEM, int FORM, int C, int SD>
int launch_runs(const Job& j, cudaStream_t stream) {
  if constexpr (!runs_supported<ELEM, FORM, C, SD>()) {
    return FES_OK;
  } else {
    if (j.plan->runs.n_warps == 8) return launch_runs_t<ELEM, FORM, C, SD, 256>(j, stream);
    if (j.plan->runs.n_warps == 7) return launch_runs_t<ELEM, FORM, C, SD, 224>(j, stream);
    if (j.plan->runs.n_warps == 6) return launch_runs_t<ELEM, FORM, C, SD, 192>(j, stream);
    return launch_runs_t<ELEM, FORM, C, SD, 128>(j, stream);
  }
}
