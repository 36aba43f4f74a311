// k_assemble_runs: fused assembly of stencil-run tiles (layout and idea: runs_template.hpp; the plan side is
// plan_runs.cu).  Included by assemble.cu inside namespace fes { namespace {, after Coeffs / element_coeffs.
//
// A tile is T <= 32 lane-steps of a run; lane t of every warp owns lane-step t (the P nodes v0 + P t + phi).
//   staging : the coordinate spans of the tile, global -> shared memory (contiguous ranges, coalesced)
//   phase 1 : thread per (element line, entry): cofactors, one rsqrt, the N gradients scaled by sqrt(w mu vol)
//             -- the arithmetic of k_assemble_tiles, to the bit -- into planes indexed by the entry
//   phase 2 : warp w walks its share of the period's pairs from the template program (uniform reads); a
//             contribution is two gradient loads at (code + t): unit stride across the warp, conflict-free.
//             Partial sums of L contributions are combined in the balanced tree the tiled kernel's xor
//             shuffles form, so a pair gets the same bits from either kernel.  The block goes into a padded
//             shared-memory image (odd row stride: the lanes' stores hit distinct banks)
//   store   : coalesced copy of the image rows to the tile's contiguous slice of `values`
// Nothing but the 32-byte tile header, the coordinates and the per-element coefficients comes from HBM.

struct RunDims {
  int rec_entries, coord_nodes, img_units, tmpl_words;
};

template <int ELEM, int FORM, int C, int SD>
constexpr bool runs_supported() {
  return ((ELEM == FES_P1_TRI && SD == 2) || (ELEM == FES_P1_TET && SD == 3)) &&
         (FORM == FES_FORM_LAPLACIAN || FORM == FES_FORM_ELASTIC);
}

template <int C, int SD>
struct RunSmem {
  static constexpr int UB = (C % 2 == 0) ? 16 : 8;  // bytes per image unit
  __host__ __device__ static size_t tmpl_bytes(int words) { return up16((size_t)words * 4); }
  __host__ __device__ static size_t xyz_bytes(int nodes) { return up16((size_t)nodes * SD * 8); }
  __host__ __device__ static size_t v2_bytes(int entries) { return up16((size_t)entries * 16); }
  __host__ __device__ static size_t v1_bytes(int entries) { return (SD & 1) ? up16((size_t)entries * 8) : 0; }
  __host__ __device__ static size_t img_bytes(int units) { return up16((size_t)units * UB); }
  __host__ __device__ static size_t total(const RunDims& d) {
    return tmpl_bytes(d.tmpl_words) + xyz_bytes(d.coord_nodes) + v2_bytes(d.rec_entries) + v1_bytes(d.rec_entries) +
           img_bytes(d.img_units);
  }
};

constexpr int kRunThreads = 128;
static_assert(kRunThreads / 32 == fes_runs_host::kWarps, "one pair group per warp");

template <int ELEM, int FORM, int C, int SD>
__global__ void __launch_bounds__(kRunThreads)
k_assemble_runs(const int32_t* __restrict__ tiles, int n_tiles, const int32_t* __restrict__ tmpl,
                const double* __restrict__ coords, int64_t n_nodes, int64_t n_el, Coeffs cf, RunDims dims,
                double* __restrict__ values) {
  using RS = RunSmem<C, SD>;
  constexpr int N = ET<ELEM>::N, RD = ET<ELEM>::RD;
  static_assert(ET<ELEM>::NQ == 1 && RD == SD, "run tiles: P1 volume elements");
  constexpr int L = item_len_of(N);
  constexpr int CH = (SD == 3) ? 3 : 2;  // contributions gathered per chunk (loads issued before their FMAs)
  static_assert(L % CH == 0, "chunks tile a partial sum");
  constexpr int UB = RS::UB;
  constexpr int NACC = (FORM == FES_FORM_ELASTIC) ? C * C : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* s_t = reinterpret_cast<int32_t*>(smem_raw);
  unsigned char* sp = smem_raw + RS::tmpl_bytes(dims.tmpl_words);
  double* s_xyz = reinterpret_cast<double*>(sp);    sp += RS::xyz_bytes(dims.coord_nodes);
  double2* v2 = reinterpret_cast<double2*>(sp);     sp += RS::v2_bytes(dims.rec_entries);
  double* v1 = reinterpret_cast<double*>(sp);       sp += RS::v1_bytes(dims.rec_entries);
  unsigned char* s_img = sp;
  const uint32_t v2_b = smem_u32(v2), v1_b = smem_u32(v1), img_b = smem_u32(s_img);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool per_elem = cf.d_E != nullptr || cf.d_scale != nullptr;
  double sq_uniform;
  {
    double c2 = cf.factor * (ref_measure(RD) / 1.0);
    if constexpr (FORM == FES_FORM_ELASTIC) c2 *= cf.E * cf.mu1;
    sq_uniform = sqrt(c2);
  }
  asm volatile("griddepcontrol.launch_dependents;");
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (n_my <= 0) return;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int cur_tmpl = -1;
  const int4* hdr4 = reinterpret_cast<const int4*>(tiles);
  int4 h0 = __ldg(hdr4 + 2 * (int64_t)blockIdx.x), h1 = __ldg(hdr4 + 2 * (int64_t)blockIdx.x + 1);
  for (int k = 0; k < n_my; ++k) {
    const int32_t v0 = h0.x, T = h0.y, toff = h0.z, E0 = h0.w, pair0 = h1.x;
    if (k + 1 < n_my) {  // the next header flies under this tile
      const int64_t nx = (int64_t)blockIdx.x + (int64_t)(k + 1) * gridDim.x;
      h0 = __ldg(hdr4 + 2 * nx);
      h1 = __ldg(hdr4 + 2 * nx + 1);
    }
    if (toff != cur_tmpl) {  // rare: a CTA mostly stays within one template
      __syncthreads();
      const int nw = __ldg(tmpl + toff + 12);
      for (int i = tid; i < nw; i += kRunThreads) s_t[i] = __ldg(tmpl + toff + i);
      cur_tmpl = toff;
      __syncthreads();
    }
    const int P = s_t[0], n_lines = s_t[1], TS = s_t[2], LUp = s_t[3], LU = s_t[4], n_spans = s_t[5];
    const int off_lines = s_t[6], off_spans = s_t[7];
    const int32_t delta = s_t[14];
    const uint32_t magic = (uint32_t)s_t[15];

    // ---- staging: coordinate spans (contiguous node ranges), granularity = one node (2D) / one double (3D)
    {
      constexpr int G = (SD == 2) ? 1 : SD;  // copy items per node
      int total = 0;
      for (int s = 0; s < n_spans; ++s) total += (P * (T - 1) + s_t[off_spans + 3 * s + 1] + 1) * G;
      for (int i = tid; i < total; i += kRunThreads) {
        int s = 0, j = i;
        for (;;) {
          const int len = (P * (T - 1) + s_t[off_spans + 3 * s + 1] + 1) * G;
          if (j < len) break;
          j -= len;
          ++s;
        }
        const int64_t g = ((int64_t)v0 + s_t[off_spans + 3 * s]) * G + j;  // global item index
        const int d = s_t[off_spans + 3 * s + 2] * G + j;                     // table item index
        if (g >= 0 && g < n_nodes * G) {
          if constexpr (SD == 2) reinterpret_cast<double2*>(s_xyz)[d] = __ldg(reinterpret_cast<const double2*>(coords) + g);
          else s_xyz[d] = __ldg(coords + g);
        }
      }
    }
    __syncthreads();  // coordinates visible; the previous tile's image has been copied out by everyone

    // ---- phase 1: one geometry record per (line, entry)
    {
      const int n_items = n_lines * TS;
      const uint32_t inv_ts = ((1u << 20) + (uint32_t)TS - 1u) / (uint32_t)TS;
      for (int i = tid; i < n_items; i += kRunThreads) {
        const int l = (int)(((uint32_t)i * inv_ts) >> 20), tp = i - l * TS;
        const int32_t* ld = s_t + off_lines + l * (2 + N);
        if (tp >= T + ld[1]) continue;
        double X[RD + 1][SD];
#pragma unroll
        for (int n = 0; n <= RD; ++n) {
          const int ci = ld[2 + n] + P * tp;
          if constexpr (SD == 2) {
            const double2 xy = reinterpret_cast<const double2*>(s_xyz)[ci];
            X[n][0] = xy.x; X[n][1] = xy.y;
          } else {
#pragma unroll
            for (int s = 0; s < SD; ++s) X[n][s] = s_xyz[ci * SD + s];
          }
        }
        double cof[RD][SD], det;
        make_cof<RD, SD>(X, cof, det);
        double f = sq_uniform;
        if (per_elem) {
          int64_t e = (int64_t)E0 + ld[0] + (int64_t)tp * delta;
          e = e < 0 ? 0 : (e >= n_el ? n_el - 1 : e);  // entries no lane references may fall outside the mesh
          Material m; double w;
          element_coeffs(cf, e, m, w);
          double c2 = w * (ref_measure(RD) / 1.0);
          if constexpr (FORM == FES_FORM_ELASTIC) c2 *= m.mu;
          f = c2 > 0.0 ? c2 * rsqrt(c2) : 0.0;
        }
        f = copysign(f * rsqrt(fabs(det)), det);
#pragma unroll
        for (int i2 = 0; i2 < RD; ++i2)
#pragma unroll
          for (int s = 0; s < SD; ++s) cof[i2][s] *= f;
#pragma unroll
        for (int n = 0; n < N; ++n) {
          double gn[SD];
          shape_grad_rows<ELEM, SD>(cof, 0, n, gn);
          const int ent = (l * N + n) * TS + tp;
          v2[ent] = make_double2(gn[0], gn[1]);
          if constexpr (SD & 1) v1[ent] = gn[SD - 1];
        }
      }
    }
    __syncthreads();  // records visible

    // ---- phase 2: warp = pair group, lane = lane-step
    {
      const int32_t* prog = s_t + s_t[8 + warp];
      const int np = prog[0];
      const bool active = lane < T;
      const uint32_t lane2 = 16u * (uint32_t)lane, lane1 = 8u * (uint32_t)lane;
      const uint32_t img_lane = img_b + (uint32_t)lane * (uint32_t)LUp * (uint32_t)UB;
      int w = 1;
      for (int pi = 0; pi < np; ++pi) {
        const uint32_t head = (uint32_t)prog[w];
        const uint32_t img = head & 0xfffu, deg = (head >> 12) & 127u, cnt = (head >> 19) & 63u, lg = (head >> 25) & 7u;
        const uint32_t m = (cnt + (1u << lg) - 1u) >> lg;
        double lvl[3][NACC];
        double acc[C][C];
        for (uint32_t j = 0; j < (1u << lg); ++j) {
#pragma unroll
          for (int a_ = 0; a_ < C; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < C; ++b_) acc[a_][b_] = 0.0;
#pragma unroll
          for (int x0 = 0; x0 < L; x0 += CH) {
            if ((uint32_t)x0 >= m || j * m + (uint32_t)x0 >= cnt) break;  // uniform
            uint32_t code[CH];
            bool on[CH];
#pragma unroll
            for (int x = 0; x < CH; ++x) {
              const uint32_t kk = j * m + (uint32_t)(x0 + x);
              on[x] = (uint32_t)(x0 + x) < m && kk < cnt;
              code[x] = on[x] ? (uint32_t)prog[w + 1 + (int)kk] : 0u;
            }
            if (active) {
              double2 ta[CH], tb[CH];
              [[maybe_unused]] double za[CH], zb[CH];
#pragma unroll
              for (int x = 0; x < CH; ++x) {
                if (!on[x]) continue;
                const uint32_t ea = code[x] & 0xffffu, eb = code[x] >> 16;
                ta[x] = lds_f64x2(v2_b + 16u * ea + lane2);
                if constexpr (SD & 1) za[x] = lds_f64(v1_b + 8u * ea + lane1);
                if (ea != eb) {
                  tb[x] = lds_f64x2(v2_b + 16u * eb + lane2);
                  if constexpr (SD & 1) zb[x] = lds_f64(v1_b + 8u * eb + lane1);
                } else {
                  tb[x] = ta[x];
                  if constexpr (SD & 1) zb[x] = za[x];
                }
              }
#pragma unroll
              for (int x = 0; x < CH; ++x) {
                if (!on[x]) continue;
                double A_[SD], B_[SD];
                A_[0] = ta[x].x; A_[1] = ta[x].y;
                B_[0] = tb[x].x; B_[1] = tb[x].y;
                if constexpr (SD & 1) { A_[SD - 1] = za[x]; B_[SD - 1] = zb[x]; }
                if constexpr (FORM == FES_FORM_LAPLACIAN) {
#pragma unroll
                  for (int s = 0; s < SD; ++s) acc[0][0] = fma(A_[s], B_[s], acc[0][0]);
                } else {
#pragma unroll
                  for (int a_ = 0; a_ < C; ++a_)
#pragma unroll
                    for (int b_ = 0; b_ < C; ++b_) acc[a_][b_] = fma(A_[a_], B_[b_], acc[a_][b_]);
                }
              }
            }
          }
          // balanced tree over the partial sums: partial j joins level o when bit o of j is set
          bool parked = false;
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            if (!parked && (uint32_t)o < lg) {
              if ((j >> o) & 1u) {
#pragma unroll
                for (int x = 0; x < NACC; ++x) acc[x / C][x % C] = lvl[o][x] + acc[x / C][x % C];
              } else {
#pragma unroll
                for (int x = 0; x < NACC; ++x) lvl[o][x] = acc[x / C][x % C];
                parked = true;
              }
            }
          }
        }
        w += 1 + (int)cnt;
        if (!active) continue;
        double K[C][C];
        if constexpr (FORM == FES_FORM_ELASTIC) {
          double tr = 0.0;
#pragma unroll
          for (int a_ = 0; a_ < C; ++a_) tr += acc[a_][a_];
#pragma unroll
          for (int a_ = 0; a_ < C; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < C; ++b_) K[a_][b_] = fma(cf.lam_over_mu, acc[a_][b_], acc[b_][a_]) + (a_ == b_ ? tr : 0.0);
        } else {
          K[0][0] = acc[0][0];
        }
        if constexpr (C == 2) {
          const uint32_t r0 = img_lane + 16u * (img >> 1);
          sts_f64x2(r0, K[0][0], K[0][1]);
          sts_f64x2(r0 + 16u * deg, K[1][0], K[1][1]);
        } else {
#pragma unroll
          for (int a_ = 0; a_ < C; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < C; ++b_) sts_f64(img_lane + 8u * (img + (uint32_t)(a_ * C) * deg + (uint32_t)b_), K[a_][b_]);
        }
      }
    }
    __syncthreads();  // image complete; records and coordinates free

    // ---- store: image rows -> the tile's contiguous slice of `values`
    {
      const int total = T * LU;
      unsigned char* dst = reinterpret_cast<unsigned char*>(values + (int64_t)C * C * pair0);
      for (int x = tid; x < total; x += kRunThreads) {
        const uint32_t r = __umulhi((uint32_t)x, magic), u = (uint32_t)x - r * (uint32_t)LU;
        const unsigned char* src = s_img + (size_t)(r * (uint32_t)LUp + u) * UB;
        if constexpr (UB == 16) *reinterpret_cast<double2*>(dst + (size_t)x * 16) = *reinterpret_cast<const double2*>(src);
        else *reinterpret_cast<double*>(dst + (size_t)x * 8) = *reinterpret_cast<const double*>(src);
      }
    }
  }
}

template <int ELEM, int FORM, int C, int SD>
int launch_runs(const Job& j) {
  if constexpr (!runs_supported<ELEM, FORM, C, SD>()) {
    return FES_OK;
  } else {
    const fes_runs& R = j.plan->runs;
    const RunDims dims{R.max_rec_entries, R.max_coord_nodes, R.max_img_units, R.max_tmpl_words};
    const size_t smem = RunSmem<C, SD>::total(dims);
    static size_t configured = 0, occ_smem = 0;
    static int per_sm = 0, sms = 0;
    if (per_sm == 0 || smem != occ_smem) {
      if (smem > configured && smem + 1024 > 48 * 1024)
        FES_CUDA(cudaFuncSetAttribute(k_assemble_runs<ELEM, FORM, C, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = std::max(configured, smem);
      int dev = 0;
      FES_CUDA(cudaGetDevice(&dev));
      FES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      FES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_runs<ELEM, FORM, C, SD>, kRunThreads, smem));
      occ_smem = smem;
      if (per_sm < 1) {
        per_sm = 0;
        return set_error(FES_ERR_CUDA, "run assembly kernel does not fit on an SM (%zu B smem)", smem);
      }
      if (getenv("FES_PLAN_DEBUG")) fprintf(stderr, "[fes assemble runs] smem=%zu B CTAs/SM=%d\n", smem, per_sm);
    }
    Coeffs cf = j.cf;
    if (cf.factor < 0.0) { cf.factor = -cf.factor; cf.sign = -1.0; }
    const int64_t grid = std::min<int64_t>(R.n_tiles, (int64_t)sms * per_sm);
    static const bool no_pdl = getenv("FES_NO_PDL") != nullptr;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid);
    lc.blockDim = dim3(kRunThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = j.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = no_pdl ? 0 : 1;
    FES_CUDA(cudaLaunchKernelEx(&lc, k_assemble_runs<ELEM, FORM, C, SD>, (const int32_t*)R.tiles.p, (int)R.n_tiles,
                                (const int32_t*)R.tmpl.p, (const double*)j.coords, (int64_t)j.plan->n_nodes, (int64_t)j.plan->n_el,
                                cf, dims, j.out));
    FES_LAUNCH_CHECK();
    return FES_OK;
  }
}
