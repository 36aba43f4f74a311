// SIMP inner-loop kernels: cone filter, per-element strain energy / compliance / von Mises,
// sensitivity, and the optimality-criteria bisection as one persistent cooperative kernel.
//
// Replaces calculate_smoothing_matrix (ref fem/topology.py:39-77), _element_quadratic +
// DensityField.gradient (ref fem/sensitivity.py:41-49,305-309,346-348),
// LinearElasticForm.derived_fields (ref fem/forms.py:335-374) + von_mises
// (ref fem/invariants.py:48-62) and optimality_criteria_update (ref fem/design.py:39-78).

#include "elements.cuh"
#include "gridsum.cuh"
#include "sortscan.cuh"

namespace cg = cooperative_groups;

struct fes_filter {
  int64_t n = 0, nnz = 0;
  fes::DevBuf<int32_t> indptr, indices;
  fes::DevBuf<double> data;
};

namespace fes {
namespace {

constexpr int kBlock = 256;

// ---- cone filter ------------------------------------------------------------------------------
template <int DIM>
__global__ void k_centroids(const int32_t* __restrict__ en, int64_t n_el, int N, const double* __restrict__ coords,
                            double* __restrict__ ctr) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
#pragma unroll
  for (int s = 0; s < DIM; ++s) {
    double acc = 0.0;
    for (int a = 0; a < N; ++a) acc = __dadd_rn(acc, coords[(int64_t)en[e * N + a] * DIM + s]);
    ctr[e * DIM + s] = acc / (double)N;  // numpy mean: sum, then one divide
  }
}

__global__ void k_minmax(const double* __restrict__ ctr, int64_t n, int dim, double* __restrict__ lohi) {
  // one block; lohi[s] = min, lohi[3+s] = max
  __shared__ double smin[kBlock], smax[kBlock];
  for (int s = 0; s < dim; ++s) {
    double lo = 1e300, hi = -1e300;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const double v = ctr[i * dim + s];
      lo = fmin(lo, v); hi = fmax(hi, v);
    }
    smin[threadIdx.x] = lo; smax[threadIdx.x] = hi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) {
        smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + o]);
        smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) { lohi[s] = smin[0]; lohi[3 + s] = smax[0]; }
    __syncthreads();
  }
}

struct Grid {
  double lo[3];
  double inv_cell;
  int dims[3];
  int dim;
};

__device__ __forceinline__ void cell_of(const Grid& g, const double* c, int (&ijk)[3]) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s < g.dim) {
      int v = (int)floor((c[s] - g.lo[s]) * g.inv_cell);
      ijk[s] = v < 0 ? 0 : (v >= g.dims[s] ? g.dims[s] - 1 : v);
    } else ijk[s] = 0;
  }
}

__global__ void k_cell_keys(Grid g, const double* __restrict__ ctr, int64_t n, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ ids) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int ijk[3];
  cell_of(g, ctr + e * g.dim, ijk);
  keys[e] = (uint32_t)((ijk[2] * g.dims[1] + ijk[1]) * g.dims[0] + ijk[0]);
  ids[e] = (uint32_t)e;
}

__global__ void k_cell_starts(const uint32_t* __restrict__ sorted_keys, int64_t n, int64_t n_cells,
                              int32_t* __restrict__ start) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cells) return;
  int64_t lo = 0, hi = n;  // lower_bound(sorted_keys, c)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)sorted_keys[mid] < c) lo = mid + 1; else hi = mid;
  }
  start[c] = (int32_t)lo;
}

// numpy's np.linalg.norm(ci - cj): separate roundings, no fma
template <int DIM>
__device__ __forceinline__ double dist(const double* a, const double* b) {
  double s2 = 0.0;
#pragma unroll
  for (int s = 0; s < DIM; ++s) {
    const double d = __dsub_rn(a[s], b[s]);
    s2 = __dadd_rn(s2, __dmul_rn(d, d));
  }
  return sqrt(s2);
}

// pass 0 counts row lengths, pass 1 writes (col, weight) then orders the row by column and
// normalises it by (row sum + 1e-16) (ref fem/topology.py:73-77)
template <int DIM, int PASS>
__global__ void k_filter_rows(Grid g, const double* __restrict__ ctr, int64_t n, double r,
                              const int32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_ids,
                              int32_t* __restrict__ counts, const int32_t* __restrict__ indptr,
                              int32_t* __restrict__ indices, double* __restrict__ data) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double ci[DIM];
#pragma unroll
  for (int s = 0; s < DIM; ++s) ci[s] = ctr[i * DIM + s];
  int ijk[3];
  cell_of(g, ci, ijk);
  int32_t cnt = 0;
  const int32_t base = PASS ? indptr[i] : 0;
  const int z0 = DIM == 3 ? -1 : 0, z1 = DIM == 3 ? 1 : 0;
  for (int dz = z0; dz <= z1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int cx = ijk[0] + dx, cy = ijk[1] + dy, cz = ijk[2] + dz;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= g.dims[0] || cy >= g.dims[1] || cz >= g.dims[2]) continue;
        const int64_t c = ((int64_t)cz * g.dims[1] + cy) * g.dims[0] + cx;
        for (int32_t k = cell_start[c]; k < cell_start[c + 1]; ++k) {
          const int64_t j = cell_ids[k];
          double w;
          if (j == i) w = r;
          else {
            const double d = dist<DIM>(ci, ctr + j * DIM);
            if (!(d <= r)) continue;
            w = r - d;
          }
          if (PASS) { indices[base + cnt] = (int32_t)j; data[base + cnt] = w; }
          ++cnt;
        }
      }
  if (!PASS) { counts[i] = cnt; return; }
  // insertion sort by column (rows are a few dozen entries), then normalise in column order
  for (int32_t a = 1; a < cnt; ++a) {
    const int32_t col = indices[base + a];
    const double w = data[base + a];
    int32_t b = a - 1;
    while (b >= 0 && indices[base + b] > col) {
      indices[base + b + 1] = indices[base + b];
      data[base + b + 1] = data[base + b];
      --b;
    }
    indices[base + b + 1] = col;
    data[base + b + 1] = w;
  }
  double sum = 0.0;
  for (int32_t a = 0; a < cnt; ++a) sum += data[base + a];
  const double denom = sum + 1e-16;
  for (int32_t a = 0; a < cnt; ++a) data[base + a] = data[base + a] / denom;
}

template <int G>
__global__ void k_csr_apply(int64_t n_rows, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const double* __restrict__ data, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / G;
  const int lane = (int)(t % G);
  double s = 0.0;
  if (row < n_rows) {
    const int32_t k1 = indptr[row + 1];
    for (int32_t k = indptr[row] + lane; k < k1; k += G) s += data[k] * __ldg(x + indices[k]);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (row < n_rows && lane == 0) y[row] = s;
}

__global__ void k_widen32(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = in[t];
}

// ---- per-element strain energy ------------------------------------------------------------------
template <int ELEM, int SD>
__global__ void k_elastic_energy(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords,
                                 const double* __restrict__ u, double mu0, double lam0,
                                 const double* __restrict__ scale, double* __restrict__ q_out,
                                 double* __restrict__ comp_out, double* __restrict__ vm_out) {
  constexpr int N = ET<ELEM>::N;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  Geom<ET<ELEM>::RD, SD> g;
  load_geom<ELEM, SD>(en + e * N, coords, g);
  // u_e^T K0_e u_e = sum_q w_q (lam tr(eps_q)^2 + 2 mu eps_q:eps_q) over the element's stiffness rule (one
  // point for P1, the three Strang points for P2; equal weights).  The per-element compliance and
  // von Mises outputs are the reference's first-point quantities (ref fem/forms.py:335-374), so H, tr
  // and ee of point 0 are kept.
  constexpr int NQ = ET<ELEM>::NQ;
  const double vol = g.scale * ref_measure(ET<ELEM>::RD);
  double H[SD][SD], tr = 0.0, ee = 0.0, q = 0.0;
#pragma unroll
  for (int qp = NQ - 1; qp >= 0; --qp) {  // point 0 last: its H, tr, ee stay live for the outputs below
#pragma unroll
    for (int i = 0; i < SD; ++i)
#pragma unroll
      for (int s = 0; s < SD; ++s) H[i][s] = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      double ga[SD];
      shape_grad<ELEM, SD>(g, qp, a, ga);
      const int64_t v = en[e * N + a];
#pragma unroll
      for (int i = 0; i < SD; ++i) {
        const double ui = u[v * SD + i];
#pragma unroll
        for (int s = 0; s < SD; ++s) H[i][s] += ui * ga[s];
      }
    }
    tr = 0.0; ee = 0.0;
#pragma unroll
    for (int i = 0; i < SD; ++i) {
      tr += H[i][i];
#pragma unroll
      for (int s = 0; s < SD; ++s) {
        const double eps = 0.5 * (H[i][s] + H[s][i]);
        ee += eps * eps;
      }
    }
    q += (vol / (double)NQ) * (lam0 * tr * tr + 2.0 * mu0 * ee);
  }
  const double q0 = vol * (lam0 * tr * tr + 2.0 * mu0 * ee);  // sigma:eps * vol at the first point
  const double sc = scale ? scale[e] : 1.0;
  if (q_out) q_out[e] = q;
  if (comp_out) comp_out[e] = sc * q0;
  if (vm_out) {
    // sigma = sc (lam0 tr I + 2 mu0 eps), with sigma_zz = sc lam0 tr under plane strain
    const double mean = sc * (lam0 * tr + 2.0 * mu0 * tr / 3.0);  // tr(sigma)/3, eps_zz = 0 in 2D
    double dd = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        double sig = 0.0;
        if (i < SD && s < SD) sig = sc * 2.0 * mu0 * 0.5 * (H[i][s] + H[s][i]);
        if (i == s) sig += sc * lam0 * tr - mean;
        dd += sig * sig;
      }
    vm_out[e] = sqrt(1.5 * dd);
  }
}

__global__ void k_sensitivity(const double* __restrict__ rho, const double* __restrict__ q, double penalty,
                              double chain, int64_t n, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = chain * (penalty * pow(rho[e], penalty - 1.0) * q[e]);
}

// ---- optimality criteria: the whole bisection in one cooperative kernel ----------------------
struct OcArgs {
  const double* rho;
  const double* sens;
  const double* vol;
  int64_t n;
  double vf, move, tol;
  int max_iters;
  double* out;
  double* partials;
  int* result;  // [0] = bisections, [1] = negative-sensitivity flag
};

__device__ __forceinline__ double oc_candidate(double rho, double s, double m, double move) {
  double v = rho * sqrt(s / m);
  v = fmin(fmax(v, rho - move), rho + move);
  return fmin(fmax(v, 1e-6), 1.0);
}

__global__ void __launch_bounds__(512) k_oc(OcArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[3 * 33];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  double v2[2] = {0.0, 0.0};  // sum of volumes, count of negative sensitivities
  for (int64_t e = tid; e < A.n; e += nth) {
    v2[0] += A.vol[e];
    v2[1] += (A.sens[e] < 0.0) ? 1.0 : 0.0;
  }
  grid_sum<2>(grid, v2, A.partials, 0, sm);
  const double vsum = v2[0];
  if (v2[1] > 0.0) {
    if (tid == 0) { A.result[0] = 0; A.result[1] = 1; }
    return;
  }
  double lo = 0.0, hi = 1e15, m = 0.0;
  int it = 0;
  for (; it < A.max_iters;) {
    m = 0.5 * (lo + hi);
    double v1[1] = {0.0};
    for (int64_t e = tid; e < A.n; e += nth) v1[0] += A.vol[e] * oc_candidate(A.rho[e], A.sens[e], m, A.move);
    grid_sum<1>(grid, v1, A.partials, 2 + (it & 1), sm);  // alternate slots: no WAR hazard between rounds
    if (v1[0] / vsum < A.vf) hi = m; else lo = m;
    ++it;
    if (hi - lo <= A.tol * hi) break;
  }
  // rho_new is the candidate of the LAST multiplier tried (ref fem/design.py:66-78)
  for (int64_t e = tid; e < A.n; e += nth)
    A.out[e] = (A.max_iters > 0) ? oc_candidate(A.rho[e], A.sens[e], m, A.move) : A.rho[e];
  if (tid == 0) { A.result[0] = it; A.result[1] = 0; }
}

}  // namespace
}  // namespace fes

using namespace fes;

extern "C" int fes_filter_create(const int32_t* d_elem_nodes, int64_t n_el, int N, const double* d_coords, int dim,
                                 double radius, void* stream_, fes_filter** out) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!out) return set_error(FES_ERR_VALUE, "fes_filter_create: out is NULL");
  *out = nullptr;
  if (dim != 2 && dim != 3) return set_error(FES_ERR_VALUE, "fes_filter_create: dim must be 2 or 3");
  if (n_el <= 0 || n_el >= (int64_t)1 << 31 || radius < 0.0)
    return set_error(FES_ERR_VALUE, "fes_filter_create: bad n_el=%lld or radius=%g", (long long)n_el, radius);
  fes_filter* F = new fes_filter();
  struct Guard { fes_filter* p; ~Guard() { delete p; } } guard{F};
  F->n = n_el;

  DevBuf<double> ctr, lohi;
  FES_CUDA(ctr.alloc((size_t)n_el * dim));
  FES_CUDA(lohi.alloc(6));
  if (dim == 2) k_centroids<2><<<grid_for(n_el, kBlock), kBlock, 0, st>>>(d_elem_nodes, n_el, N, d_coords, ctr.p);
  else k_centroids<3><<<grid_for(n_el, kBlock), kBlock, 0, st>>>(d_elem_nodes, n_el, N, d_coords, ctr.p);
  FES_LAUNCH_CHECK();
  k_minmax<<<1, kBlock, 0, st>>>(ctr.p, n_el, dim, lohi.p);
  FES_LAUNCH_CHECK();
  double h[6];
  FES_CUDA(cudaMemcpyAsync(h, lohi.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));

  // uniform grid with cells no smaller than the radius, so neighbours lie in the 3^dim block
  Grid g{};
  g.dim = dim;
  double cell = radius > 0.0 ? radius : 1.0;
  for (;;) {
    double cells = 1.0;
    for (int s = 0; s < 3; ++s) {
      g.dims[s] = s < dim ? (int)floor((h[3 + s] - h[s]) / cell) + 1 : 1;
      cells *= g.dims[s];
    }
    if (cells <= 6.4e7) break;
    cell *= 2.0;
  }
  for (int s = 0; s < 3; ++s) g.lo[s] = s < dim ? h[s] : 0.0;
  g.inv_cell = 1.0 / cell;
  const int64_t n_cells = (int64_t)g.dims[0] * g.dims[1] * g.dims[2];

  DevBuf<uint32_t> keys_a, keys_b, ids_a, ids_b;
  DevBuf<int32_t> cell_start, counts;
  FES_CUDA(keys_a.alloc(n_el)); FES_CUDA(keys_b.alloc(n_el));
  FES_CUDA(ids_a.alloc(n_el)); FES_CUDA(ids_b.alloc(n_el));
  FES_CUDA(cell_start.alloc(n_cells + 1));
  FES_CUDA(counts.alloc(n_el + 1));
  FES_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes(), st));
  k_cell_keys<<<grid_for(n_el, kBlock), kBlock, 0, st>>>(g, ctr.p, n_el, keys_a.p, ids_a.p);
  FES_LAUNCH_CHECK();
  {
    int key_bits = 1;  // cell ids are < n_cells
    while (key_bits < 32 && ((uint64_t)n_cells >> key_bits) != 0) ++key_bits;
    bool in_b = false;
    FES_TRY(radix_sort_pairs<uint32_t>(keys_a.p, keys_b.p, ids_a.p, ids_b.p, n_el, key_bits, st, &in_b));
    if (!in_b) {
      FES_CUDA(cudaMemcpyAsync(keys_b.p, keys_a.p, keys_a.bytes(), cudaMemcpyDeviceToDevice, st));
      FES_CUDA(cudaMemcpyAsync(ids_b.p, ids_a.p, ids_a.bytes(), cudaMemcpyDeviceToDevice, st));
    }
  }
  k_cell_starts<<<grid_for(n_cells + 1, kBlock), kBlock, 0, st>>>(keys_b.p, n_el, n_cells, cell_start.p);
  FES_LAUNCH_CHECK();

  FES_CUDA(F->indptr.alloc(n_el + 1));
#define FES_ROWS(DIM_, PASS_)                                                                                 \
  k_filter_rows<DIM_, PASS_><<<grid_for(n_el, 128), 128, 0, st>>>(g, ctr.p, n_el, radius, cell_start.p, ids_b.p, \
                                                                  counts.p, F->indptr.p, F->indices.p, F->data.p)
  if (dim == 2) FES_ROWS(2, 0); else FES_ROWS(3, 0);
  FES_LAUNCH_CHECK();
  FES_TRY((device_scan<int32_t, false>(counts.p, F->indptr.p, n_el + 1, st)));
  int32_t nnz = 0;
  FES_CUDA(cudaMemcpyAsync(&nnz, F->indptr.p + n_el, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  F->nnz = nnz;
  FES_CUDA(F->indices.alloc(nnz));
  FES_CUDA(F->data.alloc(nnz));
  if (dim == 2) FES_ROWS(2, 1); else FES_ROWS(3, 1);
#undef FES_ROWS
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaStreamSynchronize(st));
  guard.p = nullptr;
  *out = F;
  return FES_OK;
}

extern "C" void fes_filter_destroy(fes_filter* f) { delete f; }
extern "C" int64_t fes_filter_nnz(const fes_filter* f) { return f ? f->nnz : 0; }

extern "C" int fes_filter_export_csr(const fes_filter* f, int64_t* h_indptr, int64_t* h_indices, double* h_data) {
  if (!f || !h_indptr || !h_indices || !h_data) return set_error(FES_ERR_VALUE, "fes_filter_export_csr: NULL argument");
  const int64_t n1 = f->n + 1, big = n1 > f->nnz ? n1 : f->nnz;
  DevBuf<int64_t> wide;
  FES_CUDA(wide.alloc(big));
  k_widen32<<<grid_for(n1, kBlock), kBlock>>>(f->indptr.p, n1, wide.p);
  FES_LAUNCH_CHECK();
  FES_CUDA(cudaMemcpy(h_indptr, wide.p, n1 * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (f->nnz) {
    k_widen32<<<grid_for(f->nnz, kBlock), kBlock>>>(f->indices.p, f->nnz, wide.p);
    FES_LAUNCH_CHECK();
    FES_CUDA(cudaMemcpy(h_indices, wide.p, f->nnz * sizeof(int64_t), cudaMemcpyDeviceToHost));
    FES_CUDA(cudaMemcpy(h_data, f->data.p, f->nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return FES_OK;
}

extern "C" int fes_filter_apply(const fes_filter* f, const double* d_in, double* d_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!f) return set_error(FES_ERR_VALUE, "fes_filter_apply: NULL filter");
  const double avg = (double)f->nnz / (double)f->n;
  if (avg <= 12.0) k_csr_apply<4><<<grid_for(f->n * 4, kBlock), kBlock, 0, st>>>(f->n, f->indptr.p, f->indices.p, f->data.p, d_in, d_out);
  else if (avg <= 48.0) k_csr_apply<16><<<grid_for(f->n * 16, kBlock), kBlock, 0, st>>>(f->n, f->indptr.p, f->indices.p, f->data.p, d_in, d_out);
  else k_csr_apply<32><<<grid_for(f->n * 32, kBlock), kBlock, 0, st>>>(f->n, f->indptr.p, f->indices.p, f->data.p, d_in, d_out);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_elastic_energy(int elem_kind, const int32_t* en, int64_t n_el, const double* coords, const double* u,
                                  double E0, double nu, const double* scale, double* q, double* comp, double* vm,
                                  void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_el <= 0) return FES_OK;
  const double mu0 = E0 / (2.0 * (1.0 + nu)), lam0 = E0 * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  const int g = grid_for(n_el, kBlock);
  if (elem_kind == FES_P1_TRI) k_elastic_energy<FES_P1_TRI, 2><<<g, kBlock, 0, st>>>(en, n_el, coords, u, mu0, lam0, scale, q, comp, vm);
  else if (elem_kind == FES_P1_TET) k_elastic_energy<FES_P1_TET, 3><<<g, kBlock, 0, st>>>(en, n_el, coords, u, mu0, lam0, scale, q, comp, vm);
  else if (elem_kind == FES_P2_TRI) k_elastic_energy<FES_P2_TRI, 2><<<g, kBlock, 0, st>>>(en, n_el, coords, u, mu0, lam0, scale, q, comp, vm);
  else return set_error(FES_ERR_UNSUPPORTED, "no elastic recovery for element kind %d", elem_kind);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_simp_sensitivity(const double* d_rho, const double* d_q, double penalty, double chain, int64_t n_el,
                                    double* d_sens, void* stream) {
  if (n_el <= 0) return FES_OK;
  k_sensitivity<<<grid_for(n_el, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_rho, d_q, penalty, chain, n_el, d_sens);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_oc_update(const double* d_rho, const double* d_sens, const double* d_vol, int64_t n_el,
                             double volume_frac, double move, int max_iters, double tol, double* d_rho_out,
                             int* bisections_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_el <= 0) return FES_OK;
  static int s_grid = 0;
  if (s_grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    FES_CUDA(cudaGetDevice(&dev));
    FES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_oc, 512, 0));
    if (per_sm < 1) return set_error(FES_ERR_CUDA, "OC kernel does not fit on an SM");
    s_grid = sms * (per_sm > 2 ? 2 : per_sm);
    if (s_grid > kMaxGrid) s_grid = kMaxGrid;
  }
  DevBuf<double> partials;
  DevBuf<int> result;
  FES_CUDA(partials.alloc(4 * kMaxGrid));
  FES_CUDA(result.alloc(2));
  OcArgs a{d_rho, d_sens, d_vol, n_el, volume_frac, move, tol, max_iters, d_rho_out, partials.p, result.p};
  void* params[] = {&a};
  FES_CUDA(cudaLaunchCooperativeKernel((void*)k_oc, dim3(s_grid), dim3(512), params, 0, st));
  count_launch();
  int h[2] = {0, 0};
  FES_CUDA(cudaMemcpyAsync(h, result.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  if (h[1])
    return set_error(FES_ERR_VALUE, "the optimality-criteria update needs a nonnegative sensitivity (adding material "
                     "must not raise the objective), which holds for compliance-type objectives.");
  if (bisections_out) *bisections_out = h[0];
  return FES_OK;
}
