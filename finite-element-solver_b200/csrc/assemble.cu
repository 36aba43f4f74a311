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
#include "elements.cuh"
#include "plan.cuh"

namespace fes {
namespace {

constexpr int kBlock = 128;

struct Coeffs {
  double E, nu, factor;
  const double* d_E;
  const double* d_scale;
};

__device__ __forceinline__ void element_coeffs(const Coeffs& c, int64_t e, Material& m, double& w) {
  m = lame(c.d_E ? __ldg(c.d_E + e) : c.E, c.nu);
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
    k_assemble_pairs<ELEM, FORM, C, SD><<<grid_for(P->n_pairs, kBlock), kBlock, 0, j.stream>>>(
        P->n_pairs, P->pair_start.p, P->pair_row.p, P->node_ptr.p, P->contrib.p, P->indptr.p, j.en, j.coords, j.cf,
        j.out);
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
  out = Coeffs{c->E, c->nu, c->factor, c->d_E, c->d_scale};
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
  DevBuf<double> coords, E, scale, values;
  FES_CUDA(coords.alloc((size_t)n_nodes * spatial_dim));
  FES_CUDA(values.alloc(plan->nnz));
  FES_CUDA(cudaMemcpyAsync(coords.p, h_coords, coords.bytes(), cudaMemcpyHostToDevice, st));
  fes_coeffs c = *coeffs;
  if (h_E) {
    FES_CUDA(E.alloc(plan->n_el));
    FES_CUDA(cudaMemcpyAsync(E.p, h_E, E.bytes(), cudaMemcpyHostToDevice, st));
    c.d_E = E.p;
  }
  if (h_scale) {
    FES_CUDA(scale.alloc(plan->n_el));
    FES_CUDA(cudaMemcpyAsync(scale.p, h_scale, scale.bytes(), cudaMemcpyHostToDevice, st));
    c.d_scale = scale.p;
  }
  FES_TRY(fes_assemble(plan, elem_kind, form, spatial_dim, d_elem_nodes, coords.p, &c, values.p, st));
  FES_CUDA(cudaMemcpyAsync(h_values, values.p, values.bytes(), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  return FES_OK;
}
