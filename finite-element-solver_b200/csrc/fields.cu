// Quadrature-tabulated forms and nodal recovery: the callers either side of the fused assembly
// path.
//
//   fes_node_scatter            _VectorScatterPlan.scatter / the np.add.at of recover_nodal
//                               (ref fem/space.py:148-184, 578-596): per-(element, local node)
//                               values summed into their nodes in ascending element id -- the order
//                               np.bincount adds them -- by a GATHER over the plan's diagonal pairs.
//   fes_quadrature_points       ElementGeometry.points / weight_detJ at a tabulated rule
//                               (ref fem/elements.py:174-212, fem/quadrature.py:54-104).
//   fes_element_loads           LinearForm.element_vectors (ref fem/forms.py:296-315).
//   fes_weighted_gradient_matrices
//                               DiffusionForm (kappa per quadrature point, ref fem/forms.py:273-293) and
//                               GeometricStiffnessForm (prestress tensor per element, ref :401-441):
//                               K[(a,i),(b,j)] = delta_ij sum_q w_q kappa_q grad(phi_a)^T T grad(phi_b).
//   fes_elastic_fields          LinearElasticForm.derived_fields (ref fem/forms.py:335-374): full 3x3
//                               strain / stress at the first quadrature point, plane-strain lift.
#include "elements.cuh"
#include "plan.cuh"

namespace fes {
namespace {

constexpr int kBlock = 256;
constexpr int kMaxQ = 6;

// ---- tabulated rules (ref fem/quadrature.py:54-104): cheapest rule of degree >= min_degree ----
struct Rule {
  int nq;
  double pts[kMaxQ][3];
  double wts[kMaxQ];
};

__host__ __device__ inline bool quad_rule(int rd, int min_degree, Rule& r) {
  const double s3 = 0.28867513459481288225, s35 = 0.38729833462074168852;  // 0.5/sqrt(3), 0.5*sqrt(3/5)
  if (rd == 1) {
    if (min_degree <= 1) { r.nq = 1; r.pts[0][0] = 0.5; r.wts[0] = 1.0; return true; }
    if (min_degree <= 3) {
      r.nq = 2; r.pts[0][0] = 0.5 - s3; r.pts[1][0] = 0.5 + s3; r.wts[0] = r.wts[1] = 0.5; return true;
    }
    if (min_degree <= 5) {
      r.nq = 3; r.pts[0][0] = 0.5 - s35; r.pts[1][0] = 0.5; r.pts[2][0] = 0.5 + s35;
      r.wts[0] = 5.0 / 18; r.wts[1] = 8.0 / 18; r.wts[2] = 5.0 / 18; return true;
    }
    return false;
  }
  if (rd == 2) {
    if (min_degree <= 1) { r.nq = 1; r.pts[0][0] = r.pts[0][1] = 1.0 / 3; r.wts[0] = 0.5; return true; }
    if (min_degree <= 2) {
      r.nq = 3;
      const double p[3][2] = {{1.0 / 6, 1.0 / 6}, {2.0 / 3, 1.0 / 6}, {1.0 / 6, 2.0 / 3}};
      for (int q = 0; q < 3; ++q) { r.pts[q][0] = p[q][0]; r.pts[q][1] = p[q][1]; r.wts[q] = 1.0 / 6; }
      return true;
    }
    if (min_degree <= 4) {
      r.nq = 6;
      const double a = 0.44594849091596489, b = 0.10810301816807022, c = 0.09157621350977074, d = 0.81684757298045851;
      const double p[6][2] = {{a, a}, {b, a}, {a, b}, {c, c}, {d, c}, {c, d}};
      for (int q = 0; q < 6; ++q) {
        r.pts[q][0] = p[q][0]; r.pts[q][1] = p[q][1];
        r.wts[q] = q < 3 ? 0.11169079483900573 : 0.05497587182766094;
      }
      return true;
    }
    return false;
  }
  if (rd == 3) {
    if (min_degree <= 1) { r.nq = 1; r.pts[0][0] = r.pts[0][1] = r.pts[0][2] = 0.25; r.wts[0] = 1.0 / 6; return true; }
    if (min_degree <= 2) {
      r.nq = 4;
      const double A = 0.5854101966249685, B = 0.1381966011250105;
      for (int q = 0; q < 4; ++q) {
        for (int k = 0; k < 3; ++k) r.pts[q][k] = (q == k) ? A : B;
        r.wts[q] = 1.0 / 24;
      }
      return true;
    }
    return false;
  }
  return false;
}

// shape values / reference gradients at a reference point (ref fem/elements.py:258-297, 369-396)
template <int ELEM>
__device__ __forceinline__ void shape_at(const double* xi, double (&phi)[ET<ELEM>::N]) {
  constexpr int RD = ET<ELEM>::RD;
  if constexpr (!ET<ELEM>::P2) {
    double l0 = 1.0;
#pragma unroll
    for (int r = 0; r < RD; ++r) { l0 -= xi[r]; phi[r + 1] = xi[r]; }
    phi[0] = l0;
  } else {
    static_assert(ELEM == FES_P2_TRI, "P2 volume forms are implemented for the triangle");
    const double x = xi[0], y = xi[1], l0 = 1.0 - x - y;
    phi[0] = l0 * (2 * l0 - 1); phi[1] = x * (2 * x - 1); phi[2] = y * (2 * y - 1);
    phi[3] = 4 * x * y; phi[4] = 4 * l0 * y; phi[5] = 4 * l0 * x;
  }
}

template <int ELEM>
__device__ __forceinline__ void dshape_at(const double* xi, double (&d)[ET<ELEM>::N][ET<ELEM>::RD]) {
  constexpr int RD = ET<ELEM>::RD, N = ET<ELEM>::N;
  if constexpr (!ET<ELEM>::P2) {
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
      for (int r = 0; r < RD; ++r) d[n][r] = (n == 0) ? -1.0 : (n == r + 1 ? 1.0 : 0.0);
  } else {
    const double x = xi[0], y = xi[1], l0 = 1.0 - x - y;
    d[0][0] = d[0][1] = -(4 * l0 - 1);
    d[1][0] = 4 * x - 1; d[1][1] = 0.0;
    d[2][0] = 0.0;       d[2][1] = 4 * y - 1;
    d[3][0] = 4 * y;     d[3][1] = 4 * x;
    d[4][0] = -4 * y;    d[4][1] = 4 * l0 - 4 * y;
    d[5][0] = 4 * l0 - 4 * x; d[5][1] = -4 * x;
  }
}

template <int ELEM, int SD>
__device__ __forceinline__ void load_corners(const int32_t* __restrict__ en, const double* __restrict__ coords,
                                             double (&X)[ET<ELEM>::RD + 1][SD]) {
#pragma unroll
  for (int n = 0; n <= ET<ELEM>::RD; ++n) {
    const int64_t v = en[n];
#pragma unroll
    for (int s = 0; s < SD; ++s) X[n][s] = __ldg(coords + v * SD + s);
  }
}

// ---- node scatter as a gather over the diagonal pair of every node --------------------------
__global__ void __launch_bounds__(kBlock)
k_node_scatter(int64_t n_nodes, int N, int m, const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ node_adj,
               const uint32_t* __restrict__ pair_start, const uint32_t* __restrict__ contrib,
               const double* __restrict__ vals, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes * m) return;
  const int64_t v = t / m;
  const int j = (int)(t - v * m);
  int lo = node_ptr[v], hi = node_ptr[v + 1];
  const int end = hi;
  while (lo < hi) {  // the stored pair (v, v): neighbours ascend
    const int mid = (lo + hi) >> 1;
    if (node_adj[mid] < (int32_t)v) lo = mid + 1; else hi = mid;
  }
  double acc = 0.0;
  if (lo < end && node_adj[lo] == (int32_t)v) {
    const uint32_t N2 = (uint32_t)(N * N);
    for (uint32_t k = pair_start[lo]; k < pair_start[lo + 1]; ++k) {
      const uint32_t code = contrib[k], e = code / N2, ab = code - e * N2, a = ab / (uint32_t)N, b = ab - a * (uint32_t)N;
      if (a == b) acc = __dadd_rn(acc, __ldg(vals + ((int64_t)e * N + a) * m + j));  // ascending (e, a): bincount's order
    }
  }
  out[t] = acc;
}

// ---- physical quadrature points and weights ---------------------------------------------------
template <int ELEM, int SD>
__global__ void __launch_bounds__(kBlock)
k_quad_points(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords, Rule rule,
              double* __restrict__ points, double* __restrict__ wdet) {
  constexpr int RD = ET<ELEM>::RD, N = ET<ELEM>::N;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  double X[RD + 1][SD];
  load_corners<ELEM, SD>(en + e * N, coords, X);
  Geom<RD, SD> g;
  make_geom<RD, SD>(X, g);
  for (int q = 0; q < rule.nq; ++q) {
    if (wdet) wdet[e * rule.nq + q] = g.scale * rule.wts[q];
    if (points) {
      // the element is affine (straight-sided P2 included): x = X0 + sum_r xi_r (X_{r+1} - X0)
#pragma unroll
      for (int s = 0; s < SD; ++s) {
        double x = X[0][s];
#pragma unroll
        for (int r = 0; r < RD; ++r) x += rule.pts[q][r] * (X[r + 1][s] - X[0][s]);
        points[(e * rule.nq + q) * SD + s] = x;
      }
    }
  }
}

// ---- b[e, n, c] = sum_q w_eq phi_n(q) f[e, q, c] ---------------------------------------------
template <int ELEM, int SD>
__global__ void __launch_bounds__(kBlock)
k_element_loads(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords, Rule rule,
                const double* __restrict__ f, int c, double* __restrict__ out) {
  constexpr int RD = ET<ELEM>::RD, N = ET<ELEM>::N;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  double X[RD + 1][SD];
  load_corners<ELEM, SD>(en + e * N, coords, X);
  Geom<RD, SD> g;
  make_geom<RD, SD>(X, g);
  for (int j = 0; j < c; ++j) {
    double b[N];
#pragma unroll
    for (int n = 0; n < N; ++n) b[n] = 0.0;
    for (int q = 0; q < rule.nq; ++q) {
      double phi[N];
      shape_at<ELEM>(rule.pts[q], phi);
      const double wf = (g.scale * rule.wts[q]) * __ldg(f + (e * rule.nq + q) * c + j);
#pragma unroll
      for (int n = 0; n < N; ++n) b[n] += phi[n] * wf;
    }
#pragma unroll
    for (int n = 0; n < N; ++n) out[(e * N + n) * c + j] = b[n];
  }
}

// ---- K[(a,i),(b,j)] = delta_ij sum_q w_q kappa_q grad_a^T T grad_b ---------------------------
template <int ELEM, int SD>
__global__ void __launch_bounds__(kBlock)
k_weighted_gradient(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords, Rule rule,
                    const double* __restrict__ kappa, const double* __restrict__ tensor, int c,
                    double* __restrict__ Ke) {
  constexpr int RD = ET<ELEM>::RD, N = ET<ELEM>::N;
  static_assert(RD == SD, "gradient forms need a volume element");
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_el * N) return;
  const int64_t e = t / N;
  const int a = (int)(t - e * N);
  double X[RD + 1][SD];
  load_corners<ELEM, SD>(en + e * N, coords, X);
  Geom<RD, SD> g;
  make_geom<RD, SD>(X, g);
  double T[SD][SD];
#pragma unroll
  for (int i = 0; i < SD; ++i)
#pragma unroll
    for (int j = 0; j < SD; ++j) T[i][j] = tensor ? __ldg(tensor + (e * SD + i) * SD + j) : (i == j ? 1.0 : 0.0);
  double s[N];
#pragma unroll
  for (int b = 0; b < N; ++b) s[b] = 0.0;
  for (int q = 0; q < rule.nq; ++q) {
    double d[N][RD];
    dshape_at<ELEM>(rule.pts[q], d);
    const double w = (g.scale * rule.wts[q]) * (kappa ? __ldg(kappa + e * rule.nq + q) : 1.0);
    double ga[SD], Tg[SD];
#pragma unroll
    for (int x = 0; x < SD; ++x) {
      ga[x] = 0.0;
#pragma unroll
      for (int r = 0; r < RD; ++r) ga[x] += d[a][r] * g.Jinv[r][x];
    }
#pragma unroll
    for (int j = 0; j < SD; ++j) {  // (grad_a^T T)_j
      Tg[j] = 0.0;
#pragma unroll
      for (int i = 0; i < SD; ++i) Tg[j] += ga[i] * T[i][j];
    }
#pragma unroll
    for (int b = 0; b < N; ++b) {
      double dot = 0.0;
#pragma unroll
      for (int x = 0; x < SD; ++x) {
        double gb = 0.0;
#pragma unroll
        for (int r = 0; r < RD; ++r) gb += d[b][r] * g.Jinv[r][x];
        dot += Tg[x] * gb;
      }
      s[b] += w * dot;
    }
  }
  const int K = N * c;
  double* row = Ke + e * K * K;
  for (int i = 0; i < c; ++i)
    for (int b = 0; b < N; ++b)
      for (int j = 0; j < c; ++j) row[(a * c + i) * K + (b * c + j)] = (i == j) ? s[b] : 0.0;
}

// ---- strain / stress tensors at the first quadrature point -----------------------------------
template <int ELEM, int SD>
__global__ void __launch_bounds__(kBlock)
k_elastic_fields(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords,
                 const double* __restrict__ u, const double* __restrict__ d_E, double E, double nu, Rule rule,
                 double* __restrict__ strain, double* __restrict__ stress, double* __restrict__ compliance) {
  constexpr int RD = ET<ELEM>::RD, N = ET<ELEM>::N;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  double X[RD + 1][SD];
  load_corners<ELEM, SD>(en + e * N, coords, X);
  Geom<RD, SD> g;
  make_geom<RD, SD>(X, g);
  double d[N][RD];
  dshape_at<ELEM>(rule.pts[0], d);
  double H[SD][SD];  // H[i][x] = d u_i / d x_x
#pragma unroll
  for (int i = 0; i < SD; ++i)
#pragma unroll
    for (int x = 0; x < SD; ++x) H[i][x] = 0.0;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const int64_t v = en[e * N + n];
    double gn[SD];
#pragma unroll
    for (int x = 0; x < SD; ++x) {
      gn[x] = 0.0;
#pragma unroll
      for (int r = 0; r < RD; ++r) gn[x] += d[n][r] * g.Jinv[r][x];
    }
#pragma unroll
    for (int i = 0; i < SD; ++i) {
      const double ui = __ldg(u + v * SD + i);
#pragma unroll
      for (int x = 0; x < SD; ++x) H[i][x] += ui * gn[x];
    }
  }
  const double Ee = d_E ? __ldg(d_E + e) : E;
  const Material m = lame(Ee, nu);
  double eps[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, sig[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < SD; ++i) {
    tr += H[i][i];
#pragma unroll
    for (int j = 0; j < SD; ++j) eps[i][j] = (i == j) ? H[i][i] : (H[i][j] + H[j][i]) / 2.0;  // engineering shear / 2
  }
#pragma unroll
  for (int i = 0; i < SD; ++i)
#pragma unroll
    for (int j = 0; j < SD; ++j)
      sig[i][j] = (i == j) ? (m.lam * tr + 2.0 * m.mu * eps[i][i]) : m.mu * (H[i][j] + H[j][i]);
  if (SD == 2) sig[2][2] = m.lam * tr;  // plane strain: the stress that holds eps_zz at zero
  double vol = 0.0;
  for (int q = 0; q < rule.nq; ++q) vol += g.scale * rule.wts[q];
  double comp = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      comp += sig[i][j] * eps[i][j];
      if (strain) strain[e * 9 + i * 3 + j] = eps[i][j];
      if (stress) stress[e * 9 + i * 3 + j] = sig[i][j];
    }
  if (compliance) compliance[e] = comp * vol;
}

// ---- Cauchy stress at EVERY quadrature point (ref fem/forms.py:376-398, LinearElasticForm.stress_field) ------
// (n_el, n_qp, SD, SD), in-plane only (no plane-strain lift): the spatially resolved counterpart of the
// per-element tensors above; a P2 displacement has a strain that varies linearly within the element.
template <int ELEM, int SD>
__global__ void __launch_bounds__(kBlock)
k_stress_field(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords,
               const double* __restrict__ u, const double* __restrict__ d_E, double E, double nu, Rule rule,
               double* __restrict__ stress) {
  constexpr int RD = ET<ELEM>::RD, N = ET<ELEM>::N;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  double X[RD + 1][SD];
  load_corners<ELEM, SD>(en + e * N, coords, X);
  Geom<RD, SD> g;
  make_geom<RD, SD>(X, g);
  const Material m = lame(d_E ? __ldg(d_E + e) : E, nu);
  double ue[N][SD];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const int64_t v = en[e * N + n];
#pragma unroll
    for (int i = 0; i < SD; ++i) ue[n][i] = __ldg(u + v * SD + i);
  }
  for (int q = 0; q < rule.nq; ++q) {
    double d[N][RD];
    dshape_at<ELEM>(rule.pts[q], d);
    double H[SD][SD];  // H[i][x] = d u_i / d x_x at this point
#pragma unroll
    for (int i = 0; i < SD; ++i)
#pragma unroll
      for (int x = 0; x < SD; ++x) H[i][x] = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
#pragma unroll
      for (int x = 0; x < SD; ++x) {
        double gx = 0.0;
#pragma unroll
        for (int r = 0; r < RD; ++r) gx += d[n][r] * g.Jinv[r][x];
#pragma unroll
        for (int i = 0; i < SD; ++i) H[i][x] += ue[n][i] * gx;
      }
    }
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < SD; ++i) tr += H[i][i];
    double* out = stress + (e * rule.nq + q) * (SD * SD);
#pragma unroll
    for (int i = 0; i < SD; ++i)
#pragma unroll
      for (int j = 0; j < SD; ++j)
        out[i * SD + j] = (i == j) ? (m.lam * tr + 2.0 * m.mu * H[i][i]) : m.mu * (H[i][j] + H[j][i]);
  }
}

// ---- gradient of a scalar nodal field at the first quadrature point (ref fem/space.py:511-517) ------------
template <int ELEM, int SD>
__global__ void __launch_bounds__(kBlock)
k_scalar_gradient(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords,
                  const double* __restrict__ u, Rule rule, double* __restrict__ grad) {
  constexpr int RD = ET<ELEM>::RD, N = ET<ELEM>::N;
  static_assert(RD == SD, "gradients need a volume element");
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  double X[RD + 1][SD];
  load_corners<ELEM, SD>(en + e * N, coords, X);
  Geom<RD, SD> g;
  make_geom<RD, SD>(X, g);
  double d[N][RD];
  dshape_at<ELEM>(rule.pts[0], d);
  double out[SD];
#pragma unroll
  for (int x = 0; x < SD; ++x) out[x] = 0.0;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const double un = __ldg(u + en[e * N + n]);
#pragma unroll
    for (int x = 0; x < SD; ++x) {
      double gx = 0.0;
#pragma unroll
      for (int r = 0; r < RD; ++r) gx += d[n][r] * g.Jinv[r][x];
      out[x] += un * gx;
    }
  }
#pragma unroll
  for (int x = 0; x < SD; ++x) grad[e * SD + x] = out[x];
}

int rule_for(int elem_kind, int min_degree, Rule& r) {
  if (!quad_rule(elem_ref_dim(elem_kind), min_degree, r))
    return set_error(FES_ERR_UNSUPPORTED, "no quadrature rule of degree >= %d for reference_dim=%d", min_degree,
                     elem_ref_dim(elem_kind));
  return FES_OK;
}

}  // namespace
}  // namespace fes

using namespace fes;

#define FES_VOLUME_DISPATCH(CALL)                                                              \
  if (elem_kind == FES_P1_LINE && sd == 1) { CALL(FES_P1_LINE, 1); }                            \
  else if (elem_kind == FES_P1_TRI && sd == 2) { CALL(FES_P1_TRI, 2); }                         \
  else if (elem_kind == FES_P1_TET && sd == 3) { CALL(FES_P1_TET, 3); }                         \
  else if (elem_kind == FES_P2_TRI && sd == 2) { CALL(FES_P2_TRI, 2); }                         \
  else return set_error(FES_ERR_UNSUPPORTED, "no volume element of kind %d in %dD", elem_kind, sd)

extern "C" int fes_node_scatter(const fes_plan* P, const double* d_vals, int m, double* d_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!P || m < 1) return set_error(FES_ERR_VALUE, "fes_node_scatter: bad arguments");
  if (P->n_nodes == 0) return FES_OK;
  if (P->n_pairs == 0) {
    FES_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * P->n_nodes * m, st));
    return FES_OK;
  }
  k_node_scatter<<<grid_for(P->n_nodes * m, kBlock), kBlock, 0, st>>>(P->n_nodes, P->N, m, P->node_ptr.p, P->node_adj.p,
                                                                     P->pair_start.p, P->contrib.p, d_vals, d_out);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_quadrature_size(int elem_kind, int min_degree) {
  Rule r;
  return quad_rule(elem_ref_dim(elem_kind), min_degree, r) ? r.nq : 0;
}

extern "C" int fes_quadrature_points(int elem_kind, int sd, int min_degree, const int32_t* en, int64_t n_el,
                                     const double* coords, double* d_points, double* d_wdet, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Rule r;
  FES_TRY(rule_for(elem_kind, min_degree, r));
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el, kBlock);
#define FES_CALL(EL, SD_) k_quad_points<EL, SD_><<<g, kBlock, 0, st>>>(en, n_el, coords, r, d_points, d_wdet)
  FES_VOLUME_DISPATCH(FES_CALL);
#undef FES_CALL
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_element_loads(int elem_kind, int sd, int min_degree, const int32_t* en, int64_t n_el,
                                 const double* coords, const double* d_f_qp, int n_comp, double* d_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Rule r;
  FES_TRY(rule_for(elem_kind, min_degree, r));
  if (n_comp < 1) return set_error(FES_ERR_VALUE, "fes_element_loads: n_comp must be >= 1");
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el, kBlock);
#define FES_CALL(EL, SD_) k_element_loads<EL, SD_><<<g, kBlock, 0, st>>>(en, n_el, coords, r, d_f_qp, n_comp, d_out)
  FES_VOLUME_DISPATCH(FES_CALL);
#undef FES_CALL
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_weighted_gradient_matrices(int elem_kind, int sd, int n_comp, int min_degree, const int32_t* en,
                                              int64_t n_el, const double* coords, const double* d_kappa_qp,
                                              const double* d_tensor, double* d_Ke, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Rule r;
  FES_TRY(rule_for(elem_kind, min_degree, r));
  if (n_comp < 1 || n_comp > 3) return set_error(FES_ERR_VALUE, "fes_weighted_gradient_matrices: n_comp must be 1..3");
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el * elem_nodes_of(elem_kind), kBlock);
#define FES_CALL(EL, SD_) \
  k_weighted_gradient<EL, SD_><<<g, kBlock, 0, st>>>(en, n_el, coords, r, d_kappa_qp, d_tensor, n_comp, d_Ke)
  FES_VOLUME_DISPATCH(FES_CALL);
#undef FES_CALL
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_elastic_fields(int elem_kind, const int32_t* en, int64_t n_el, const double* coords,
                                  const double* d_u, const double* d_E, double E, double nu, int min_degree,
                                  double* d_strain, double* d_stress, double* d_compliance, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Rule r;
  FES_TRY(rule_for(elem_kind, min_degree, r));
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el, kBlock);
  if (elem_kind == FES_P1_TRI)
    k_elastic_fields<FES_P1_TRI, 2><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_strain, d_stress, d_compliance);
  else if (elem_kind == FES_P1_TET)
    k_elastic_fields<FES_P1_TET, 3><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_strain, d_stress, d_compliance);
  else if (elem_kind == FES_P2_TRI)
    k_elastic_fields<FES_P2_TRI, 2><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_strain, d_stress, d_compliance);
  else
    return set_error(FES_ERR_UNSUPPORTED, "no elastic constitutive matrix for element kind %d", elem_kind);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_elastic_stress_field(int elem_kind, const int32_t* en, int64_t n_el, const double* coords,
                                        const double* d_u, const double* d_E, double E, double nu, int min_degree,
                                        double* d_stress, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Rule r;
  FES_TRY(rule_for(elem_kind, min_degree, r));
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el, kBlock);
  if (elem_kind == FES_P1_TRI)
    k_stress_field<FES_P1_TRI, 2><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_stress);
  else if (elem_kind == FES_P1_TET)
    k_stress_field<FES_P1_TET, 3><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_stress);
  else if (elem_kind == FES_P2_TRI)
    k_stress_field<FES_P2_TRI, 2><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_stress);
  else
    return set_error(FES_ERR_UNSUPPORTED, "no elastic constitutive matrix for element kind %d", elem_kind);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

// ---- von Mises stress and its derivative with respect to the element DOFs (plane strain) -------
// ref fem/sensitivity.py:137-199 (_VonMisesStress._voigt_stress / _Q / _dvm_du): s = D B u_e at the
// first quadrature point, sigma_zz = nu (s_xx + s_yy), vm = sqrt(Q), dvm/du_e = (D B)^T dQ/ds / (2 vm).
namespace fes {
namespace {
template <int ELEM>
__global__ void __launch_bounds__(kBlock)
k_von_mises_gradient(const int32_t* __restrict__ en, int64_t n_el, const double* __restrict__ coords,
                     const double* __restrict__ u, const double* __restrict__ d_E, double E, double nu, Rule rule,
                     double* __restrict__ vm_out, double* __restrict__ dvm_du) {
  constexpr int N = ET<ELEM>::N, SD = 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_el) return;
  double X[3][SD];
  load_corners<ELEM, SD>(en + e * N, coords, X);
  Geom<2, SD> g;
  make_geom<2, SD>(X, g);
  double d[N][2];
  dshape_at<ELEM>(rule.pts[0], d);
  double gn[N][SD];
  double exx = 0.0, eyy = 0.0, gxy = 0.0;  // Voigt strain (engineering shear)
#pragma unroll
  for (int n = 0; n < N; ++n) {
#pragma unroll
    for (int x = 0; x < SD; ++x) gn[n][x] = d[n][0] * g.Jinv[0][x] + d[n][1] * g.Jinv[1][x];
    const int64_t v = en[e * N + n];
    const double ux = __ldg(u + 2 * v), uy = __ldg(u + 2 * v + 1);
    exx += gn[n][0] * ux;
    eyy += gn[n][1] * uy;
    gxy += gn[n][1] * ux + gn[n][0] * uy;
  }
  const Material m = lame(d_E ? __ldg(d_E + e) : E, nu);
  const double sxx = (m.lam + 2.0 * m.mu) * exx + m.lam * eyy, syy = m.lam * exx + (m.lam + 2.0 * m.mu) * eyy,
               sxy = m.mu * gxy, szz = nu * (sxx + syy);
  const double Q = sxx * sxx + syy * syy + szz * szz - sxx * syy - syy * szz - szz * sxx + 3.0 * sxy * sxy;
  const double vm = sqrt(fmax(Q, 0.0));
  if (vm_out) vm_out[e] = vm;
  if (!dvm_du) return;
  double ds[3] = {0.0, 0.0, 0.0};
  if (vm > 0.0) {
    ds[0] = (2 * sxx - syy - szz + nu * (2 * szz - syy - sxx)) / (2.0 * vm);
    ds[1] = (2 * syy - sxx - szz + nu * (2 * szz - sxx - syy)) / (2.0 * vm);
    ds[2] = 6.0 * sxy / (2.0 * vm);
  }
  // (D B)^T ds: t = D ds (D symmetric), then B^T t
  const double txx = (m.lam + 2.0 * m.mu) * ds[0] + m.lam * ds[1], tyy = m.lam * ds[0] + (m.lam + 2.0 * m.mu) * ds[1],
               txy = m.mu * ds[2];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    dvm_du[(e * N + n) * 2 + 0] = gn[n][0] * txx + gn[n][1] * txy;
    dvm_du[(e * N + n) * 2 + 1] = gn[n][1] * tyy + gn[n][0] * txy;
  }
}
}  // namespace
}  // namespace fes

extern "C" int fes_von_mises_gradient(int elem_kind, const int32_t* en, int64_t n_el, const double* coords,
                                      const double* d_u, const double* d_E, double E, double nu, int min_degree,
                                      double* d_vm, double* d_dvm_du, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Rule r;
  FES_TRY(rule_for(elem_kind, min_degree, r));
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el, kBlock);
  if (elem_kind == FES_P1_TRI)
    k_von_mises_gradient<FES_P1_TRI><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_vm, d_dvm_du);
  else if (elem_kind == FES_P2_TRI)
    k_von_mises_gradient<FES_P2_TRI><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, d_E, E, nu, r, d_vm, d_dvm_du);
  else
    return set_error(FES_ERR_UNSUPPORTED, "stress quantities of interest are plane-strain 2D only (element kind %d)", elem_kind);
  FES_LAUNCH_CHECK();
  return FES_OK;
}

extern "C" int fes_scalar_gradient(int elem_kind, int sd, int min_degree, const int32_t* en, int64_t n_el,
                                   const double* coords, const double* d_u, double* d_grad, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  Rule r;
  FES_TRY(rule_for(elem_kind, min_degree, r));
  if (n_el == 0) return FES_OK;
  const int g = grid_for(n_el, kBlock);
#define FES_CALL(EL, SD_) k_scalar_gradient<EL, SD_><<<g, kBlock, 0, st>>>(en, n_el, coords, d_u, r, d_grad)
  FES_VOLUME_DISPATCH(FES_CALL);
#undef FES_CALL
  FES_LAUNCH_CHECK();
  return FES_OK;
}
