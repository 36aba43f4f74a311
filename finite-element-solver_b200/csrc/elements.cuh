// Element math shared by every kernel: affine geometry (ref fem/elements.py:174-212), shape
// gradients (ref fem/elements.py:258-297, 379-396), quadrature tables (ref fem/quadrature.py:
// 54-104), reference mass matrices (ref fem/elements.py:113-126, 267-274) and the (a,b) node
// block of each bilinear form (ref fem/forms.py:202-209, 254-258, 323-333).
//
// Elasticity uses the isotropic closed form of B^T D B:
//   K[(a,i),(b,j)] = sum_q w_q ( lam g_ai g_bj + mu g_aj g_bi + mu (g_a.g_b) delta_ij )
// which is the Voigt contraction of fem/forms.py:37-71 with fem/materials.py:38-66 written out.
#pragma once
#include "common.cuh"

namespace fes {

template <int ELEM> struct ET;
template <> struct ET<FES_P1_LINE> { static constexpr int N = 2, RD = 1, NQ = 1; static constexpr bool P2 = false; };
template <> struct ET<FES_P1_TRI>  { static constexpr int N = 3, RD = 2, NQ = 1; static constexpr bool P2 = false; };
template <> struct ET<FES_P1_TET>  { static constexpr int N = 4, RD = 3, NQ = 1; static constexpr bool P2 = false; };
template <> struct ET<FES_P2_LINE> { static constexpr int N = 3, RD = 1, NQ = 2; static constexpr bool P2 = true; };
template <> struct ET<FES_P2_TRI>  { static constexpr int N = 6, RD = 2, NQ = 3; static constexpr bool P2 = true; };

__host__ __device__ inline int elem_nodes_of(int kind) {
  return kind == FES_P1_LINE ? 2 : kind == FES_P1_TRI ? 3 : kind == FES_P1_TET ? 4 : kind == FES_P2_LINE ? 3 : 6;
}
__host__ __device__ inline int elem_ref_dim(int kind) {
  return (kind == FES_P1_LINE || kind == FES_P2_LINE) ? 1 : (kind == FES_P1_TET ? 3 : 2);
}

// P2 triangle reference gradients at the Strang points (1/6,1/6), (2/3,1/6), (1/6,2/3);
// node order c0,c1,c2,m12,m02,m01.  L = (l0,l1,l2) barycentrics at the point.
#define FES_P2G(L0, L1, L2)                                                                   \
  { {-(4.0 * (L0) - 1.0), -(4.0 * (L0) - 1.0)}, {4.0 * (L1) - 1.0, 0.0}, {0.0, 4.0 * (L2) - 1.0}, \
    {4.0 * (L2), 4.0 * (L1)}, {-4.0 * (L2), 4.0 * (L0) - 4.0 * (L2)}, {4.0 * (L0) - 4.0 * (L1), -4.0 * (L1)} }
static __constant__ double c_p2tri_dshape[3][6][2] = {
    FES_P2G(1.0 - 1.0 / 6 - 1.0 / 6, 1.0 / 6, 1.0 / 6),
    FES_P2G(1.0 - 2.0 / 3 - 1.0 / 6, 2.0 / 3, 1.0 / 6),
    FES_P2G(1.0 - 1.0 / 6 - 2.0 / 3, 1.0 / 6, 2.0 / 3)};
// The same table as a constexpr function, and its quadrature-summed products
//   M_ab[r][s] = sum_q dshape[q][a][r] * dshape[q][b][s]
// (the three Strang points carry equal weights).  On an affine element the physical gradient of
// shape function a at point q is dshape[q][a][0] G_0 + dshape[q][a][1] G_1 with G_r the rows of the
// inverse Jacobian, so  sum_q grad(phi_a) (x) grad(phi_b) = sum_rs M_ab[r][s] G_r (x) G_s: the tiled
// assembly kernel stores only G_0, G_1 per element and contracts with this 2 x 2 table per node pair.
constexpr double p2tri_dshape(int q, int a, int r) {
  const double L1 = (q == 1) ? 2.0 / 3 : 1.0 / 6, L2 = (q == 2) ? 2.0 / 3 : 1.0 / 6, L0 = 1.0 - L1 - L2;
  switch (a * 2 + r) {
    case 0: case 1: return -(4.0 * L0 - 1.0);
    case 2: return 4.0 * L1 - 1.0;
    case 3: return 0.0;
    case 4: return 0.0;
    case 5: return 4.0 * L2 - 1.0;
    case 6: return 4.0 * L2;
    case 7: return 4.0 * L1;
    case 8: return -4.0 * L2;
    case 9: return 4.0 * L0 - 4.0 * L2;
    case 10: return 4.0 * L0 - 4.0 * L1;
    default: return -4.0 * L1;
  }
}
struct P2TriPairTable {
  double m[36][4];  // [a * 6 + b][r * 2 + s]
};
constexpr P2TriPairTable make_p2tri_pair_table() {
  P2TriPairTable t{};
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b)
      for (int r = 0; r < 2; ++r)
        for (int s = 0; s < 2; ++s) {
          double acc = 0.0;
          for (int q = 0; q < 3; ++q) acc += p2tri_dshape(q, a, r) * p2tri_dshape(q, b, s);
          t.m[a * 6 + b][r * 2 + s] = acc;
        }
  return t;
}
static __constant__ P2TriPairTable c_p2tri_pair = make_p2tri_pair_table();
// Mass per unit measure: P2 triangle (k/180) and P2 line (k/30), exact rationals; the
// divisions are folded at compile time so no kernel divides per contribution.
#define FES_M180(a, b, c, d, e, f) {(a) / 180.0, (b) / 180.0, (c) / 180.0, (d) / 180.0, (e) / 180.0, (f) / 180.0}
static __constant__ double c_p2tri_mass[6][6] = {
    FES_M180(6, -1, -1, -4, 0, 0), FES_M180(-1, 6, -1, 0, -4, 0), FES_M180(-1, -1, 6, 0, 0, -4),
    FES_M180(-4, 0, 0, 32, 16, 16), FES_M180(0, -4, 0, 16, 32, 16), FES_M180(0, 0, -4, 16, 16, 32)};
static __constant__ double c_p2line_mass[3][3] = {
    {4 / 30.0, -1 / 30.0, 2 / 30.0}, {-1 / 30.0, 4 / 30.0, 2 / 30.0}, {2 / 30.0, 2 / 30.0, 16 / 30.0}};

// Inverse Jacobian rows (d xi_r / d x_s) and the measure scale |det J| (or sqrt det J^T J for
// an embedded facet, where only the scale is defined here: boundary integrals are mass-type).
template <int RD, int SD>
struct Geom {
  double Jinv[RD][SD];
  double scale;
};

template <int RD, int SD>
__device__ __forceinline__ void make_geom(const double (&X)[RD + 1][SD], Geom<RD, SD>& g) {
  if constexpr (RD == 1 && SD == 1) {
    const double j = X[1][0] - X[0][0];
    g.Jinv[0][0] = 1.0 / j;
    g.scale = fabs(j);
  } else if constexpr (RD == 2 && SD == 2) {
    const double a = X[1][0] - X[0][0], b = X[2][0] - X[0][0];  // J[0][0], J[0][1]
    const double c = X[1][1] - X[0][1], d = X[2][1] - X[0][1];  // J[1][0], J[1][1]
    const double det = a * d - b * c, inv = 1.0 / det;
    g.Jinv[0][0] = d * inv;  g.Jinv[0][1] = -b * inv;
    g.Jinv[1][0] = -c * inv; g.Jinv[1][1] = a * inv;
    g.scale = fabs(det);
  } else if constexpr (RD == 3 && SD == 3) {
    double J[3][3];  // J[s][r] = X[r+1][s] - X[0][s]
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int r = 0; r < 3; ++r) J[s][r] = X[r + 1][s] - X[0][s];
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, inv = 1.0 / det;
    // Jinv[r][s] = cofactor(J)[s][r] / det
    g.Jinv[0][0] = c00 * inv;
    g.Jinv[1][0] = c01 * inv;
    g.Jinv[2][0] = c02 * inv;
    g.Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * inv;
    g.Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv;
    g.Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * inv;
    g.Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv;
    g.Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * inv;
    g.Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv;
    g.scale = fabs(det);
  } else if constexpr (RD == 1) {  // line embedded in 2D/3D: length
    double s2 = 0.0;
#pragma unroll
    for (int s = 0; s < SD; ++s) { const double d = X[1][s] - X[0][s]; s2 += d * d; }
    g.scale = sqrt(s2);
  } else {  // triangle embedded in 3D: sqrt(det(J^T J)) = |a x b|  (ref fem/elements.py:194-203)
    double aa = 0, ab = 0, bb = 0;
#pragma unroll
    for (int s = 0; s < SD; ++s) {
      const double a = X[1][s] - X[0][s], b = X[2][s] - X[0][s];
      aa += a * a; ab += a * b; bb += b * b;
    }
    g.scale = sqrt(fabs(aa * bb - ab * ab));
  }
}

// Cofactor form of the same geometry for volume elements (RD == SD): cof = det * Jinv, no
// division.  The tiled assembly kernel scales it by one rsqrt instead of a divide and a sqrt.
template <int RD, int SD>
__device__ __forceinline__ void make_cof(const double (&X)[RD + 1][SD], double (&cof)[RD][SD], double& det) {
  static_assert(RD == SD, "cofactor geometry is for volume elements");
  if constexpr (RD == 1) {
    det = X[1][0] - X[0][0];
    cof[0][0] = 1.0;
  } else if constexpr (RD == 2) {
    const double a = X[1][0] - X[0][0], b = X[2][0] - X[0][0];
    const double c = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
    det = a * d - b * c;
    cof[0][0] = d;  cof[0][1] = -b;
    cof[1][0] = -c; cof[1][1] = a;
  } else {
    double J[3][3];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int r = 0; r < 3; ++r) J[s][r] = X[r + 1][s] - X[0][s];
    cof[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    cof[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    cof[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * cof[0][0] + J[0][1] * cof[1][0] + J[0][2] * cof[2][0];
    cof[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    cof[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    cof[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    cof[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    cof[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    cof[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  }
}

template <int ELEM, int SD>
__device__ __forceinline__ void load_geom(const int32_t* __restrict__ en, const double* __restrict__ coords,
                                          Geom<ET<ELEM>::RD, SD>& g) {
  constexpr int RD = ET<ELEM>::RD;
  double X[RD + 1][SD];
#pragma unroll
  for (int n = 0; n <= RD; ++n) {
    const int64_t v = en[n];  // corners come first in every element's node list
#pragma unroll
    for (int s = 0; s < SD; ++s) X[n][s] = __ldg(coords + v * SD + s);
  }
  make_geom<RD, SD>(X, g);
}

__device__ __forceinline__ double ref_measure(int rd) { return rd == 1 ? 1.0 : (rd == 2 ? 0.5 : 1.0 / 6.0); }

// physical gradient of shape function `a` at quadrature point q, from the rows of the inverse
// Jacobian (or of any multiple of it: the map is linear)
template <int ELEM, int SD>
__device__ __forceinline__ void shape_grad_rows(const double (&Jinv)[ET<ELEM>::RD][SD], int q, int a, double (&out)[SD]) {
  constexpr int RD = ET<ELEM>::RD;
  if constexpr (!ET<ELEM>::P2) {
#pragma unroll
    for (int s = 0; s < SD; ++s) {
      double first = 0.0, pick = 0.0;
#pragma unroll
      for (int r = 0; r < RD; ++r) {
        first -= Jinv[r][s];
        pick = (a == r + 1) ? Jinv[r][s] : pick;
      }
      out[s] = (a == 0) ? first : pick;
    }
  } else {
    static_assert(ELEM == FES_P2_TRI, "gradient forms on P2 are implemented for the triangle only");
#pragma unroll
    for (int s = 0; s < SD; ++s)
      out[s] = c_p2tri_dshape[q][a][0] * Jinv[0][s] + c_p2tri_dshape[q][a][1] * Jinv[1][s];
  }
}

template <int ELEM, int SD>
__device__ __forceinline__ void shape_grad(const Geom<ET<ELEM>::RD, SD>& g, int q, int a, double (&out)[SD]) {
  shape_grad_rows<ELEM, SD>(g.Jinv, q, a, out);
}

template <int ELEM>
__device__ __forceinline__ double mass_ref(int a, int b) {
  constexpr int N = ET<ELEM>::N;
  if constexpr (!ET<ELEM>::P2) return (a == b) ? 2.0 / (double)(N * (N + 1)) : 1.0 / (double)(N * (N + 1));
  else if constexpr (ELEM == FES_P2_TRI) return c_p2tri_mass[a][b];
  else return c_p2line_mass[a][b];
}

struct Material {
  double mu, lam;
};
__device__ __forceinline__ Material lame(double E, double nu) {
  return {E / (2.0 * (1.0 + nu)), E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))};
}

// The C x C block coupling local nodes a and b of one element, times `w` (global factor x
// per-element scale).  SD == RD for LAPLACIAN / ELASTIC.
template <int ELEM, int FORM, int C, int SD>
__device__ __forceinline__ void eval_block(const Geom<ET<ELEM>::RD, SD>& g, int a, int b, Material m,
                                           double w, double (&blk)[C][C]) {
  constexpr int RD = ET<ELEM>::RD;
  if constexpr (FORM == FES_FORM_MASS) {
    const double val = w * (g.scale * ref_measure(RD)) * mass_ref<ELEM>(a, b);
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) blk[i][j] = (i == j) ? val : 0.0;
  } else {
    static_assert(SD == RD, "gradient forms need a volume element");
    constexpr int NQ = ET<ELEM>::NQ;
    const double wq = w * g.scale * ref_measure(RD) / (double)NQ;  // equal weights in every rule used
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) blk[i][j] = 0.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double ga[SD], gb[SD];
      shape_grad<ELEM, SD>(g, q, a, ga);
      shape_grad<ELEM, SD>(g, q, b, gb);
      double dot = 0.0;
#pragma unroll
      for (int s = 0; s < SD; ++s) dot += ga[s] * gb[s];
      if constexpr (FORM == FES_FORM_LAPLACIAN) {
        static_assert(C == 1, "Laplacian is scalar");
        blk[0][0] += wq * dot;
      } else {
        static_assert(C == SD, "elasticity has one component per spatial dimension");
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
          for (int j = 0; j < C; ++j) {
            double t = m.lam * ga[i] * gb[j] + m.mu * ga[j] * gb[i];
            if (i == j) t += m.mu * dot;
            blk[i][j] += wq * t;
          }
      }
    }
  }
}

}  // namespace fes
