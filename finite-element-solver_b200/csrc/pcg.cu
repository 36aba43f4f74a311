// Jacobi-preconditioned CG as ONE persistent cooperative kernel per solve.
//
// Replaces Backend.prepare / LinearSolver.solve (ref fem/backends.py:62-116,186-213) and the
// Dirichlet elimination of DiscreteSystem (ref fem/system.py:26-65).  The iteration, x0 = 0,
// the stopping rule ||r|| < rtol*||b|| tested at the top of each iteration on the recurrence
// residual, and the failure semantics mirror scipy.sparse.linalg.cg(A, b, rtol, atol=0,
// maxiter, M=diag(A)^-1).
//
// Device design: the whole solve is one cooperative launch (grid = resident CTAs on all SMs).
// Each iteration is three phases separated by grid barriers: p = z + beta p | q = A p with the
// p.q partials | x += alpha p, r -= alpha q with the r.z and r.r partials.  alpha, beta and the
// convergence test are computed redundantly by every CTA from per-CTA partial sums reduced in a
// fixed order (warp shuffle tree -> shared memory -> global partials -> same tree again), so
// no scalar ever visits the host and results are bit-reproducible run to run.  A 0/1 free mask
// restricts the iteration to the free-free block without forming A[free][:, free].
#include <cooperative_groups.h>

#include "gridsum.cuh"
#include "plan.cuh"

namespace cg = cooperative_groups;

struct PcgState {
  long long iters;
  int status;  // 0 converged, 1 not converged in maxiter, 2 breakdown (p.Ap <= 0 or NaN)
  double relres;
  double bnorm;
};

struct fes_solver {
  int64_t n = 0, n_free = 0, nnz = 0;
  const int32_t* indptr = nullptr;
  const int32_t* indices = nullptr;
  const double* data = nullptr;
  const uint8_t* mask = nullptr;
  fes::DevBuf<int32_t> own_indptr, own_indices;
  fes::DevBuf<double> own_data;
  fes::DevBuf<double> dinv, r, p, q, partials, hb, hx;
  fes::DevBuf<PcgState> state;
  int lanes = 8, grid = 0, block = 0;
  int block_c = 0;  // 0 = scalar CSR rows; 2 / 3 = node-block rows
  int64_t n_nodes = 0;
  const int32_t* node_ptr = nullptr;
  const int32_t* node_adj = nullptr;
};

namespace fes {
namespace {

constexpr int kPcgBlock = 512;

struct PcgArgs {
  int64_t n;
  const int32_t* indptr;
  const int32_t* indices;
  const double* data;
  const uint8_t* mask;
  const double* dinv;
  const double* b;
  double* x;
  double* r;
  double* p;
  double* q;
  double* partials;  // [4][kMaxGrid]
  PcgState* state;
  double rtol;
  long long maxiter;
  int warm;
  // node-block mode (operators assembled from a plan): row c*v+i holds, for each neighbour node
  // w = node_adj[node_ptr[v]+s], the c values at c*c*node_ptr[v] + i*c*deg + c*s .. +c-1
  int64_t n_nodes;
  const int32_t* node_ptr;
  const int32_t* node_adj;
};

__device__ __forceinline__ bool is_free(const uint8_t* mask, int64_t i) { return mask == nullptr || mask[i] != 0; }

// Row sum over G cooperating lanes.  Every lane of the warp must call this together (the shuffle
// tree spans the warp), so inactive rows pass active = false instead of skipping the call.
template <int G, bool MASKED_X>
__device__ __forceinline__ double row_dot(const PcgArgs& A, bool active, int64_t row, int lane, const double* vec) {
  double s = 0.0;
  if (active) {
    const int32_t k1 = __ldg(A.indptr + row + 1);
    for (int32_t k = __ldg(A.indptr + row) + lane; k < k1; k += G) {
      const int32_t j = __ldg(A.indices + k);
      double xv = vec[j];
      if (MASKED_X) xv = is_free(A.mask, j) ? xv : 0.0;
      s += __ldg(A.data + k) * xv;
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}


// Node-block row product: G lanes share node v and stride over its neighbour blocks; one column
// index and one C-wide gather of `vec` serve C*C matrix values.  All lanes of the warp call this
// together.  out[i] is the full row sum for dof C*v+i (valid in every lane of the group).
template <int G, int C, bool MASKED_X>
__device__ __forceinline__ void node_row_dot(const PcgArgs& A, bool active, int64_t v, int lane, const double* vec,
                                             double (&out)[C]) {
#pragma unroll
  for (int i = 0; i < C; ++i) out[i] = 0.0;
  if (active) {
    const int32_t np0 = __ldg(A.node_ptr + v), deg = __ldg(A.node_ptr + v + 1) - np0;
    const double* base = A.data + (int64_t)C * C * np0;
    for (int32_t s = lane; s < deg; s += G) {
      const int64_t w = __ldg(A.node_adj + np0 + s);
      double pw[C];
      if constexpr (C == 2) {
        const double2 t = *reinterpret_cast<const double2*>(vec + 2 * w);
        pw[0] = t.x; pw[1] = t.y;
      } else {
#pragma unroll
        for (int j = 0; j < C; ++j) pw[j] = vec[C * w + j];
      }
      if (MASKED_X) {
#pragma unroll
        for (int j = 0; j < C; ++j) pw[j] = is_free(A.mask, C * w + j) ? pw[j] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const double* row = base + (int64_t)i * C * deg + C * s;
        if constexpr (C == 2) {
          const double2 a = __ldg(reinterpret_cast<const double2*>(row));
          out[i] = fma(a.x, pw[0], fma(a.y, pw[1], out[i]));
        } else {
#pragma unroll
          for (int j = 0; j < C; ++j) out[i] = fma(__ldg(row + j), pw[j], out[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) out[i] += __shfl_xor_sync(0xffffffffu, out[i], o);
}

// One sparse product y = A vec restricted to the free rows, with the optional dot partial
// sum_i vec2[i] * y[i].  Rows are walked in warp-convergent strides; UNROLL independent rows per
// group are in flight together.
template <int G, int C, bool MASKED_X, bool RESIDUAL>
__device__ __forceinline__ double sparse_product(const PcgArgs& A, int64_t tid, int64_t nth, const double* vec,
                                                 double* y) {
  constexpr int GPW = 32 / G, UNROLL = 2;
  const int64_t ngrp = nth / G;
  const int lane = (int)(tid % G), sub = (int)((tid / G) % GPW);
  const int64_t rbase0 = tid / G - sub;
  double partial = 0.0;
  if constexpr (C == 0) {
    for (int64_t rbase = rbase0; rbase < A.n; rbase += ngrp * UNROLL) {
      double acc[UNROLL];
      bool act[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t row = rbase + u * ngrp + sub;
        act[u] = row < A.n && is_free(A.mask, row);
        acc[u] = row_dot<G, MASKED_X>(A, act[u], row, lane, vec);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t row = rbase + u * ngrp + sub;
        if (act[u] && lane == 0) {
          if (RESIDUAL) y[row] = A.b[row] - acc[u];
          else { y[row] = acc[u]; partial += vec[row] * acc[u]; }
        }
      }
    }
  } else {
    constexpr int CC = C == 0 ? 1 : C;
    for (int64_t vbase = rbase0; vbase < A.n_nodes; vbase += ngrp * UNROLL) {
      double acc[UNROLL][CC];
      bool act[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t v = vbase + u * ngrp + sub;
        bool any = false;
        if (v < A.n_nodes) {
#pragma unroll
          for (int i = 0; i < CC; ++i) any = any || is_free(A.mask, CC * v + i);
        }
        act[u] = any;
        node_row_dot<G, CC, MASKED_X>(A, any, v, lane, vec, acc[u]);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t v = vbase + u * ngrp + sub;
        if (act[u] && lane < CC) {
          // lane i of the group finishes dof C*v+i (every lane holds all C sums after the shuffles)
          double mine = acc[u][0];
#pragma unroll
          for (int i = 1; i < CC; ++i) mine = (lane == i) ? acc[u][i] : mine;
          const int64_t row = CC * v + lane;
          if (is_free(A.mask, row)) {
            if (RESIDUAL) y[row] = A.b[row] - mine;
            else { y[row] = mine; partial += vec[row] * mine; }
          }
        }
      }
    }
  }
  return partial;
}

template <int G, int C>
__global__ void __launch_bounds__(kPcgBlock) k_pcg(PcgArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[33];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;

  // ---- r0 = b - A x0 on the free rows (x0 = 0 unless warm), p = 0 everywhere
  if (A.warm) {
    sparse_product<G, C, true, true>(A, tid, nth, A.x, A.r);
    grid.sync();
  }
  double v3[3] = {0.0, 0.0, 0.0};  // b.b, r.r, r.z
  for (int64_t i = tid; i < A.n; i += nth) {
    A.p[i] = 0.0;
    if (!is_free(A.mask, i)) continue;
    const double bi = A.b[i];
    double ri = bi;
    if (A.warm) ri = A.r[i];
    else { A.r[i] = bi; A.x[i] = 0.0; }
    v3[0] += bi * bi;
    v3[1] += ri * ri;
    v3[2] += ri * ri * A.dinv[i];
  }
  grid_sum<3>(grid, v3, A.partials, 0, sm);
  const double bb = v3[0];
  double rr = v3[1], rho = v3[2], rho_prev = 1.0;
  long long it = 0;
  int status = 0;
  if (bb == 0.0) {  // scipy returns x = 0 for a zero right-hand side
    for (int64_t i = tid; i < A.n; i += nth)
      if (is_free(A.mask, i)) A.x[i] = 0.0;
    rr = 0.0;
  } else {
    for (;;) {
      if (sqrt(rr) < A.rtol * sqrt(bb)) break;
      if (it >= A.maxiter) { status = 1; break; }
      const double beta = (it == 0) ? 0.0 : rho / rho_prev;
      for (int64_t i = tid; i < A.n; i += nth)
        if (is_free(A.mask, i)) A.p[i] = A.dinv[i] * A.r[i] + beta * A.p[i];
      grid.sync();

      double v1[1] = {sparse_product<G, C, false, false>(A, tid, nth, A.p, A.q)};
      grid_sum<1>(grid, v1, A.partials, 3, sm);
      const double pq = v1[0];
      if (!(pq > 0.0)) { status = 2; break; }
      const double alpha = rho / pq;

      double v2[2] = {0.0, 0.0};
      for (int64_t i = tid; i < A.n; i += nth) {
        if (!is_free(A.mask, i)) continue;
        A.x[i] += alpha * A.p[i];
        const double ri = A.r[i] - alpha * A.q[i];
        A.r[i] = ri;
        v2[0] += ri * ri;
        v2[1] += ri * ri * A.dinv[i];
      }
      grid_sum<2>(grid, v2, A.partials, 1, sm);
      rr = v2[0];
      rho_prev = rho;
      rho = v2[1];
      ++it;
    }
  }
  if (tid == 0) {
    A.state->iters = it;
    A.state->status = status;
    A.state->bnorm = sqrt(bb);
    A.state->relres = bb > 0.0 ? sqrt(rr) / sqrt(bb) : 0.0;
  }
}

template <int G>
__global__ void k_diag_inv(int64_t n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                           const double* __restrict__ data, const uint8_t* __restrict__ mask,
                           double* __restrict__ dinv, int* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / G;
  const int lane = (int)(t % G);
  double d = 0.0;
  if (row < n) {
    const int32_t k1 = indptr[row + 1];
    for (int32_t k = indptr[row] + lane; k < k1; k += G)
      if (indices[k] == row) d += data[k];
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (row < n && lane == 0) {
    const bool fr = mask == nullptr || mask[row] != 0;
    if (fr && !(d != 0.0)) atomicExch(bad, 1);
    dinv[row] = fr && d != 0.0 ? 1.0 / d : 0.0;
  }
}

// rhs = b - A xfixed on free rows, 0 on fixed rows (ref fem/system.py:49-50)
template <int G>
__global__ void k_lift(int64_t n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const double* __restrict__ data, const uint8_t* __restrict__ mask,
                       const double* __restrict__ b, const double* __restrict__ xf, double* __restrict__ rhs) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / G;
  const int lane = (int)(t % G);
  double s = 0.0;
  const bool active = row < n && (mask == nullptr || mask[row] != 0);
  if (active) {
    const int32_t k1 = indptr[row + 1];
    for (int32_t k = indptr[row] + lane; k < k1; k += G) s += data[k] * xf[indices[k]];
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (row < n && lane == 0) rhs[row] = active ? b[row] - s : 0.0;
}

__global__ void k_count_free(const uint8_t* __restrict__ mask, int64_t n, unsigned long long* out) {
  __shared__ double sm[33];
  double c = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c += mask[i] ? 1.0 : 0.0;
  c = block_sum(c, sm);
  if (threadIdx.x == 0) atomicAdd(out, (unsigned long long)c);
}

int pick_lanes(int64_t n, int64_t nnz) {
  const double avg = n > 0 ? (double)nnz / (double)n : 1.0;
  if (avg <= 3.0) return 2;
  if (avg <= 10.0) return 4;
  if (avg <= 24.0) return 8;
  if (avg <= 56.0) return 16;
  return 32;
}

template <int G, int C>
int launch_pcg(fes_solver* s, PcgArgs& args, cudaStream_t st) {
  if (s->grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    FES_CUDA(cudaGetDevice(&dev));
    FES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg<G, C>, kPcgBlock, 0));
    if (per_sm < 1) return set_error(FES_ERR_CUDA, "PCG kernel does not fit on an SM");
    s->grid = sms * per_sm;
    if (s->grid > kMaxGrid) s->grid = kMaxGrid;
    s->block = kPcgBlock;
  }
  void* params[] = {&args};
  FES_CUDA(cudaLaunchCooperativeKernel((void*)k_pcg<G, C>, dim3(s->grid), dim3(s->block), params, 0, st));
  count_launch();
  return FES_OK;
}

int init_solver(fes_solver* s, cudaStream_t st) {
  s->lanes = pick_lanes(s->n, s->nnz);
  FES_CUDA(s->dinv.alloc(s->n));
  FES_CUDA(s->r.alloc(s->n));
  FES_CUDA(s->p.alloc(s->n));
  FES_CUDA(s->q.alloc(s->n));
  FES_CUDA(s->partials.alloc(4 * kMaxGrid));
  FES_CUDA(s->state.alloc(1));
  s->n_free = s->n;
  if (s->mask) {
    DevBuf<unsigned long long> cnt;
    FES_CUDA(cnt.alloc(1));
    FES_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
    k_count_free<<<148, 256, 0, st>>>(s->mask, s->n, cnt.p);
    FES_LAUNCH_CHECK();
    unsigned long long h = 0;
    FES_CUDA(cudaMemcpyAsync(&h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    FES_CUDA(cudaStreamSynchronize(st));
    s->n_free = (int64_t)h;
  }
  return fes_solver_refresh(s, st);
}

}  // namespace
}  // namespace fes

using namespace fes;

#define FES_BY_LANES(lanes, CALL)                \
  switch (lanes) {                               \
    case 2: CALL(2); break;                      \
    case 4: CALL(4); break;                      \
    case 8: CALL(8); break;                      \
    case 16: CALL(16); break;                    \
    default: CALL(32); break;                    \
  }

extern "C" int fes_solver_refresh(fes_solver* s, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!s) return set_error(FES_ERR_VALUE, "fes_solver_refresh: NULL solver");
  if (s->n == 0) return FES_OK;
  DevBuf<int> bad;
  FES_CUDA(bad.alloc(1));
  FES_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
#define FES_DIAG(G) \
  k_diag_inv<G><<<grid_for(s->n * G, 256), 256, 0, st>>>(s->n, s->indptr, s->indices, s->data, s->mask, s->dinv.p, bad.p)
  FES_BY_LANES(s->lanes, FES_DIAG)
#undef FES_DIAG
  FES_LAUNCH_CHECK();
  int h_bad = 0;
  FES_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  if (h_bad) return set_error(FES_ERR_CG, "CG failed: zero diagonal entry, Jacobi preconditioner undefined");
  return FES_OK;
}

extern "C" int fes_solver_create(int64_t n, const int32_t* d_indptr, const int32_t* d_indices, const double* d_data,
                                 const uint8_t* d_free_mask, void* stream, fes_solver** out) {
  if (!out) return set_error(FES_ERR_VALUE, "fes_solver_create: out is NULL");
  *out = nullptr;
  if (n < 0) return set_error(FES_ERR_VALUE, "fes_solver_create: n < 0");
  fes_solver* s = new fes_solver();
  s->n = n; s->indptr = d_indptr; s->indices = d_indices; s->data = d_data; s->mask = d_free_mask;
  int32_t last = 0;
  if (n > 0 && cudaMemcpy(&last, d_indptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
    delete s;
    return set_error(FES_ERR_CUDA, "fes_solver_create: cannot read indptr");
  }
  s->nnz = last;
  const int rc = init_solver(s, (cudaStream_t)stream);
  if (rc != FES_OK) { delete s; return rc; }
  *out = s;
  return FES_OK;
}

extern "C" int fes_solver_create_host(int64_t n, const int64_t* h_indptr, const int64_t* h_indices, const double* h_data,
                                      fes_solver** out) {
  if (!out) return set_error(FES_ERR_VALUE, "fes_solver_create_host: out is NULL");
  *out = nullptr;
  if (n < 0 || !h_indptr) return set_error(FES_ERR_VALUE, "fes_solver_create_host: bad arguments");
  const int64_t nnz = h_indptr[n];
  if (nnz >= (int64_t)1 << 31) return set_error(FES_ERR_VALUE, "fes_solver_create_host: nnz exceeds int32");
  fes_solver* s = new fes_solver();
  struct Guard { fes_solver* p; ~Guard() { delete p; } } guard{s};
  s->n = n; s->nnz = nnz;
  FES_CUDA(s->own_indptr.alloc(n + 1));
  FES_CUDA(s->own_indices.alloc(nnz));
  FES_CUDA(s->own_data.alloc(nnz));
  {  // narrow the reference's int64 index arrays on the host, then one copy each
    int32_t* tmp = new int32_t[(size_t)(nnz > n + 1 ? nnz : n + 1)];
    for (int64_t i = 0; i <= n; ++i) tmp[i] = (int32_t)h_indptr[i];
    cudaError_t e = cudaMemcpy(s->own_indptr.p, tmp, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice);
    for (int64_t i = 0; i < nnz; ++i) tmp[i] = (int32_t)h_indices[i];
    if (e == cudaSuccess && nnz) e = cudaMemcpy(s->own_indices.p, tmp, nnz * sizeof(int32_t), cudaMemcpyHostToDevice);
    delete[] tmp;
    FES_CUDA(e);
  }
  if (nnz) FES_CUDA(cudaMemcpy(s->own_data.p, h_data, nnz * sizeof(double), cudaMemcpyHostToDevice));
  s->indptr = s->own_indptr.p; s->indices = s->own_indices.p; s->data = s->own_data.p; s->mask = nullptr;
  FES_TRY(init_solver(s, nullptr));
  guard.p = nullptr;
  *out = s;
  return FES_OK;
}

extern "C" void fes_solver_destroy(fes_solver* s) { delete s; }

extern "C" int fes_solver_use_plan(fes_solver* s, const fes_plan* plan) {
  if (!s || !plan) return set_error(FES_ERR_VALUE, "fes_solver_use_plan: NULL argument");
  if (plan->n_dofs != s->n || s->indptr != plan->indptr.p)
    return set_error(FES_ERR_VALUE, "fes_solver_use_plan: the operator was not assembled from this plan");
  if (plan->c == 2 || plan->c == 3) {
    s->block_c = plan->c;
    s->n_nodes = plan->n_nodes;
    s->node_ptr = plan->node_ptr.p;
    s->node_adj = plan->node_adj.p;
    s->grid = 0;  // occupancy of the block-mode kernel differs
  }
  return FES_OK;
}

extern "C" int fes_solver_solve(fes_solver* s, const double* d_b, double* d_x, double rtol, int64_t maxiter,
                                int warm_start, int64_t* iters_out, double* relres_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!s) return set_error(FES_ERR_VALUE, "fes_solver_solve: NULL solver");
  if (iters_out) *iters_out = 0;
  if (relres_out) *relres_out = 0.0;
  if (s->n == 0 || s->n_free == 0) return FES_OK;
  if (maxiter <= 0) maxiter = 10 * s->n_free;
  PcgArgs a{s->n, s->indptr, s->indices, s->data, s->mask, s->dinv.p, d_b, d_x, s->r.p, s->p.p, s->q.p,
            s->partials.p, s->state.p, rtol, (long long)maxiter, warm_start ? 1 : 0, s->n_nodes, s->node_ptr,
            s->node_adj};
  if (s->block_c == 2) {
    FES_TRY((launch_pcg<8, 2>(s, a, st)));
  } else if (s->block_c == 3) {
    FES_TRY((launch_pcg<16, 3>(s, a, st)));
  } else {
#define FES_PCG(G) FES_TRY((launch_pcg<G, 0>(s, a, st)))
    FES_BY_LANES(s->lanes, FES_PCG)
#undef FES_PCG
  }
  PcgState h;
  FES_CUDA(cudaMemcpyAsync(&h, s->state.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  if (iters_out) *iters_out = h.iters;
  if (relres_out) *relres_out = h.relres;
  if (h.status == 1)
    return set_error(FES_ERR_CG, "CG failed to solve the free-free block: did not converge in %lld iterations "
                     "(relative residual %.3e). The operator must be symmetric positive-definite for CG.",
                     (long long)h.iters, h.relres);
  if (h.status == 2)
    return set_error(FES_ERR_CG, "CG failed to solve the free-free block: illegal input or breakdown "
                     "(p.Ap <= 0 at iteration %lld). The operator must be symmetric positive-definite for CG.",
                     (long long)h.iters);
  return FES_OK;
}

extern "C" int fes_solver_solve_host(fes_solver* s, const double* h_b, double* h_x, double rtol, int64_t maxiter,
                                     int64_t* iters_out, double* relres_out) {
  if (!s || !h_b || !h_x) return set_error(FES_ERR_VALUE, "fes_solver_solve_host: NULL argument");
  if (s->hb.n != (size_t)s->n) { FES_CUDA(s->hb.alloc(s->n)); FES_CUDA(s->hx.alloc(s->n)); }
  if (s->n) FES_CUDA(cudaMemcpy(s->hb.p, h_b, s->n * sizeof(double), cudaMemcpyHostToDevice));
  const int rc = fes_solver_solve(s, s->hb.p, s->hx.p, rtol, maxiter, 0, iters_out, relres_out, nullptr);
  if (rc != FES_OK) return rc;
  if (s->n) FES_CUDA(cudaMemcpy(h_x, s->hx.p, s->n * sizeof(double), cudaMemcpyDeviceToHost));
  return FES_OK;
}

extern "C" int fes_lift_rhs(const fes_solver* s, const double* d_b, const double* d_xfixed, double* d_rhs,
                            void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!s) return set_error(FES_ERR_VALUE, "fes_lift_rhs: NULL solver");
  if (s->n == 0) return FES_OK;
#define FES_LIFT(G) \
  k_lift<G><<<grid_for(s->n * G, 256), 256, 0, st>>>(s->n, s->indptr, s->indices, s->data, s->mask, d_b, d_xfixed, d_rhs)
  FES_BY_LANES(s->lanes, FES_LIFT)
#undef FES_LIFT
  FES_LAUNCH_CHECK();
  return FES_OK;
}
