// Jacobi-preconditioned CG as ONE persistent cooperative kernel per solve.
//
// Replaces Backend.prepare / LinearSolver.solve (ref fem/backends.py:62-116,186-213) and the
// Dirichlet elimination of DiscreteSystem (ref fem/system.py:26-65).  The iteration, x0 = 0,
// the stopping rule ||r|| < rtol*||b|| tested at the top of each iteration on the recurrence
// residual, and the failure semantics mirror scipy.sparse.linalg.cg(A, b, rtol, atol=0,
// maxiter, M=diag(A)^-1).
//
// Device design
//   * The whole solve is one cooperative launch (grid = resident CTAs on all SMs).  Each
//     iteration is three phases separated by grid barriers: p = z + beta p | q = A p with the p.q
//     partials | x += alpha p, r -= alpha q with the r.z and r.r partials.  alpha, beta and the
//     convergence test are computed redundantly by every CTA from per-CTA partial sums reduced
//     in a fixed order (warp shuffle tree -> shared memory -> global partials -> same tree), so no
//     scalar visits the host and results are bit-reproducible run to run.
//   * The operator is re-laid out once per refresh into SELL-32-sigma over BLOCK rows: a node's
//     c dofs form one block row whose entries are c x c blocks sharing one column index.  A thread
//     owns one block row; the 32 lanes of a warp read one slice column by column, so every load
//     is lane-contiguous and all of a row's loads are independent (no shuffles, no dependent
//     index chain beyond column -> gather).  Rows are sorted by length inside windows of 256 so
//     slices carry almost no padding.
//   * A 0/1 free mask restricts the iteration to the free-free block without forming
//     A[free][:, free]: fixed rows carry dinv = r = p = q = 0, so the vector phases need no mask.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "gridsum.cuh"
#include "plan.cuh"

namespace cg = cooperative_groups;

struct PcgState {
  long long iters;
  int status;  // 0 converged, 1 not converged in maxiter, 2 breakdown (p.Ap <= 0 or NaN)
  double relres;
  double bnorm;
  long long cycles[6];  // thread 0 with FES_PCG_TIMERS: p update / barrier, halo, product, product reduce, x/r update, update reduce (+ loop top)
};

// Peer-memory communicator: one cudaMalloc'd region per rank, exported over CUDA IPC.
//   [0,128)      u64 halo_flag[16]   peer r stores its halo epoch here after writing my ghosts
//   [256,8448)   {double value; u64 step} red_box[4][16][8]   ring of reduction mailboxes, [step & 3][src rank][value]:
//                every value travels with the step it belongs to in ONE 16-byte store, so a reader needs no
//                flag and the writer no fence -- a reduction costs one NVLink one-way latency, not 1.5 round trips
//   [8448,8464)  u64 counters[2]     this rank's halo epoch / reduction step (persist across solves)
//   [16384,...)  double p[max_dofs]  the CG direction vector; peers write its ghost entries
constexpr int kCommMaxWorld = 16, kBoxVals = 8;
constexpr size_t kCommHaloFlag = 0, kCommRedBox = 256, kCommCounters = 8448, kCommVec = 16384;

struct fes_comm {
  int rank = 0, world = 1;
  int64_t max_dofs = 0;
  size_t bytes = 0;
  unsigned char* base = nullptr;                 // this rank's region
  unsigned char* peer[kCommMaxWorld] = {};       // every rank's region in this process' address space
  fes::DevBuf<unsigned char*> d_peer;            // the same table on the device
  bool connected = false;
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
  int lanes = 8, grid = 0, block = 0, grid_fused = -1;
  // block-row source layout: c = 1 means plain CSR rows (row_ptr = indptr, cols = indices)
  int c = 1;
  int64_t n_brow = 0;
  const int32_t* row_ptr = nullptr;
  const int32_t* row_col = nullptr;
  // SELL-32-sigma copy of the operator
  int64_t n_slices = 0, sell_entries = 0;
  fes::DevBuf<int32_t> sell_row, sell_slice_ptr, sell_cols;
  fes::DevBuf<double> sell_vals;
  // distributed mode: rows [row_begin, row_end) are owned; the rest of the local vector is ghost
  fes_comm* comm = nullptr;
  int64_t row_begin = 0, row_end = 0;
  fes::DevBuf<int32_t> send_ptr, send_src, send_dst;
  unsigned recv_mask = 0, send_mask = 0;
  int64_t n_free_global = 0;          // free dofs over all ranks: the default iteration cap must not depend on the rank
  int64_t n_slots_interior = 0;       // SELL slots [0, n_slots_interior) multiply rows that touch no ghost column
  int n_ship = 0;
  fes::DevBuf<int32_t> ship_dof, ship_ptr, ship_rank, ship_dst;
  fes::DevBuf<uint8_t> shipped;
};

namespace fes {
namespace {

constexpr int kPcgBlock = 512;
constexpr int kSigma = 256;  // rows sorted by length inside windows of this many block rows

struct PcgArgs {
  int64_t n;
  const uint8_t* mask;
  const double* dinv;
  const double* b;
  double* x;
  double* r;
  double* p;
  double* q;
  double* partials;  // [8][kMaxGrid]
  PcgState* state;
  double rtol;
  long long maxiter;
  int warm;
  int64_t n_slots;  // n_slices * 32
  const int32_t* sell_row;
  const int32_t* sell_slice_ptr;
  const int32_t* sell_cols;
  const double* sell_vals;
  // owned row range (the whole vector on one GPU) and the peer-memory communicator
  int64_t row_begin, row_end;
  int world, rank;
  unsigned char* comm;               // this rank's region (NULL on one GPU)
  unsigned char* const* comm_peer;   // device table of every rank's region
  const int32_t* send_ptr;           // [world+1] into send_src / send_dst (dof level)
  const int32_t* send_src;           // local dof to read
  const int32_t* send_dst;           // dof index in the receiver's local numbering
  unsigned recv_mask, send_mask;     // bit r: this rank receives from / sends to rank r
  int fused;                         // one reduction + two grid barriers per iteration (see k_pcg)
  int overlap;                       // distributed three-phase loop: halo shipped inside the p update, interior rows first
  int timers;                        // FES_PCG_TIMERS: per-phase cycle counters in state->cycles
  int64_t n_slots_interior;          // slots below this bound read no ghost entry (distributed overlap)
  // overlapped loop: the owned dofs some peer needs, each ONCE, with its (rank, ghost index) targets; the thread
  // that updates such a dof in the vector phase also stores it into the peers' vectors
  int n_ship;
  const int32_t* ship_dof;           // [n_ship] local dof
  const int32_t* ship_ptr;           // [n_ship + 1] into ship_rank / ship_dst
  const int32_t* ship_rank;
  const int32_t* ship_dst;
  const uint8_t* shipped;            // [n] 1 = the dof is in ship_dof (the plain vector loop skips it)
};

constexpr long long kSpinLimit = 40000000000LL;  // ~20 s of SM clocks: a lost peer traps instead of hanging

__device__ __forceinline__ void spin_until(const volatile unsigned long long* flag, unsigned long long want) {
  const long long t0 = clock64();
  while (*flag < want) {
    if (clock64() - t0 > kSpinLimit) {
      printf("fes_b200: peer wait timed out (flag %llu, want %llu)\n", (unsigned long long)*flag, want);
      __trap();
    }
  }
}

// Write this rank's owned boundary values of `vec` straight into the ghost entries of every
// neighbour's vector (NVLink peer stores), publish the epoch, then wait for the neighbours'.
__device__ __forceinline__ void halo_exchange(const PcgArgs& A, cg::grid_group& grid, const double* vec,
                                              unsigned long long epoch, int64_t tid, int64_t nth) {
  const int32_t total = A.send_ptr[A.world];
  for (int64_t k = tid; k < total; k += nth) {
    int r = 0;
    while (k >= A.send_ptr[r + 1]) ++r;
    double* peer_vec = reinterpret_cast<double*>(A.comm_peer[r] + kCommVec);
    peer_vec[A.send_dst[k]] = vec[A.send_src[k]];
  }
  __threadfence_system();
  grid.sync();
  if (blockIdx.x == 0 && (int)threadIdx.x < A.world && ((A.send_mask >> threadIdx.x) & 1u)) {
    volatile unsigned long long* f =
        reinterpret_cast<volatile unsigned long long*>(A.comm_peer[threadIdx.x] + kCommHaloFlag) + A.rank;
    *f = epoch;
  }
  if (threadIdx.x == 0) {
    for (int r = 0; r < A.world; ++r)
      if ((A.recv_mask >> r) & 1u)
        spin_until(reinterpret_cast<volatile unsigned long long*>(A.comm + kCommHaloFlag) + r, epoch);
  }
  __syncthreads();
  __threadfence_system();  // drop stale L1 lines of the ghost entries before the gathers
}

// Overlapped loop.  The boundary values of p travel INSIDE the vector phase: the thread that computes the new
// p of a dof some peer needs stores it straight into the peers' vectors (ship), fences, and runs into the grid
// barrier that ends the iteration; one thread then publishes the epoch (halo_publish) and nobody waits --
// the rows that read ghost entries call halo_wait just before they are multiplied, after the interior rows,
// so the NVLink latency hides under the interior product.  Ghost entries are read through L2 (ld.cg) in
// those rows: no L1 invalidation is needed on the receiving side.
__device__ __forceinline__ void ship(const PcgArgs& A, int k, double value) {
  for (int32_t t = A.ship_ptr[k]; t < A.ship_ptr[k + 1]; ++t)
    reinterpret_cast<double*>(A.comm_peer[A.ship_rank[t]] + kCommVec)[A.ship_dst[t]] = value;
}

__device__ __forceinline__ void halo_publish(const PcgArgs& A, unsigned long long epoch) {
  // lanes of the LAST warp of CTA 0 publish (thread 0 of every CTA is the one that waits)
  const int r = (int)threadIdx.x - ((int)blockDim.x - 32);
  if (blockIdx.x == 0 && r >= 0 && r < A.world && ((A.send_mask >> r) & 1u)) {
    *(reinterpret_cast<volatile unsigned long long*>(A.comm_peer[r] + kCommHaloFlag) + A.rank) = epoch;
    __threadfence_system();  // push the flag out now instead of whenever the write buffer drains
  }
}

__device__ __forceinline__ void halo_wait(const PcgArgs& A, unsigned long long epoch) {
  if (threadIdx.x == 0) {
    for (int r = 0; r < A.world; ++r)
      if ((A.recv_mask >> r) & 1u)
        spin_until(reinterpret_cast<volatile unsigned long long*>(A.comm + kCommHaloFlag) + r, epoch);
    __threadfence_system();
  }
  __syncthreads();
}

// Sum NV values over all ranks through the peers' mailboxes; fixed rank order, so every rank and
// every CTA computes the same bits.  v holds this rank's totals on entry (valid in every thread).
__device__ __forceinline__ void st_tagged(unsigned char* slot, double value, unsigned long long step) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"((unsigned long long)__double_as_longlong(value)),
               "l"(step)
               : "memory");
}
__device__ __forceinline__ double ld_tagged(const unsigned char* slot, unsigned long long step) {
  unsigned long long bits, tag;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tag) : "l"(slot) : "memory");
    if (tag == step) break;
    if (clock64() - t0 > kSpinLimit) {
      printf("fes_b200: peer reduction timed out (tag %llu, want %llu)\n", tag, step);
      __trap();
    }
  }
  return __longlong_as_double((long long)bits);
}

template <int NV>
__device__ __forceinline__ void rank_sum(const PcgArgs& A, double (&v)[NV], unsigned long long step, double* sm) {
  static_assert(NV <= kBoxVals, "reduction mailbox too small");
  const size_t ring = ((step & 3) * kCommMaxWorld) * kBoxVals * 16;
  if (blockIdx.x == 0 && (int)threadIdx.x < A.world * NV) {  // thread (r, k): value k into rank r's mailbox for this rank
    const int r = (int)threadIdx.x / NV, k = (int)threadIdx.x - r * NV;
    double mine = v[0];
#pragma unroll
    for (int j = 1; j < NV; ++j) mine = (k == j) ? v[j] : mine;
    st_tagged(A.comm_peer[r] + kCommRedBox + ring + ((size_t)A.rank * kBoxVals + k) * 16, mine, step);
  }
  __syncthreads();  // sm may still be read by the block reduction that produced v
  if ((int)threadIdx.x < NV) {  // thread k sums value k over the ranks in rank order: the same bits everywhere
    double t = 0.0;
    for (int r = 0; r < A.world; ++r)
      t += ld_tagged(A.comm + kCommRedBox + ring + ((size_t)r * kBoxVals + threadIdx.x) * 16, step);
    sm[threadIdx.x] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = sm[k];
}

__device__ __forceinline__ bool is_free(const uint8_t* mask, int64_t i) { return mask == nullptr || mask[i] != 0; }

// y = A vec on the free rows (0 on fixed rows); returns this thread's share of vec . y.
// RESIDUAL writes b - A vec instead.  MASKED_X treats vec as zero on fixed dofs (warm start).
// Slots [slot0, slot1) only; with DOTS the six extra sums are ADDED to dots[0..5] (the overlapped loop
// calls this twice per iteration: interior rows, then the rows that read ghost entries).
template <int C, bool MASKED_X, bool RESIDUAL, bool DOTS = false, bool VEC_L2 = false>
__device__ __forceinline__ double sparse_product(const PcgArgs& A, int64_t tid, int64_t nth, const double* vec,
                                                 double* y, double* dots = nullptr, int64_t slot0 = 0,
                                                 int64_t slot1 = -1) {
  double partial = 0.0;
  [[maybe_unused]] double d_rho = 0.0, d_s1 = 0.0, d_s2 = 0.0, d_rr = 0.0, d_t1 = 0.0, d_t2 = 0.0;
  if (slot1 < 0) slot1 = A.n_slots;
  for (int64_t slot = slot0 + tid; slot < slot1; slot += nth) {
    const int lane = (int)(slot & 31);
    const int64_t slice = slot >> 5;
    const int32_t brow = __ldcs(A.sell_row + slot);
    const int32_t base = __ldg(A.sell_slice_ptr + slice);
    const int32_t width = (__ldg(A.sell_slice_ptr + slice + 1) - base) >> 5;
    uint8_t fr[C];
#pragma unroll
    for (int i = 0; i < C; ++i) fr[i] = (brow >= 0 && is_free(A.mask, (int64_t)C * brow + i)) ? 1 : 0;
    const int32_t* cp = A.sell_cols + base + lane;
    const double* vp = A.sell_vals + (int64_t)base * (C * C) + lane;
    double acc[C];
#pragma unroll
    for (int i = 0; i < C; ++i) acc[i] = 0.0;
#pragma unroll 4
    for (int32_t s = 0; s < width; ++s) {
      const int32_t col = __ldcs(cp + s * 32);  // the operator streams through L2 (evict-first): the vectors stay resident
      double pw[C];
      if constexpr (C == 2) {
        const double2* src = reinterpret_cast<const double2*>(vec + 2 * (int64_t)col);
        const double2 t = VEC_L2 ? __ldcg(src) : *src;   // VEC_L2: ghost entries were written by a peer, read them at L2
        pw[0] = t.x; pw[1] = t.y;
      } else {
#pragma unroll
        for (int j = 0; j < C; ++j) pw[j] = VEC_L2 ? __ldcg(vec + (int64_t)C * col + j) : vec[(int64_t)C * col + j];
      }
      if (MASKED_X) {
#pragma unroll
        for (int j = 0; j < C; ++j) pw[j] = is_free(A.mask, (int64_t)C * col + j) ? pw[j] : 0.0;
      }
      const double* a = vp + (int64_t)s * (C * C * 32);
#pragma unroll
      for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) acc[i] = fma(__ldcs(a + (i * C + j) * 32), pw[j], acc[i]);
    }
    if (brow >= 0) {
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const int64_t row = (int64_t)C * brow + i;
        if (RESIDUAL) y[row] = fr[i] ? A.b[row] - acc[i] : 0.0;
        else {
          const double qi = fr[i] ? acc[i] : 0.0;
          y[row] = qi;
          partial += fr[i] ? vec[row] * acc[i] : 0.0;
          if constexpr (DOTS) {  // fixed rows carry r = dinv = 0
            const double ri = A.r[row], di = A.dinv[row];
            d_rho += di * ri * ri; d_s1 += di * ri * qi; d_s2 += di * qi * qi;
            d_rr += ri * ri;       d_t1 += ri * qi;      d_t2 += qi * qi;
          }
        }
      }
    }
  }
  if constexpr (DOTS) {
    dots[0] += d_rho; dots[1] += d_s1; dots[2] += d_s2; dots[3] += d_rr; dots[4] += d_t1; dots[5] += d_t2;
  }
  return partial;
}

#ifndef FES_PCG_FUSED_BLOCKS
#define FES_PCG_FUSED_BLOCKS 3
#endif
// FUSED (single GPU, FES_PCG_FUSED=1): one reduction and two grid barriers per iteration; its own
// instantiation, so the register allocation of each loop is independent of the other's.  Measured on the
// B200 (bench.py extras): config 2 179 us / iteration against 172 for the three-barrier loop, config 4
// SIMP 1.49 against 1.68 design iterations / s, config 5 tets 1237 against 1031 us -- the six extra dot
// products and the two extra vector reads inside the product phase cost the HBM-bound SpMV more than the
// barrier and the reduction they remove, so the textbook loop stays the default.
//
// Distributed FUSED variant (FES_PCG_DIST_FUSED=1; the default is the three-phase loop with overlap level 3, see
// fes_solver_solve): one rank reduction instead of two per iteration (a whole NVLink
// round trip less), and the halo of p is posted without a grid barrier at the top of the iteration and
// awaited only by the rows that read ghost entries, after the interior rows have been multiplied.
#ifndef FES_PCG_DIST_BLOCKS
#define FES_PCG_DIST_BLOCKS 1  // CTAs per SM of the overlapped distributed loop: 1 = 128 registers, no spills (2 spills ~300 B)
#endif
template <int C, bool DIST, bool FUSED>
__global__ void __launch_bounds__(kPcgBlock, (DIST && FUSED) ? FES_PCG_DIST_BLOCKS
                                             : (C == 3 || DIST) ? 2 : (FUSED ? FES_PCG_FUSED_BLOCKS : 3)) k_pcg(PcgArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[7 * 33];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;

  constexpr bool dist = DIST;  // the single-GPU instantiation carries none of the peer code
  unsigned long long halo_epoch = 0, red_step = 0;
  if (dist) {
    const volatile unsigned long long* ctr = reinterpret_cast<const volatile unsigned long long*>(A.comm + kCommCounters);
    halo_epoch = ctr[0];
    red_step = ctr[1];
  }
  const int64_t lo = A.row_begin, hi = A.row_end;

  // ---- r0 = b - A x0 on the owned free rows (x0 = 0 unless warm), p = 0
  if (A.warm) {
    const double* x_all = A.x;
    if (dist) {  // the product needs x on the ghost nodes: ship it through the shared vector
      for (int64_t i = lo + tid; i < hi; i += nth) A.p[i] = A.x[i];
      grid.sync();
      halo_exchange(A, grid, A.p, ++halo_epoch, tid, nth);
      x_all = A.p;
    }
    sparse_product<C, true, true>(A, tid, nth, x_all, A.r);
    grid.sync();
  }
  // Fixed rows carry dinv = 0 and r = p = q = 0 from here on, so the vector phases below need no
  // mask: every update leaves them at zero and x untouched.
  double v3[3] = {0.0, 0.0, 0.0};  // b.b, r.r, r.z
  for (int64_t i = lo + tid; i < hi; i += nth) {
    const bool fr = is_free(A.mask, i);
    const double bi = fr ? A.b[i] : 0.0;
    double ri = bi;
    if (A.warm) ri = A.r[i];  // the residual product already wrote 0 on fixed rows
    else { A.r[i] = bi; if (fr) A.x[i] = 0.0; }
    A.p[i] = FUSED ? A.dinv[i] * ri : 0.0;  // fused variant: p0 = z0 before the first product
    A.q[i] = 0.0;
    v3[0] += bi * bi;
    v3[1] += ri * ri;
    v3[2] += ri * ri * A.dinv[i];
  }
  grid_sum<3>(grid, v3, A.partials, 0, sm);
  if (dist) rank_sum<3>(A, v3, ++red_step, sm);
  if constexpr (dist && FUSED) {
    // p0 of the dofs the peers need.  Only now: rank_sum was a barrier across the ranks, so no peer is still
    // multiplying the warm-start x out of the ghost entries these stores overwrite.  The grid barrier orders
    // the stores before the first halo_publish.
    for (int k = (int)tid; k < A.n_ship; k += (int)nth) ship(A, k, A.p[A.ship_dof[k]]);
    __threadfence_system();
    grid.sync();
  }
  const double bb = v3[0];
  double rr = v3[1], rho = v3[2], rho_prev = 1.0;
  long long it = 0;
  int status = 0;
  // phase timers (FES_PCG_TIMERS): thread 0 adds the SM cycles since the previous tick to phase k, in global memory
  long long t_last = clock64();
  if (A.timers && tid == 0)
    for (int k = 0; k < 6; ++k) A.state->cycles[k] = 0;
  auto tick = [&](int k) {
    if (A.timers) {
      const long long t = clock64();
      if (tid == 0) A.state->cycles[k] += t - t_last;
      t_last = t;
    }
  };
  if (bb == 0.0) {  // scipy returns x = 0 for a zero right-hand side
    for (int64_t i = lo + tid; i < hi; i += nth)
      if (is_free(A.mask, i)) A.x[i] = 0.0;
    rr = 0.0;
  } else if constexpr (FUSED) {
    // One reduction and two grid barriers per iteration.  The product phase also accumulates, over
    // the rows it just produced,  rho = sum dinv r^2, s1 = sum dinv r q, s2 = sum dinv q^2 and the same
    // three sums without dinv; then  alpha = rho / p.q  and the NEXT residual's norms follow without
    // touching memory:  (r - alpha q)^T D (r - alpha q) = rho - 2 alpha s1 + alpha^2 s2, so beta is
    // known before the vector phase, which does x += alpha p, r -= alpha q, p = D r + beta p in one
    // pass.  rho is re-summed from the stored residual every iteration (no drift); only beta and the
    // stopping test use the one-step prediction, whose cancellation error is eps * rho_k / rho_{k+1}.
    for (;;) {
      if (sqrt(rr) < A.rtol * sqrt(bb)) break;
      if (it >= A.maxiter) { status = 1; break; }
      tick(5);
      double v7[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (dist) {
        // the boundary values of p were shipped and fenced before the last grid barrier: publish, do not wait
        ++halo_epoch;
        halo_publish(A, halo_epoch);
        v7[0] = sparse_product<C, false, false, true>(A, tid, nth, A.p, A.q, v7 + 1, 0, A.n_slots_interior);
        if (A.n_slots_interior + (int64_t)blockIdx.x * blockDim.x < A.n_slots) {  // CTA-uniform: this CTA multiplies interface rows
          tick(2);
          halo_wait(A, halo_epoch);
          tick(1);
          v7[0] += sparse_product<C, false, false, true, true>(A, tid, nth, A.p, A.q, v7 + 1, A.n_slots_interior, A.n_slots);
        }
      } else {
        v7[0] = sparse_product<C, false, false, true>(A, tid, nth, A.p, A.q, v7 + 1);
      }
      tick(2);
      grid_sum<7>(grid, v7, A.partials, 0, sm);
      if (dist) rank_sum<7>(A, v7, ++red_step, sm);
      tick(3);
      const double pq = v7[0], rho_k = v7[1];
      if (!(pq > 0.0)) { status = 2; break; }
      const double alpha = rho_k / pq;
      const double rho_next = fmax(fma(alpha, fma(alpha, v7[3], -2.0 * v7[2]), rho_k), 0.0);
      const double rr_next = fmax(fma(alpha, fma(alpha, v7[6], -2.0 * v7[5]), v7[4]), 0.0);
      const double beta = rho_next / rho_k;
      if (dist) {  // dofs some peer needs: updated once, by the thread that also ships the new p
        bool sent = false;
        for (int k = (int)tid; k < A.n_ship; k += (int)nth) {
          const int32_t i = A.ship_dof[k];
          const double pi = A.p[i];
          if (pi != 0.0) A.x[i] = fma(alpha, pi, A.x[i]);
          const double ri = A.r[i] - alpha * A.q[i];
          A.r[i] = ri;
          const double pn = fma(beta, pi, A.dinv[i] * ri);
          A.p[i] = pn;
          ship(A, k, pn);
          sent = true;
        }
        if (sent) __threadfence_system();  // peer stores performed before this thread reaches the grid barrier
      }
      for (int64_t i = lo + tid; i < hi; i += nth) {
        if (dist && A.shipped[i]) continue;
        const double pi = A.p[i];
        if (pi != 0.0) A.x[i] = fma(alpha, pi, A.x[i]);  // fixed rows (p = 0) keep their prescribed value
        const double ri = A.r[i] - alpha * A.q[i];
        A.r[i] = ri;
        A.p[i] = fma(beta, pi, A.dinv[i] * ri);
      }
      tick(4);
      grid.sync();
      tick(0);
      rr = rr_next;
      ++it;
    }
  } else {
    for (;;) {
      if (sqrt(rr) < A.rtol * sqrt(bb)) break;
      if (it >= A.maxiter) { status = 1; break; }
      const double beta = (it == 0) ? 0.0 : rho / rho_prev;
      tick(5);
      double v1[1];
      if (dist && A.overlap) {
        // The textbook recurrences with the exchange hidden: the thread that computes the new p of a dof some peer
        // needs ships it (NVLink store) inside the p update; after the grid barrier one thread publishes the epoch,
        // the interior rows are multiplied, and only the rows that read ghost entries wait for the neighbours.
        bool sent = false;
        for (int k = (int)tid; k < A.n_ship; k += (int)nth) {
          const int32_t i = A.ship_dof[k];
          const double pn = fma(beta, A.p[i], A.dinv[i] * A.r[i]);
          A.p[i] = pn;
          ship(A, k, pn);
          sent = true;
        }
        for (int64_t i = lo + tid; i < hi; i += nth)
          if (!A.shipped[i]) A.p[i] = fma(beta, A.p[i], A.dinv[i] * A.r[i]);
        if (sent) __threadfence_system();  // after the plain update: the NVLink stores have been in flight meanwhile
        grid.sync();
        tick(0);
        ++halo_epoch;
        halo_publish(A, halo_epoch);
        if (A.overlap >= 2) { halo_wait(A, halo_epoch); tick(1); }   // wait first
        if (A.overlap == 3) {
          // one pass, plain loads: L1 was invalidated by the grid barrier above and no ghost line has been read since
          v1[0] = sparse_product<C, false, false>(A, tid, nth, A.p, A.q);
        } else {
          v1[0] = sparse_product<C, false, false>(A, tid, nth, A.p, A.q, nullptr, 0, A.n_slots_interior);
          if (A.n_slots_interior + (int64_t)blockIdx.x * blockDim.x < A.n_slots) {  // CTA-uniform
            tick(2);
            if (A.overlap == 1) halo_wait(A, halo_epoch);
            tick(1);
            v1[0] += sparse_product<C, false, false, false, true>(A, tid, nth, A.p, A.q, nullptr, A.n_slots_interior, A.n_slots);
          }
        }
      } else {
        for (int64_t i = lo + tid; i < hi; i += nth) A.p[i] = fma(beta, A.p[i], A.dinv[i] * A.r[i]);
        grid.sync();
        tick(0);
        if (dist) halo_exchange(A, grid, A.p, ++halo_epoch, tid, nth);
        tick(1);
        v1[0] = sparse_product<C, false, false>(A, tid, nth, A.p, A.q);
      }
      tick(2);
      grid_sum<1>(grid, v1, A.partials, 3, sm);
      if (dist) rank_sum<1>(A, v1, ++red_step, sm);
      tick(3);
      const double pq = v1[0];
      if (!(pq > 0.0)) { status = 2; break; }
      const double alpha = rho / pq;

      double v2[2] = {0.0, 0.0};
      for (int64_t i = lo + tid; i < hi; i += nth) {
        const double pi = A.p[i];
        if (pi != 0.0) A.x[i] = fma(alpha, pi, A.x[i]);  // fixed rows (p = 0) keep their prescribed value
        const double ri = A.r[i] - alpha * A.q[i];
        A.r[i] = ri;
        v2[0] += ri * ri;
        v2[1] += ri * ri * A.dinv[i];
      }
      tick(4);
      grid_sum<2>(grid, v2, A.partials, 1, sm);
      if (dist) rank_sum<2>(A, v2, ++red_step, sm);
      tick(5);
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
    if (dist) {
      volatile unsigned long long* ctr = reinterpret_cast<volatile unsigned long long*>(A.comm + kCommCounters);
      ctr[0] = halo_epoch;
      ctr[1] = red_step;
    }
  }
}

// one thread per SELL row slot: column index and the c x c values of block (brow, s), zero padding
template <int C>
__global__ void k_sell_fill(int64_t n_slots, const int32_t* __restrict__ sell_row,
                            const int32_t* __restrict__ slice_ptr, const int32_t* __restrict__ row_ptr,
                            const int32_t* __restrict__ row_col, const double* __restrict__ data,
                            int32_t* __restrict__ cols, double* __restrict__ vals, int fill_cols) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const int lane = (int)(slot & 31);
  const int64_t slice = slot >> 5;
  const int32_t base = slice_ptr[slice], width = (slice_ptr[slice + 1] - base) >> 5;
  const int32_t brow = sell_row[slot];
  int32_t p0 = 0, deg = 0;
  if (brow >= 0) { p0 = row_ptr[brow]; deg = row_ptr[brow + 1] - p0; }
  for (int32_t s = 0; s < width; ++s) {
    const int64_t e = (int64_t)base + (int64_t)s * 32 + lane;
    const bool live = s < deg;
    if (fill_cols) cols[e] = live ? row_col[p0 + s] : (brow >= 0 ? brow : 0);
    double* v = vals + ((int64_t)base + (int64_t)s * 32) * (C * C) + lane;
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j)
        v[(i * C + j) * 32] = live ? data[(int64_t)C * C * p0 + ((int64_t)i * deg + s) * C + j] : 0.0;
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

template <int C, bool DIST, bool FUSED>
int launch_pcg_variant(fes_solver* s, PcgArgs& args, cudaStream_t st) {
  if (s->grid == 0 || s->grid_fused != (FUSED ? 1 : 0)) {
    s->grid_fused = FUSED ? 1 : 0;
    int dev = 0, sms = 0, per_sm = 0;
    FES_CUDA(cudaGetDevice(&dev));
    FES_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg<C, DIST, FUSED>, kPcgBlock, 0));
    if (per_sm < 1) return set_error(FES_ERR_CUDA, "PCG kernel does not fit on an SM");
    s->grid = sms * per_sm;
    if (s->grid > kMaxGrid) s->grid = kMaxGrid;
    s->block = kPcgBlock;
  }
  void* params[] = {&args};
  FES_CUDA(cudaLaunchCooperativeKernel((void*)k_pcg<C, DIST, FUSED>, dim3(s->grid), dim3(s->block), params, 0, st));
  count_launch();
  return FES_OK;
}

template <int C, bool DIST>
int launch_pcg(fes_solver* s, PcgArgs& args, cudaStream_t st) {
  if (args.fused) return launch_pcg_variant<C, DIST, true>(s, args, st);
  return launch_pcg_variant<C, DIST, false>(s, args, st);
}

// 1 = the block row reads a column outside the owned block range [b0, b1) (a ghost node)
__global__ void k_interface_rows(int64_t b0, int64_t b1, const int32_t* __restrict__ row_ptr,
                                 const int32_t* __restrict__ row_col, uint8_t* __restrict__ flag) {
  const int64_t r = b0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b1) return;
  uint8_t f = 0;
  for (int32_t k = row_ptr[r]; k < row_ptr[r + 1]; ++k) {
    const int32_t c = row_col[k];
    f |= (c < b0 || c >= b1) ? 1 : 0;
  }
  flag[r - b0] = f;
}

// SELL-32-sigma structure over block rows: the row order and slice offsets are computed on the
// host from the row lengths (O(rows), once per operator); columns and values are filled on the
// device.
int build_sell(fes_solver* s, cudaStream_t st) {
  const int64_t nb_all = s->n_brow;
  std::vector<int32_t> ptr(nb_all + 1);
  FES_CUDA(cudaMemcpyAsync(ptr.data(), s->row_ptr, (nb_all + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  // only the owned block rows are multiplied (all of them on one GPU)
  const int64_t b0 = s->row_begin / s->c, b1 = s->row_end / s->c, nb = b1 - b0;
  // Distributed: rows that read no ghost column come first (their slices are multiplied while the halo
  // is in flight), the interface rows follow from the next slice boundary.
  std::vector<int32_t> seq;   // block rows in multiplication order, -1 = padding
  seq.reserve(nb + 32);
  int64_t n_interior_slots = 0;
  if (s->comm != nullptr && s->comm->world > 1 && nb > 0) {
    DevBuf<uint8_t> d_flag;
    FES_CUDA(d_flag.alloc(nb));
    k_interface_rows<<<grid_for(nb, 256), 256, 0, st>>>(b0, b1, s->row_ptr, s->row_col, d_flag.p);
    FES_LAUNCH_CHECK();
    std::vector<uint8_t> flag(nb);
    FES_CUDA(cudaMemcpyAsync(flag.data(), d_flag.p, nb, cudaMemcpyDeviceToHost, st));
    FES_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < nb; ++i) if (!flag[i]) seq.push_back((int32_t)(b0 + i));
    while (seq.size() % 32) seq.push_back(-1);
    n_interior_slots = (int64_t)seq.size();
    for (int64_t i = 0; i < nb; ++i) if (flag[i]) seq.push_back((int32_t)(b0 + i));
  } else {
    for (int64_t i = 0; i < nb; ++i) seq.push_back((int32_t)(b0 + i));
    n_interior_slots = 0;  // unused on one GPU
  }
  const int64_t n_seq = (int64_t)seq.size();
  const int64_t n_slices = (n_seq + 31) / 32;
  s->n_slots_interior = n_interior_slots;
  auto len_of = [&](int32_t r) { return r < 0 ? -1 : ptr[r + 1] - ptr[r]; };
  std::vector<int32_t> row(n_slices * 32, -1), slice_ptr(n_slices + 1, 0);
  std::vector<int32_t> order;
  // sigma windows never straddle the interior / interface boundary (a multiple of 32, not of sigma)
  for (int part = 0; part < 2; ++part) {
    const int64_t p0 = part == 0 ? 0 : n_interior_slots, p1 = part == 0 ? (n_interior_slots ? n_interior_slots : n_seq) : n_seq;
    if (part == 1 && n_interior_slots == 0) break;
    for (int64_t w0 = p0; w0 < p1; w0 += kSigma) {
      const int64_t w1 = std::min<int64_t>(p1, w0 + kSigma);
      order.assign(seq.begin() + w0, seq.begin() + w1);
      std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return len_of(a) > len_of(b); });
      for (int64_t i = w0; i < w1; ++i) row[i] = order[i - w0];
    }
  }
  int64_t total = 0;
  for (int64_t sl = 0; sl < n_slices; ++sl) {
    int32_t width = 0;
    for (int l = 0; l < 32; ++l) {
      const int32_t r = row[sl * 32 + l];
      if (r >= 0) width = std::max(width, ptr[r + 1] - ptr[r]);
    }
    slice_ptr[sl] = (int32_t)total;
    total += (int64_t)width * 32;
    if (total >= ((int64_t)1 << 31)) return set_error(FES_ERR_VALUE, "operator too large for int32 SELL offsets");
  }
  slice_ptr[n_slices] = (int32_t)total;
  s->n_slices = n_slices;
  s->sell_entries = total;
  FES_CUDA(s->sell_row.alloc(row.size()));
  FES_CUDA(s->sell_slice_ptr.alloc(slice_ptr.size()));
  FES_CUDA(s->sell_cols.alloc(total));
  FES_CUDA(s->sell_vals.alloc((size_t)total * s->c * s->c));
  FES_CUDA(cudaMemcpyAsync(s->sell_row.p, row.data(), row.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FES_CUDA(cudaMemcpyAsync(s->sell_slice_ptr.p, slice_ptr.data(), slice_ptr.size() * sizeof(int32_t),
                           cudaMemcpyHostToDevice, st));
  FES_CUDA(cudaStreamSynchronize(st));
  return FES_OK;
}

int fill_sell(fes_solver* s, int fill_cols, cudaStream_t st) {
  const int64_t n_slots = s->n_slices * 32;
  if (n_slots == 0) return FES_OK;
  const int g = grid_for(n_slots, 256);
#define FES_FILL(C_)                                                                                         \
  k_sell_fill<C_><<<g, 256, 0, st>>>(n_slots, s->sell_row.p, s->sell_slice_ptr.p, s->row_ptr, s->row_col, s->data, \
                                     s->sell_cols.p, s->sell_vals.p, fill_cols)
  if (s->c == 1) FES_FILL(1);
  else if (s->c == 2) FES_FILL(2);
  else FES_FILL(3);
#undef FES_FILL
  FES_LAUNCH_CHECK();
  return FES_OK;
}

int init_solver(fes_solver* s, cudaStream_t st) {
  s->lanes = pick_lanes(s->n, s->nnz);
  FES_CUDA(s->dinv.alloc(s->n));
  FES_CUDA(s->r.alloc(s->n));
  FES_CUDA(s->p.alloc(s->n));
  FES_CUDA(s->q.alloc(s->n));
  FES_CUDA(s->partials.alloc(8 * kMaxGrid));
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
  s->c = 1; s->n_brow = s->n; s->row_ptr = s->indptr; s->row_col = s->indices;
  s->row_begin = 0; s->row_end = s->n;
  return FES_OK;
}

// (re)build the SELL copy lazily: structure on first use, values whenever the operator changed
int ensure_sell(fes_solver* s, cudaStream_t st) {
  if (s->n == 0) return FES_OK;
  if (s->sell_row.p == nullptr) {
    FES_TRY(build_sell(s, st));
    FES_TRY(fill_sell(s, 1, st));
  }
  return FES_OK;
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
  if (s->sell_row.p != nullptr) FES_TRY(fill_sell(s, 0, st));  // same structure, new values
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
  int rc = init_solver(s, (cudaStream_t)stream);
  if (rc == FES_OK) rc = fes_solver_refresh(s, stream);
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
    std::vector<int32_t> tmp((size_t)std::max<int64_t>(nnz, n + 1));
    for (int64_t i = 0; i <= n; ++i) tmp[i] = (int32_t)h_indptr[i];
    FES_CUDA(cudaMemcpy(s->own_indptr.p, tmp.data(), (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    for (int64_t i = 0; i < nnz; ++i) tmp[i] = (int32_t)h_indices[i];
    if (nnz) FES_CUDA(cudaMemcpy(s->own_indices.p, tmp.data(), nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  if (nnz) FES_CUDA(cudaMemcpy(s->own_data.p, h_data, nnz * sizeof(double), cudaMemcpyHostToDevice));
  s->indptr = s->own_indptr.p; s->indices = s->own_indices.p; s->data = s->own_data.p; s->mask = nullptr;
  FES_TRY(init_solver(s, nullptr));
  FES_TRY(fes_solver_refresh(s, nullptr));
  guard.p = nullptr;
  *out = s;
  return FES_OK;
}

extern "C" void fes_solver_destroy(fes_solver* s) { delete s; }

// Bytes of the operator copy one product streams (block-SELL values + one column index per block + the
// row-slot table + slice offsets); 0 before the first solve built it.
extern "C" int64_t fes_solver_operator_bytes(const fes_solver* s) {
  if (!s || s->sell_row.p == nullptr) return 0;
  return s->sell_entries * ((int64_t)8 * s->c * s->c + 4) + s->n_slices * 32 * 4 + (s->n_slices + 1) * 4;
}

extern "C" int fes_solver_use_plan(fes_solver* s, const fes_plan* plan) {
  if (!s || !plan) return set_error(FES_ERR_VALUE, "fes_solver_use_plan: NULL argument");
  if (plan->n_dofs != s->n || s->indptr != plan->indptr.p)
    return set_error(FES_ERR_VALUE, "fes_solver_use_plan: the operator was not assembled from this plan");
  if ((plan->c == 2 || plan->c == 3) && s->sell_row.p == nullptr) {
    s->c = plan->c;
    s->n_brow = plan->n_nodes;
    s->row_ptr = plan->node_ptr.p;
    s->row_col = plan->node_adj.p;
    s->grid = 0;  // occupancy of the block-row kernel differs
  }
  return FES_OK;
}

extern "C" int fes_comm_create(int rank, int world, int64_t max_local_dofs, fes_comm** out) {
  if (!out) return set_error(FES_ERR_VALUE, "fes_comm_create: out is NULL");
  *out = nullptr;
  if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || max_local_dofs < 0)
    return set_error(FES_ERR_VALUE, "fes_comm_create: bad rank/world (%d/%d; at most %d ranks)", rank, world, kCommMaxWorld);
  fes_comm* c = new fes_comm();
  c->rank = rank; c->world = world; c->max_dofs = max_local_dofs;
  c->bytes = kCommVec + (size_t)max_local_dofs * sizeof(double);
  if (cudaMalloc((void**)&c->base, c->bytes) != cudaSuccess || cudaMemset(c->base, 0, c->bytes) != cudaSuccess) {
    delete c;
    return set_error(FES_ERR_CUDA, "fes_comm_create: cannot allocate %zu bytes", c->bytes);
  }
  c->peer[rank] = c->base;
  *out = c;
  return FES_OK;
}

extern "C" int fes_comm_handle(const fes_comm* c, void* out64) {
  if (!c || !out64) return set_error(FES_ERR_VALUE, "fes_comm_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  FES_CUDA(cudaIpcGetMemHandle(&h, c->base));
  memcpy(out64, &h, 64);
  return FES_OK;
}

extern "C" int fes_comm_connect(fes_comm* c, const void* handles) {
  if (!c || !handles) return set_error(FES_ERR_VALUE, "fes_comm_connect: NULL argument");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    FES_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = (unsigned char*)p;
  }
  FES_CUDA(c->d_peer.alloc(kCommMaxWorld));
  FES_CUDA(cudaMemcpy(c->d_peer.p, c->peer, sizeof(c->peer), cudaMemcpyHostToDevice));
  c->connected = true;
  return FES_OK;
}

extern "C" void fes_comm_destroy(fes_comm* c) {
  if (!c) return;
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->base) cudaFree(c->base);
  delete c;
}

extern "C" int fes_solver_set_dist(fes_solver* s, fes_comm* comm, int64_t row_begin, int64_t row_end,
                                   const int32_t* h_send_ptr, const int32_t* h_send_src, const int32_t* h_send_dst,
                                   int64_t n_free_global) {
  if (!s || !comm || !h_send_ptr) return set_error(FES_ERR_VALUE, "fes_solver_set_dist: NULL argument");
  if (n_free_global < 0) return set_error(FES_ERR_VALUE, "fes_solver_set_dist: n_free_global < 0");
  if (row_begin < 0 || row_end > s->n || row_begin > row_end || row_begin % s->c || row_end % s->c)
    return set_error(FES_ERR_VALUE, "fes_solver_set_dist: bad owned row range [%lld, %lld)", (long long)row_begin,
                     (long long)row_end);
  if (s->n > comm->max_dofs) return set_error(FES_ERR_VALUE, "fes_solver_set_dist: communicator vector too small");
  if (s->sell_row.p != nullptr) return set_error(FES_ERR_VALUE, "fes_solver_set_dist: call before the first solve");
  s->comm = comm;
  s->row_begin = row_begin; s->row_end = row_end;
  s->n_free_global = n_free_global;
  const int world = comm->world;
  {  // every sent dof once, with all its (rank, ghost index) targets
    const int32_t total = h_send_ptr[world];
    std::vector<int32_t> order(total), rank_of(total);
    for (int r = 0; r < world; ++r)
      for (int32_t k = h_send_ptr[r]; k < h_send_ptr[r + 1]; ++k) rank_of[k] = r;
    for (int32_t k = 0; k < total; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return h_send_src[a] < h_send_src[b]; });
    std::vector<int32_t> dof, ptr, rk(total), dst(total);
    std::vector<uint8_t> flag(s->n > 0 ? s->n : 1, 0);
    for (int32_t j = 0; j < total; ++j) {
      const int32_t k = order[j], i = h_send_src[k];
      if (i < row_begin || i >= row_end) return set_error(FES_ERR_VALUE, "fes_solver_set_dist: send list names a dof this rank does not own");
      if (dof.empty() || dof.back() != i) { dof.push_back(i); ptr.push_back(j); flag[i] = 1; }
      rk[j] = rank_of[k];
      dst[j] = h_send_dst[k];
    }
    ptr.push_back(total);
    s->n_ship = (int)dof.size();
    FES_CUDA(s->ship_dof.alloc(dof.size() ? dof.size() : 1));
    FES_CUDA(s->ship_ptr.alloc(ptr.size()));
    FES_CUDA(s->ship_rank.alloc(total > 0 ? total : 1));
    FES_CUDA(s->ship_dst.alloc(total > 0 ? total : 1));
    FES_CUDA(s->shipped.alloc(flag.size()));
    if (!dof.empty()) FES_CUDA(cudaMemcpy(s->ship_dof.p, dof.data(), dof.size() * 4, cudaMemcpyHostToDevice));
    FES_CUDA(cudaMemcpy(s->ship_ptr.p, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
    if (total > 0) {
      FES_CUDA(cudaMemcpy(s->ship_rank.p, rk.data(), (size_t)total * 4, cudaMemcpyHostToDevice));
      FES_CUDA(cudaMemcpy(s->ship_dst.p, dst.data(), (size_t)total * 4, cudaMemcpyHostToDevice));
    }
    FES_CUDA(cudaMemcpy(s->shipped.p, flag.data(), flag.size(), cudaMemcpyHostToDevice));
  }
  const int32_t total = h_send_ptr[world];
  FES_CUDA(s->send_ptr.alloc(world + 1));
  FES_CUDA(s->send_src.alloc(total > 0 ? total : 1));
  FES_CUDA(s->send_dst.alloc(total > 0 ? total : 1));
  FES_CUDA(cudaMemcpy(s->send_ptr.p, h_send_ptr, (world + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (total > 0) {
    FES_CUDA(cudaMemcpy(s->send_src.p, h_send_src, total * sizeof(int32_t), cudaMemcpyHostToDevice));
    FES_CUDA(cudaMemcpy(s->send_dst.p, h_send_dst, total * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  s->send_mask = 0;
  for (int r = 0; r < world; ++r)
    if (h_send_ptr[r + 1] > h_send_ptr[r]) s->send_mask |= 1u << r;
  // the halo is symmetric for a symmetric pattern: whoever needs my nodes owns nodes I need
  s->recv_mask = s->send_mask;
  return FES_OK;
}

extern "C" int fes_solver_solve(fes_solver* s, const double* d_b, double* d_x, double rtol, int64_t maxiter,
                                int warm_start, int64_t* iters_out, double* relres_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (!s) return set_error(FES_ERR_VALUE, "fes_solver_solve: NULL solver");
  if (iters_out) *iters_out = 0;
  if (relres_out) *relres_out = 0.0;
  const bool dist = s->comm != nullptr && s->comm->world > 1;
  // Every exit condition of a distributed solve is global: a rank whose local dofs are all fixed still
  // joins the collective kernel, and the default iteration cap comes from the global free-dof count
  // (scipy: maxiter = 10 n), so all ranks stop at the same iteration.
  const int64_t n_free = dist ? s->n_free_global : s->n_free;
  if (s->n == 0 || n_free == 0) return FES_OK;
  if (maxiter <= 0) maxiter = 10 * n_free;
  if (warm_start && s->c == 2 && !dist && (reinterpret_cast<uintptr_t>(d_x) & 15))
    return set_error(FES_ERR_VALUE, "fes_solver_solve: a warm-start vector of a 2-component operator must be 16-byte aligned");
  FES_TRY(ensure_sell(s, st));
  if (dist && !s->comm->connected) return set_error(FES_ERR_VALUE, "fes_solver_solve: communicator is not connected");
  double* pvec = dist ? reinterpret_cast<double*>(s->comm->base + kCommVec) : s->p.p;
  PcgArgs a{s->n, s->mask, s->dinv.p, d_b, d_x, s->r.p, pvec, s->q.p, s->partials.p, s->state.p, rtol,
            (long long)maxiter, warm_start ? 1 : 0, s->n_slices * 32, s->sell_row.p, s->sell_slice_ptr.p,
            s->sell_cols.p, s->sell_vals.p, s->row_begin, s->row_end, dist ? s->comm->world : 1,
            dist ? s->comm->rank : 0, dist ? s->comm->base : nullptr, dist ? s->comm->d_peer.p : nullptr,
            s->send_ptr.p, s->send_src.p, s->send_dst.p, s->recv_mask, s->send_mask,
            // one GPU: the textbook three-barrier loop unless FES_PCG_FUSED is set (A-B knob; measured slower, see
            // k_pcg).  Distributed: the textbook loop with the exchange overlapped (default); FES_PCG_DIST_FUSED
            // selects the single-reduction overlapped loop, FES_PCG_DIST_CLASSIC the loop with the halo barrier.
            dist ? (getenv("FES_PCG_DIST_FUSED") ? 1 : 0) : (getenv("FES_PCG_FUSED") ? 1 : 0),
            dist ? (getenv("FES_PCG_DIST_CLASSIC") ? 0 : (getenv("FES_PCG_DIST_MODE") ? atoi(getenv("FES_PCG_DIST_MODE")) : 3)) : 0,
            getenv("FES_PCG_TIMERS") ? 1 : 0, s->n_slots_interior, s->n_ship, s->ship_dof.p, s->ship_ptr.p,
            s->ship_rank.p, s->ship_dst.p, s->shipped.p};
  if (dist) {
    if (s->c == 2) FES_TRY((launch_pcg<2, true>(s, a, st)));
    else if (s->c == 3) FES_TRY((launch_pcg<3, true>(s, a, st)));
    else FES_TRY((launch_pcg<1, true>(s, a, st)));
  } else {
    if (s->c == 2) FES_TRY((launch_pcg<2, false>(s, a, st)));
    else if (s->c == 3) FES_TRY((launch_pcg<3, false>(s, a, st)));
    else FES_TRY((launch_pcg<1, false>(s, a, st)));
  }
  PcgState h;
  FES_CUDA(cudaMemcpyAsync(&h, s->state.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  FES_CUDA(cudaStreamSynchronize(st));
  if (iters_out) *iters_out = h.iters;
  if (relres_out) *relres_out = h.relres;
  if (getenv("FES_PCG_TIMERS") && h.iters > 0) {
    const char* names[6] = {"p-update", "halo", "product", "product-reduce", "xr-update", "update-reduce"};
    fprintf(stderr, "[fes pcg] rank %d, %lld iterations, SM cycles per iteration:", dist ? s->comm->rank : 0, h.iters);
    for (int k = 0; k < 6; ++k) fprintf(stderr, " %s=%lld", names[k], h.cycles[k] / h.iters);
    fprintf(stderr, "\n");
  }
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
