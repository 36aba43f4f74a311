// The CSR pattern + scatter plan handle shared between plan.cu (builder) and assemble.cu.
#pragma once
#include "common.cuh"

struct fes_plan {
  int64_t n_el = 0, n_nodes = 0, n_dofs = 0, n_pairs = 0, nnz = 0, n_contrib = 0;
  int N = 0, c = 0;
  // node-level (block) CSR: row v holds its neighbour nodes in ascending order
  fes::DevBuf<int32_t> node_ptr;    // n_nodes + 1
  fes::DevBuf<int32_t> node_adj;    // n_pairs
  fes::DevBuf<int32_t> pair_row;    // n_pairs : row node of each pair
  // contributions: for pair p, contrib[pair_start[p] .. pair_start[p+1]) = e*N*N + a*N + b in
  // ascending element id -- the order np.bincount sums them in
  fes::DevBuf<uint32_t> pair_start;  // n_pairs + 1
  fes::DevBuf<uint32_t> contrib;     // n_el * N * N
  // scalar CSR, bit-identical to the reference's _ScatterPlan.indptr / .indices
  fes::DevBuf<int32_t> indptr;   // n_dofs + 1
  fes::DevBuf<int32_t> indices;  // nnz
};
