/* fes_b200.h -- C ABI of libfes_b200.so: the B200 (sm_100a) hot path of
 * janetyq/Finite-Element-Solver (batched element matrices -> CSR assembly -> Jacobi-PCG ->
 * SIMP inner loop), fp64, behind the reference's own Python seams.
 *
 * The reference is pure Python and has no FFI; each entry point below names the reference
 * function(s) it replaces (file:line into the upstream tree).  INTEGRATION.md shows the
 * ctypes stubs a maintainer would add to fem/space.py, fem/backends.py and fem/topology.py.
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer, `h_` = host pointer.
 *   - every function returns an int status: FES_OK or a negative FES_ERR_*; the message is
 *     available from fes_last_error() (thread-local).  The Python layer maps
 *     FES_ERR_VALUE -> ValueError, FES_ERR_UNSUPPORTED -> NotImplementedError,
 *     FES_ERR_CG -> RuntimeError("CG failed ..."), everything else -> RuntimeError.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous on that stream unless stated; *_host variants synchronise before returning.
 *   - connectivity on the device is int32 (the reference holds int64; values < 2^31);
 *     DOF of (node v, component d) = n_comp*v + d (ref fem/space.py:55-66).
 *   - element matrices are (n_el, k, k) row-major, k = N*n_comp, node-major / component-minor
 *     (ref fem/space.py:766-772, fem/forms.py:165-171).
 */
#ifndef FES_B200_H
#define FES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FES_OK 0
#define FES_ERR_VALUE (-1)       /* bad argument / size mismatch  -> ValueError            */
#define FES_ERR_CUDA (-2)        /* CUDA runtime failure          -> RuntimeError          */
#define FES_ERR_UNSUPPORTED (-3) /* element/form combination      -> NotImplementedError   */
#define FES_ERR_CG (-4)          /* PCG did not converge/breakdown-> RuntimeError("CG failed") */
#define FES_ERR_NCCL (-5)

/* element kinds (ref fem/elements.py:300-413) */
#define FES_P1_LINE 0
#define FES_P1_TRI 1
#define FES_P1_TET 2
#define FES_P2_LINE 3
#define FES_P2_TRI 4

/* bilinear forms (ref fem/forms.py:190-258,318-333,444-480) */
#define FES_FORM_LAPLACIAN 0 /* LaplacianForm: sum_q w_q grad phi_i . grad phi_j                 */
#define FES_FORM_MASS 1      /* MassForm / MaskedMassForm: vol * kron(M_ref, I_c) [* mask]        */
#define FES_FORM_ELASTIC 2   /* LinearElasticForm: sum_q w_q B^T D B, plane strain in 2D           */

typedef struct fes_plan fes_plan;     /* CSR pattern + scatter plan of one (connectivity, n_comp)  */
typedef struct fes_solver fes_solver; /* Jacobi-PCG bound to one operator                          */
typedef struct fes_filter fes_filter; /* SIMP cone filter (row-normalised CSR over elements)       */
typedef struct fes_comm fes_comm;     /* peer-memory communicator of the distributed PCG           */

/* Coefficients of a form.  ScaledForm(factor, form) sets `factor`; per-element modulus
 * (LinearElasticMaterial(E=array)) sets d_E; MaskedMassForm's 0/1 mask and SIMP's rho^p go in
 * d_scale (one fp64 per element, multiplies the whole element matrix). */
typedef struct fes_coeffs {
  double E;              /* Young's modulus when d_E == NULL                  */
  double nu;             /* Poisson ratio                                      */
  double factor;         /* global scale (1.0 = none)                          */
  const double* d_E;     /* optional (n_el) per-element modulus                */
  const double* d_scale; /* optional (n_el) per-element scale                  */
} fes_coeffs;

/* ---- library / errors ------------------------------------------------------------------ */
int fes_version(void);
const char* fes_last_error(void);
int fes_set_device(int device); /* bind the calling thread to a CUDA device (one process per GPU) */
int fes_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t fes_launch_count(void);

/* ---- sparsity pattern + scatter plan ------------------------------------------------------
 * Replaces _ScatterPlan.build via FunctionSpace._scatter_plan (ref fem/space.py:107-131,
 * 757-772): built once per (element set, n_comp), reused by every assembly.  The pattern is
 * every (row, col) any element touches, explicit zeros kept, columns ascending -- indptr and
 * indices are bit-identical to the reference's.  Built on the device from a stable LSD radix
 * sort of the N^2*n_el node-pair keys, head-flag scan and block expansion by n_comp. */
int fes_plan_create(const int32_t* d_elem_nodes, int64_t n_el, int nodes_per_elem, int64_t n_nodes,
                    int n_comp, void* stream, fes_plan** out);
void fes_plan_destroy(fes_plan* plan);
int64_t fes_plan_n_dofs(const fes_plan* plan);
int64_t fes_plan_nnz(const fes_plan* plan);        /* stored scalar entries                        */
int64_t fes_plan_n_pairs(const fes_plan* plan);    /* node-level (block) entries                   */
int64_t fes_plan_bytes(const fes_plan* plan);      /* device bytes held by the plan                */
int64_t fes_plan_read_bytes(const fes_plan* plan); /* plan bytes one assembly reads (overhead)     */
int64_t fes_plan_n_tiles(const fes_plan* plan);    /* tiles of the fused kernel; 0 = untiled fallback */
/* Stencil runs (structured node ranges assembled from a per-template program instead of per-tile index
 * streams; P1 triangles / tets, Laplacian and elasticity): tiles and nodes they cover, 0 = none detected. */
int64_t fes_plan_n_run_tiles(const fes_plan* plan);
int64_t fes_plan_run_nodes(const fes_plan* plan);
const int32_t* fes_plan_indptr(const fes_plan* plan);  /* device, n_dofs+1                         */
const int32_t* fes_plan_indices(const fes_plan* plan); /* device, nnz                              */
/* host export as int64, the dtypes scipy's csr_array holds in the reference */
int fes_plan_export_csr(const fes_plan* plan, int64_t* h_indptr, int64_t* h_indices);
/* The builder's primitives (hand-written, csrc/sortscan.cuh), exposed for direct tests: a stable LSD radix
 * sort of (uint64 key, uint32 value) pairs over the low end_bit key bits, in place -- np.argsort(kind='stable')
 * of ref fem/space.py:113 -- and an exclusive prefix sum (np.cumsum of ref fem/space.py:126), d_out may alias d_in. */
int fes_sort_pairs_u64(uint64_t* d_keys, uint32_t* d_vals, int64_t n, int end_bit, void* stream);
int fes_exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, void* stream);

/* ---- element matrices (unfused) ----------------------------------------------------------
 * Replaces Form.element_matrices(geometry) (ref fem/forms.py:202-209,245-247,254-258,323-333)
 * fused with Element._affine_geometry (ref fem/elements.py:174-212).  d_coords is
 * (n_nodes, spatial_dim) row-major; spatial_dim > reference dim means embedded boundary facets
 * (mass only).  Writes d_Ke (n_el, k, k).  Parity/debug export and the K0 source of the SIMP
 * PrecomputedForm path. */
int fes_element_matrices(int elem_kind, int form, int n_comp, int spatial_dim,
                         const int32_t* d_elem_nodes, int64_t n_el, const double* d_coords,
                         const fes_coeffs* coeffs, double* d_Ke, void* stream);
/* (n_el) element measures = weight_detJ.sum(1) (ref fem/space.py:479-482) */
int fes_element_volumes(int elem_kind, int spatial_dim, const int32_t* d_elem_nodes, int64_t n_el,
                        const double* d_coords, double* d_vol, void* stream);

/* ---- fused assembly ----------------------------------------------------------------------
 * Replaces FunctionSpace.assemble(form) = form.element_matrices + _ScatterPlan.scatter
 * (ref fem/space.py:133-145,679-696,791-803).  One thread per stored node pair gathers that
 * pair's element contributions in ascending element id (np.bincount's order),
 * evaluating each element block on the fly; element matrices never touch HBM.  Writes
 * d_values (nnz) in the plan's CSR order. */
int fes_assemble(const fes_plan* plan, int elem_kind, int form, int spatial_dim,
                 const int32_t* d_elem_nodes, const double* d_coords, const fes_coeffs* coeffs,
                 double* d_values, void* stream);
/* PrecomputedForm (ref fem/forms.py:459-480): scatter caller-supplied (n_el,k,k) matrices,
 * optionally scaled per element (SIMP: rho^p * K0, ref fem/topology.py:247). */
int fes_assemble_precomputed(const fes_plan* plan, const double* d_Ke, const double* d_scale,
                             double factor, double* d_values, void* stream);
/* Host-buffer drop-in: H2D of coords (+coefficients), assembly, D2H of the values.
 * h_* coefficient pointers replace the d_* ones in `coeffs` (which must be NULL). */
int fes_assemble_host(const fes_plan* plan, int elem_kind, int form, int spatial_dim,
                      const int32_t* d_elem_nodes, const double* h_coords, int64_t n_nodes,
                      const fes_coeffs* coeffs, const double* h_E, const double* h_scale,
                      double* h_values, void* stream);

/* ---- CSR helpers ------------------------------------------------------------------------- */
/* y = A x for a device CSR (int32 indices) */
int fes_spmv(int64_t n_rows, const int32_t* d_indptr, const int32_t* d_indices, const double* d_data,
             const double* d_x, double* d_y, void* stream);
/* A.values += alpha * B.values where B's pattern is a subset of A's (the Robin boundary term
 * kappa * boundary_mass added to the volume operator, ref fem/problem.py:107-121,138-140).
 * Explicit zeros are kept (scipy's `+` prunes them; compare after eliminate_zeros). */
int fes_csr_add_subpattern(int64_t n_rows, const int32_t* d_a_indptr, const int32_t* d_a_indices,
                           double* d_a_values, const int32_t* d_b_indptr, const int32_t* d_b_indices,
                           const double* d_b_values, double alpha, void* stream);
/* d_out[i] = d_in[i]^p (SIMP dilution rho^p, ref fem/topology.py:214-224) */
int fes_pow(const double* d_in, double p, int64_t n, double* d_out, void* stream);

/* ---- Jacobi-PCG ---------------------------------------------------------------------------
 * Replaces Backend.prepare / LinearSolver.solve (ref fem/backends.py:62-116,186-213) and the
 * Dirichlet elimination of DiscreteSystem (ref fem/system.py:26-65).  The operator is a device
 * CSR over ALL dofs; `d_free_mask` (n, 1 = free, 0 = fixed; NULL = all free) restricts the
 * iteration to the free-free block without forming it.  Stopping rule, x0 = 0 and failure
 * semantics follow scipy.sparse.linalg.cg(rtol, atol=0, maxiter, M=diag^-1). */
int fes_solver_create(int64_t n, const int32_t* d_indptr, const int32_t* d_indices,
                      const double* d_data, const uint8_t* d_free_mask, void* stream,
                      fes_solver** out);
/* host int64 CSR (what Backend.prepare receives from DiscreteSystem), copied to the device */
int fes_solver_create_host(int64_t n, const int64_t* h_indptr, const int64_t* h_indices,
                           const double* h_data, fes_solver** out);
void fes_solver_destroy(fes_solver* s);
/* bytes of the operator copy ONE product streams (block-SELL values, one column index per block,
 * row-slot table): the measured side of the PCG roofline next to SURVEY 8(d)'s scalar-CSR formula */
int64_t fes_solver_operator_bytes(const fes_solver* s);
/* The operator's indptr/indices/data are the plan's CSR (values from fes_assemble): let the
 * SpMV walk node blocks (one column index and one c-wide gather per c x c block). */
int fes_solver_use_plan(fes_solver* s, const fes_plan* plan);
/* re-read d_data (same pattern) and refresh the Jacobi diagonal: a new SIMP iteration */
int fes_solver_refresh(fes_solver* s, void* stream);
/* Solve A_ff x_f = b_f.  d_b, d_x are full length n; entries at fixed dofs of d_x are left
 * untouched, of d_b ignored.  warm_start != 0 starts from the d_x passed in.  Synchronises.
 * Returns FES_ERR_CG when not converged in maxiter (maxiter <= 0 -> 10*n_free). */
int fes_solver_solve(fes_solver* s, const double* d_b, double* d_x, double rtol, int64_t maxiter,
                     int warm_start, int64_t* iters_out, double* relres_out, void* stream);
int fes_solver_solve_host(fes_solver* s, const double* h_b, double* h_x, double rtol,
                          int64_t maxiter, int64_t* iters_out, double* relres_out);
/* b_f - A_fx x_fixed: d_rhs = d_b - A * d_xfixed on free rows (ref fem/system.py:49-50) */
int fes_lift_rhs(const fes_solver* s, const double* d_b, const double* d_xfixed, double* d_rhs,
                 void* stream);

/* ---- multi-GPU (one process per GPU) -------------------------------------------
 * The operator is sharded by contiguous node blocks with ghost elements; assembly needs no
 * exchange.  The distributed PCG keeps ONE persistent kernel per rank: the halo of the direction
 * vector is written by its owners straight into the neighbours' vectors over NVLink peer memory
 * (CUDA IPC), and the dot products are summed through per-rank mailboxes in fixed rank order, so
 * every rank computes bit-identical alpha/beta.  The host only exchanges the 64-byte IPC handles
 * and the halo index lists (torch.distributed). */
int fes_comm_create(int rank, int world, int64_t max_local_dofs, fes_comm** out);
int fes_comm_handle(const fes_comm* comm, void* out64);         /* cudaIpcMemHandle_t of this rank   */
int fes_comm_connect(fes_comm* comm, const void* handles);      /* world x 64 bytes, rank order      */
void fes_comm_destroy(fes_comm* comm);
/* Rows [row_begin,row_end) of the local numbering are owned; h_send_* (dof level) list, per peer
 * rank, the owned dofs it needs and their index in ITS numbering.  n_free_global = free dofs summed
 * over the ranks' OWNED rows (identical on every rank): the default iteration cap 10 n of scipy's cg
 * (ref fem/backends.py:201-213) and the "nothing to solve" test use it, so every rank leaves the
 * collective kernel at the same iteration.  Call before the first solve.
 * Per iteration the distributed kernel posts the halo of p without a grid barrier, multiplies the
 * rows that read no ghost entry while it is in flight, and sums all dot products of the iteration in
 * ONE cross-rank reduction (single-reduction CG); FES_PCG_DIST_CLASSIC=1 selects the textbook loop
 * (halo barrier + two reductions) for A-B timing. */
int fes_solver_set_dist(fes_solver* s, fes_comm* comm, int64_t row_begin, int64_t row_end,
                        const int32_t* h_send_ptr, const int32_t* h_send_src, const int32_t* h_send_dst,
                        int64_t n_free_global);

/* ---- SIMP inner loop ----------------------------------------------------------------------
 * Replaces calculate_smoothing_matrix (ref fem/topology.py:39-77), DensityField.gradient /
 * _element_quadratic (ref fem/sensitivity.py:41-49,305-309,346-348),
 * LinearElasticForm.derived_fields + von_mises (ref fem/forms.py:335-374,
 * fem/invariants.py:48-62) and optimality_criteria_update (ref fem/design.py:39-78). */
/* centroids are the mean of each element's first `nodes_per_elem` node coordinates */
int fes_filter_create(const int32_t* d_elem_nodes, int64_t n_el, int nodes_per_elem, const double* d_coords,
                      int dim, double radius, void* stream, fes_filter** out);
void fes_filter_destroy(fes_filter* f);
int64_t fes_filter_nnz(const fes_filter* f);
int fes_filter_export_csr(const fes_filter* f, int64_t* h_indptr, int64_t* h_indices, double* h_data);
int fes_filter_apply(const fes_filter* f, const double* d_in, double* d_out, void* stream);
/* per element: q = u_e^T K0_e u_e (solid modulus E0; summed over the element's stiffness rule),
 * compliance = scale_e * sigma:eps * volume and the von Mises stress at modulus scale_e*E0, both at
 * the first quadrature point like LinearElasticForm.derived_fields (equal to scale_e * q for P1).
 * Outputs may be NULL.  d_scale = rho^p. */
int fes_elastic_energy(int elem_kind, const int32_t* d_elem_nodes, int64_t n_el, const double* d_coords,
                       const double* d_u, double E0, double nu, const double* d_scale,
                       double* d_q, double* d_compliance, double* d_von_mises, void* stream);
/* sens_e = chain * p * rho_e^(p-1) * q_e */
int fes_simp_sensitivity(const double* d_rho, const double* d_q, double penalty, double chain,
                         int64_t n_el, double* d_sens, void* stream);
/* OC bisection, whole loop on the device.  Returns FES_ERR_VALUE if any sens < 0. Synchronises. */
int fes_oc_update(const double* d_rho, const double* d_sens, const double* d_vol, int64_t n_el,
                  double volume_frac, double move, int max_iters, double tol, double* d_rho_out,
                  int* bisections_out, void* stream);
/* deterministic reductions used by the drivers: sum(a*b) (b NULL -> sum(a)) */
int fes_dot(const double* d_a, const double* d_b, int64_t n, double* h_out, void* stream);

/* ---- quadrature-tabulated forms and nodal recovery ------------------ */
/* _VectorScatterPlan.scatter (ref fem/space.py:148-184) and the np.add.at of recover_nodal
 * (ref fem/space.py:578-596): d_elem_values is (n_el, nodes_per_elem, m); d_out (n_nodes, m)
 * receives, per node, the sum over its incident (element, local node) entries in ascending
 * element id -- np.bincount's order, so the result is bit-identical to the reference's. */
int fes_node_scatter(const fes_plan* plan, const double* d_elem_values, int m, double* d_out, void* stream);
/* points of the cheapest tabulated rule of degree >= min_degree (ref fem/quadrature.py:54-114); 0 = none */
int fes_quadrature_size(int elem_kind, int min_degree);
/* ElementGeometry.points (n_el, nq, spatial_dim) and weight_detJ (n_el, nq) at that rule
 * (ref fem/elements.py:174-212); either output may be NULL. Volume elements only. */
int fes_quadrature_points(int elem_kind, int spatial_dim, int min_degree, const int32_t* d_elem_nodes,
                          int64_t n_el, const double* d_coords, double* d_points, double* d_wdet, void* stream);
/* LinearForm.element_vectors (ref fem/forms.py:296-315): d_f_qp is the field sampled at the rule's
 * points, (n_el, nq, n_comp); d_out is (n_el, nodes_per_elem, n_comp). */
int fes_element_loads(int elem_kind, int spatial_dim, int min_degree, const int32_t* d_elem_nodes, int64_t n_el,
                      const double* d_coords, const double* d_f_qp, int n_comp, double* d_out, void* stream);
/* K[(a,i),(b,j)] = delta_ij sum_q w_q kappa_q grad(phi_a)^T T grad(phi_b):  DiffusionForm (d_kappa_qp
 * (n_el, nq), d_tensor NULL; ref fem/forms.py:273-293) and GeometricStiffnessForm (d_tensor (n_el, d, d)
 * prestress, d_kappa_qp NULL; ref fem/forms.py:401-441).  d_Ke is (n_el, k, k), k = nodes * n_comp;
 * scatter it with fes_assemble_precomputed. */
int fes_weighted_gradient_matrices(int elem_kind, int spatial_dim, int n_comp, int min_degree,
                                   const int32_t* d_elem_nodes, int64_t n_el, const double* d_coords,
                                   const double* d_kappa_qp, const double* d_tensor, double* d_Ke, void* stream);
/* LinearElasticForm.derived_fields (ref fem/forms.py:335-374): (n_el, 3, 3) strain and stress at the
 * first point of the rule (plane-strain lift in 2D) and compliance = stress:strain * volume.  d_E NULL ->
 * uniform modulus E.  Outputs may be NULL. */
int fes_elastic_fields(int elem_kind, const int32_t* d_elem_nodes, int64_t n_el, const double* d_coords,
                       const double* d_u, const double* d_E, double E, double nu, int min_degree,
                       double* d_strain, double* d_stress, double* d_compliance, void* stream);
/* LinearElasticForm.stress_field (ref fem/forms.py:376-398): the in-plane Cauchy stress at EVERY point of the
 * cheapest rule of degree >= min_degree, d_stress (n_el, n_qp, sd, sd) row-major, n_qp = fes_quadrature_size;
 * no plane-strain lift.  d_E NULL -> the scalar modulus E. */
int fes_elastic_stress_field(int elem_kind, const int32_t* d_elem_nodes, int64_t n_el, const double* d_coords,
                             const double* d_u, const double* d_E, double E, double nu, int min_degree,
                             double* d_stress, void* stream);
/* FunctionSpace.gradient (ref fem/space.py:511-517): (n_el, spatial_dim) gradient of a scalar nodal field at the
 * first point of the rule (constant per element for P1). */
int fes_scalar_gradient(int elem_kind, int spatial_dim, int min_degree, const int32_t* d_elem_nodes, int64_t n_el,
                        const double* d_coords, const double* d_u, double* d_grad, void* stream);
/* _VonMisesStress (ref fem/sensitivity.py:137-199): per-element plane-strain von Mises stress at the first
 * point of the rule and d(von Mises)/d(u_e), (n_el, nodes_per_elem, 2); either output may be NULL.
 * Scatter the weighted derivative with fes_node_scatter to get the adjoint load of MeanStress /
 * SoftMaxStress (ref fem/sensitivity.py:202-276). */
int fes_von_mises_gradient(int elem_kind, const int32_t* d_elem_nodes, int64_t n_el, const double* d_coords,
                           const double* d_u, const double* d_E, double E, double nu, int min_degree,
                           double* d_von_mises, double* d_dvm_du, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FES_B200_H */
