/*
 * fem_b200.h -- C ABI of the B200-native finite-element hot path (libfem_b200.so).
 *
 * Drop-in boundary for ONE path of deepmodeling/jax-fem (all citations relative to the reference
 * tree): per-cell residual/tangent evaluation -> global sparse assembly -> Dirichlet row
 * elimination -> Jacobi-preconditioned Krylov solve -> implicit adjoint.  The reference has no
 * native FFI (it is pure Python on XLA/PETSc/SciPy); the entry points below are what a
 * `jax.ffi` / ctypes binding for that path would bind, one per Python seam that is replaced.
 * INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; plain pointers and sizes
 *     only, no library types.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - all functions are stream-ordered, allocate nothing, keep no global mutable state except the
 *     thread-local error string, and return 0 on success or a negative FEM_E* code.
 *   - float64 values, int32 indices (PETSc.IntType of the reference, jax_fem/solver.py:474-475).
 *   - there is NO CPU fallback: every entry point fails with FEM_ENODEV when no CUDA device exists.
 */
#ifndef FEM_B200_H
#define FEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEM_OK 0
#define FEM_EINVAL (-1)      /* bad argument / unregistered element-law combination            */
#define FEM_ECUDA (-2)       /* CUDA runtime error (see fem_last_error)                         */
#define FEM_ENODEV (-3)      /* no CUDA device: the product path never falls back to the CPU    */

/* element types (jax_fem/basis.py:52-57, 58-65, 86-91) */
#define FEM_ELE_HEX8 0
#define FEM_ELE_QUAD4 1
#define FEM_ELE_HEX27 2

/* registered constitutive laws == the reference's get_tensor_map bodies          */
#define FEM_LAW_POISSON 0         /* params: k                       ; optional per-quad scale   */
#define FEM_LAW_LINEAR_ELASTIC 1  /* params: E, nu, plane_stress(0/1, 2-D elements only)          */
#define FEM_LAW_NEO_HOOKEAN 2     /* params: E, nu, clamp_J(0/1)     ; optional per-quad rho     */
#define FEM_LAW_SIMP 3            /* params: Emax, Emin, nu, penal, plane_stress(0/1, 2-D only); REQUIRED per-quad theta */

const char* fem_last_error(void);
int fem_version(void);
/* number of CUDA devices visible, or FEM_ENODEV */
int fem_device_count(void);

/* ---- (0) the assembly plan: the (I, J) COO pattern of Problem.__post_init__ (jax_fem/problem.py:86-107) and PETSc's
 *      setPreallocationCOO (jax_fem/solver.py:476-478), built ON THE DEVICE from the raw connectivity: sparsity pattern
 *      (int32 indptr / indices exactly as PETSc produces them: every (I, J) pair kept, columns ascending), node-block graph,
 *      the node-sorted corner order of the element tangents, the deterministic gather schedule and the transpose map.
 *      The handle owns its tables (cudaMalloc); no Python and no torch are involved.
 * cells: device (n_cells, nodes_per_cell) int32.  Fails with FEM_EINVAL when nnz or n_cells*N*N exceed int32 (shard the
 * mesh) or a node belongs to more than 16 cells (8 for 27-node cells).
 * fem_plan_sizes: [0] nnzb (node-block entries) [1] nnz [2] gather work items (n_blocks of fem_gather_csr) [3] emeta rows
 *                 [4] sources (= n_cells*N*N) [5] n_dofs [6] doubles per corner row block of Ke [7] row blocks of Ke.
 * fem_plan_table: device pointer + length of one table (FEM_PLAN_*); the tables are the arguments of the entry points below
 *                 (corner_pos -> fem_element_residual_jacobian; gdesc, src -> fem_gather_csr; nc_ptr, nc -> fem_gather_residual;
 *                 indptr, indices, brow_ptr, bcol -> fem_spmv / fem_pcg / fem_pbicgstab; tperm -> fem_csr_transpose_values).
 * fem_plan_entry_meta: the emeta argument of fem_gather_csr (n_rows x 4 int32, 16-byte aligned) for a Dirichlet mask
 *                 bc_flag (n_dofs bytes, 1 = Dirichlet row; NULL = none): rows -> zeroed, unit diagonal, pattern kept
 *                 (Mat.zeroRows, jax_fem/solver.py:477,527-528).                                                        */
#define FEM_PLAN_BROW_PTR 0
#define FEM_PLAN_BCOL 1
#define FEM_PLAN_INDPTR 2
#define FEM_PLAN_INDICES 3
#define FEM_PLAN_CORNER_POS 4
#define FEM_PLAN_NC_PTR 5
#define FEM_PLAN_NC 6
#define FEM_PLAN_GDESC 7
#define FEM_PLAN_SRC 8
#define FEM_PLAN_SRC_PTR 9
#define FEM_PLAN_TPERM 10
#define FEM_PLAN_M_SB 11
#define FEM_PLAN_M_SE 12
#define FEM_PLAN_M_ENT 13
#define FEM_PLAN_M_ADD 14
#define FEM_PLAN_EDST 15
#define FEM_PLAN_EROW 16
int fem_plan_create(const int32_t* cells, int64_t n_cells, int64_t n_nodes, int nodes_per_cell, int vec,
                    void* stream, void** plan_out);
int fem_plan_destroy(void* plan);
int fem_plan_sizes(const void* plan, int64_t* sizes_host);
int fem_plan_table(const void* plan, int which, const int32_t** table_out, int64_t* count_out);
int fem_plan_entry_meta(const void* plan, const uint8_t* bc_flag, int32_t* emeta, void* stream);

/* ---- (1) element kernels: Problem.compute_residual_vars / compute_newton_vars
 *      (jax_fem/problem.py:439-460) with the geometry of FiniteElement.get_shape_grads
 *      (jax_fem/fe.py:112-141) recomputed in-kernel instead of materialised.
 *
 * ref_tables: [NQ*NN*DIM] reference shape gradients (q,n,d), then [NQ] quadrature weights
 *             (jax_fem/basis.py:141-175), device memory.
 * internal_var: (n_cells, NQ) per-quadrature-point parameter (problem.internal_vars[0],
 *             jax_fem/problem.py:125,493-556) or NULL.
 * Ke: (n_cells*NN, NN, vec, vec): for every corner (cell c, local node a) the row block
 *     d r_(a,i) / d u_(b,k), b = 0..NN-1, stored at row-block position corner_pos[c*NN + a]
 *     (corner_pos == NULL: c*NN + a, which is exactly the reference's problem.V cell blocks,
 *     problem.py:265,453).  The assembly plan passes its node-sorted corner order so that all row blocks
 *     of a mesh node are adjacent ("COO sorted by row").  Ke == NULL: residual only.
 * Re: (n_cells, ndof) element residuals (weak_form_flat, problem.py:443).
 */
int fem_element_residual_jacobian(int ele_type, int vec, int law_id, const double* law_params_host,
                                  const double* points, const int32_t* cells, int64_t n_cells,
                                  const double* sol, const double* internal_var,
                                  const double* ref_tables, const int32_t* corner_pos, double* Ke, double* Re,
                                  void* stream);

/* HEX27 (jax_fem/basis.py:58-65; 81x81 element tangents on FP64 DMMA tiles), isotropic elasticity (linear, SIMP).
 * ref_tables: [n_quad*27*3] dN then [n_quad] weights; ref_tables_t: the same dN as [27*3][n_quad] (point index fastest:
 * coalesced reads in the per-point phase) or NULL; internal_var (n_cells, n_quad) or NULL.
 * affine_tables (optional, with affine_work): [27][3] reference coordinates xi_n of the 27 nodes (unit cube) followed by
 * [3][3][27][27] Ghat[e][f][a][b] = sum_q w_q dN_a^e(q) dN_b^f(q).  When given, a first pass tests every cell for an affine
 * geometry map (X_n = X_0 + J xi_n for all 27 nodes, J from the corners xi = 0, e_1, e_2, e_3) and a cell-constant internal
 * variable, and forms K_e = E detJ J^-T Ghat J^-1 and R_e = K_e u_e for those cells (exact up to rounding; 15x fewer FLOPs);
 * the remaining cells are listed for the general DMMA kernel.  affine_work: device workspace of at least
 * 100 * n_cells + 16 bytes, 16-byte aligned ([n_cells][12] doubles of per-cell records, then int32 count + cell ids of the
 * list).  Both NULL: general kernel for every cell.
 * Ke: (n_cells*27, 244): row block of corner (c,a) = 27 blocks of 3x3 (+1 pad double) at corner_pos[c*27+a];
 * NULL = residual only.  Re: (n_cells, 81).                                                                    */
int fem_hex27_residual_jacobian(int law_id, const double* law_params_host, const double* points,
                                const int32_t* cells, int64_t n_cells, const double* sol,
                                const double* internal_var, const double* ref_tables, const double* ref_tables_t,
                                int n_quad, const double* affine_tables, void* affine_work,
                                const int32_t* corner_pos, double* Ke, double* Re, void* stream);

/* ---- (1)+(2) fused: compute_newton_vars (jax_fem/problem.py:447-460) + _PetscTangentCache.update / get_A
 *      (jax_fem/solver.py:469-553) + compute_residual_vars_helper (jax_fem/problem.py:426-437) in one kernel for
 *      HEX8 / vec 3 / isotropic elasticity (linear, SIMP): the element tangents never reach HBM.
 *
 * The mesh nodes are partitioned into patches of <= 64 owned nodes; one CTA evaluates every cell touching its
 * patch, accumulates the rows of its own nodes in shared memory in ascending cell order (deterministic, no atomics)
 * and writes each CSR value once.  Tables (jax_fem_b200/patch_plan.py documents the bit layouts):
 *   phdr (n_patches+1, 8): first owned node / local node / patch-cell / chunk, accumulator doubles
 *   pn_node, pn_out, pn_acc, pn_info: per owned node: global id, offset of its rows in `data` (= 9 brow_ptr[n]),
 *            offset in the patch accumulator, len(n) | diagonal slot << 8
 *   lnodes: per patch the global ids of its local nodes (owned first);  pc_cell / pc_ln: per (patch, cell) the
 *            global cell id and the 8 local node numbers (uint8 x 8);  pc_lm: per (patch, cell) its first lane and
 *            the mask of corners the patch owns
 *   ck_cell / ck_lane / ck_rnd: per chunk (<= 32 or 16 cells, see config) the first patch-cell, the first lane and
 *            the number of accumulation rounds
 *   ln_desc / ln_slot: per owned corner: cell-in-chunk | a<<5 | owned index<<8 | round<<16, and the 8 column slots
 * bc_flag: (3 n_nodes) bytes, 1 = Dirichlet row (zeroed, unit diagonal, pattern kept; solver.py:477,527-528).
 * f_ext: (n_nodes, 3) constant load vector or NULL.  Outputs: data (nnz) CSR values, res (n_nodes, 3).
 * config: which (patch size, chunk size) the tables were built for (patch_plan.CONFIGS index).                  */
int fem_assemble_fused(int ele_type, int vec, int law_id, const double* law_params_host,
                       const double* points, const double* sol, const double* internal_var,
                       const double* ref_tables, int64_t n_patches, const int32_t* phdr,
                       const int32_t* pn_node, const int32_t* pn_out, const int32_t* pn_acc,
                       const int32_t* pn_info, const int32_t* lnodes, const int32_t* pc_cell,
                       const int32_t* pc_ln, const int32_t* pc_lm, const int32_t* ck_cell, const int32_t* ck_lane, const int32_t* ck_rnd,
                       const int32_t* ln_desc, const int32_t* ln_slot, const uint8_t* bc_flag,
                       const double* f_ext, double* data, double* res, int config, void* stream);

/* ---- (1)+(2) staged in one kernel: compute_newton_vars (jax_fem/problem.py:447-460) + _PetscTangentCache.update / get_A
 *      (jax_fem/solver.py:469-553) for HEX8 / vec 3 / isotropic elasticity (linear, SIMP) as work items of ONE persistent
 *      kernel.  E item i evaluates cells_p[16 i .. 16 i + 16) (the plan's processing order; corder = their cell ids for
 *      internal_var / Re) and writes the row block of corner (slot, a) to staging row dest_row[8 slot + a] once the G
 *      item prev_g[8 slot + a] (-1: none) that read the row's previous occupant is done; G item g sums the CSR rows of its
 *      nodes (the fem_gather_csr work item) once the E items holding the cells around its nodes are done.
 *      tdesc (n_e + n_gather, 32): one descriptor per ticket, in ticket order: [0] E: 0x80000000 | item, G: item; G only:
 *      [1] first staging row, [2..5] first corner / entry / source / emeta row, [6..9] their ends, [10] number of E items
 *      it waits for, [11] -1 or the offset of their list in gdep when they do not fit, [12..31] those E items.
 *      The staging buffer `stage` (rows of 72 doubles, 128-byte aligned) is a ring that stays in L2 plus a spill area;
 *      jax_fem_b200/stage_plan.py builds the tables and checks that every wait is on an EARLIER ticket (no deadlock).
 *      Staging rows hold G = sum_q E_q w_q g_a (x) g_b as 9 tiles [I][J][b]; K = lam' G + mu' G^T + mu' tr(G) I is
 *      applied after the sum.  ctrl: fem_staged_ctrl_ints(n_e, n_gather) int32 of scratch (zeroed by the call).
 *      Outputs: data (nnz) CSR values with Dirichlet rows treated (emeta, see fem_gather_csr), Re (n_cells, 24).
 *      fem_staged_status: *status_host != 0 after a call whose waits exceeded their limit (plan bug; values invalid).  */
int64_t fem_staged_ctrl_ints(int64_t n_e, int64_t n_gather);
int fem_assemble_staged(int law_id, const double* law_params_host, const double* points, const double* sol,
                        const double* internal_var, const double* ref_tables, int64_t n_cells,
                        const int32_t* cells_p, const int32_t* corder, const int32_t* dest_row,
                        const int32_t* prev_g, int64_t n_gather, const int32_t* tdesc, const int32_t* gdep,
                        const int32_t* emeta, const int32_t* src, double* stage,
                        int32_t* ctrl, double* Re, double* data, void* stream);
int fem_staged_status(const int32_t* ctrl, int32_t* status_host, void* stream);

/* Plan construction helper (HOST pointers, no CUDA): greedy split of every patch's cells into chunks of <= chunk
 * cells in which no owned node occurs more than rmax times.  cell_ptr (n_patches+1); owned_idx (M, nodes_per_cell):
 * owned-node index of the corner or 255; outputs: chunk of every patch-cell (inside its patch), round of every
 * corner (255 = not owned), chunks per patch.                                                                    */
int fem_patch_chunks_host(int64_t n_patches, const int64_t* cell_ptr_host, const uint8_t* owned_idx_host,
                          int nodes_per_cell, int max_owned, int chunk, int rmax,
                          int32_t* cell_chunk_host, uint8_t* rank_host, int32_t* n_chunks_host);

/* ---- (2) global assembly: _PetscTangentCache.update / get_A (jax_fem/solver.py:469-553)
 *      as a precomputed cell->CSR-slot permutation + deterministic segmented sum (no atomics).
 *
 * The node-block graph has one "entry" per (row node n, neighbour node m), rows ascending, columns
 * ascending.  src_ptr (nnzb+1) / src (n_items): for entry e the Ke block index corner_pos[c*NN+a]*NN + b of
 * every (cell, local row node, local col node) contributing to it, in ascending (c,a,b) order (fixed
 * summation order => bit-reproducible).  Ke is the output of fem_element_residual_jacobian.
 * gdesc (4*(n_blocks+1)): work split.  CTA b owns the nodes whose first corner (in node-sorted order) lies in
 *              [W b, W (b+1)), W = 32 corners (8 for 27-node cells); gdesc[4b..4b+3] = first corner, first entry,
 *              first source, first emeta row of item b (the next item's quadruple closes the ranges).  No node may have more than
 *              16 corners (8 for 27-node cells).  Row blocks are padded to an even number of doubles.
 * emeta (n_rows, 4): one row per entry, two for an entry with more than 4 sources (the second, flagged by bit 20 of
 *              [3], covers the second half of the sources and is added to the first half's sum), IN PROCESSING ORDER
 *              (inside each CTA's range the rows are sorted by descending source count so that the lanes of a warp
 *              loop equally long; results do not depend on the order):
 *              [0],[1] = its source range in `src`, [2] = offset in `data` of element (row vec*n, col vec*m) of the
 *              scalar CSR pattern (indptr[vec*n+i] = vec*vec*brow_ptr[n] + i*vec*len(n)), [3] = bits 0..15 vec*len(n)
 *              (distance between the entry's consecutive scalar rows), bit 16 = m == n, bit 17+i = row vec*n+i is a
 *              Dirichlet row -> row zeroed, unit diagonal, pattern kept (Mat.zeroRows, solver.py:477,527-528).
 *              16-byte aligned: the kernel fetches one int4 per entry.
 */
int fem_gather_csr(int vec, int nn, int64_t n_blocks, const int32_t* gdesc, const int32_t* emeta,
                   const int32_t* src, const double* Ke, double* data, void* stream);

/* The same two steps for HEX8 / vec 3 / {linear elasticity, SIMP, Neo-Hookean} with the element tangents staged in
 * TILE-MAJOR rows: the FP64 tensor-core accumulator fragments are stored straight from registers (row of corner (c, a) =
 * 9 tiles (I, J) of 8 doubles b = 0..7, at corner_pos[c*8 + a], 16-byte aligned), and for the isotropic laws the rows hold
 * G = sum_q E_q w_q g_a (x) g_b; K = lam' G + mu' G^T + mu' tr(G) I is applied by the gather after the sum over the cells.
 * post_host[3] (written by fem_element_tiles, passed to fem_gather_csr_tiles): {apply the map (0/1), lam', mu'}.
 * All other arguments as in fem_element_residual_jacobian / fem_gather_csr.                                               */
int fem_element_tiles(int law_id, const double* law_params_host, const double* points, const int32_t* cells,
                      int64_t n_cells, const double* sol, const double* internal_var, const double* ref_tables,
                      const int32_t* corner_pos, double* Ke_tiles, double* Re, double* post_host, void* stream);
int fem_gather_csr_tiles(int64_t n_blocks, const int32_t* gdesc, const int32_t* emeta, const int32_t* src,
                         const double* Ke_tiles, double* data, const double* post_host, void* stream);

/* residual scatter-add of problem.py:426-437 as a per-node gather: nc_ptr (n_nodes+1) / nc codes
 * c*NN + a ascending; res = sum Re + f_ext (f_ext may be NULL).                                  */
int fem_gather_residual(int vec, int nn, int64_t n_nodes, const int32_t* nc_ptr, const int32_t* nc,
                        const double* Re, const double* f_ext, double* res, void* stream);

/* ---- (1c) solution-dependent mass maps: get_mass_kernel (jax_fem/problem.py:216-236) and its share of value_and_jacfwd
 *      (problem.py:262-266) for the registered mass law m_i(u) = a u_i + b_i with a = coef or coef_field (n_cells, n_quad)
 *      and b = const_host[vec] or const_field (n_cells, n_quad, vec): backward-Euler heat capacity, phase-field driving term,
 *      elastic foundation.  Runs between the element kernel and the gathers and ADDS, in place, to the element residuals Re
 *      (n_cells, nn*vec) and to the staged row blocks Ke in the REFERENCE block layout of fem_element_residual_jacobian (not
 *      the tile-major rows of fem_element_tiles); Ke == NULL: residual only.  shape_vals: [n_quad][nn] (basis.py:141-175).
 *      HEX8 (vec 1, 3), QUAD4 (vec 1, 2) and HEX27 (vec 3; up to 216 points); one thread per (cell, node) owns its row block:
 *      deterministic, no atomics.                                                                                        */
int fem_mass_term(int ele_type, int vec, const double* points, const int32_t* cells, int64_t n_cells, const double* sol,
                  const double* ref_tables, const double* shape_vals, int n_quad, double coef, const double* coef_field,
                  const double* const_host, const double* const_field, const int32_t* corner_pos, double* Ke, double* Re,
                  void* stream);

/* ---- (1b) solution-dependent surface maps: get_surface_kernel (jax_fem/problem.py:238-259) and its tangent
 *      (problem.py:289-325) for the registered surface law val_i(u) = coef_i (u_i - uref_i)^power (Robin / convection / spring
 *      foundation; applications/robin_bc/example.py:59-67 is coef 5, power 2).  law_host[7] = coef[3], uref[3], power.
 *      One boundary set per call: its F faces (face_cell, face_lid = local face of the cell), nanson (F, fq) = Nanson scale x
 *      face weight (fe.py:180-182, precomputed: independent of u), face_vals (local faces, fq, nn) = face shape values
 *      (basis.py:178-250); bnode: the set's boundary nodes with their incident (face, local node) pairs in bf_ptr / bf_face /
 *      bf_local, ascending face order.  fem_face_residual ADDS the face residual to res (nodes, vec); fem_face_tangent ADDS the
 *      face tangent to the assembled CSR values (rows flagged in bc_flag stay unit rows).  Each output is written by one
 *      thread in a fixed order: deterministic, no atomics.                                                            */
int fem_face_residual(int vec, int nn, int fq, int64_t n_bnodes, const int32_t* bnode, const int32_t* bf_ptr,
                      const int32_t* bf_face, const int32_t* bf_local, const int32_t* face_cell,
                      const int32_t* face_lid, const double* nanson, const double* face_vals, const int32_t* cells,
                      const double* sol, const double* law_host, double* res, void* stream);
int fem_face_tangent(int vec, int nn, int fq, int64_t n_bnodes, const int32_t* bnode, const int32_t* bf_ptr,
                     const int32_t* bf_face, const int32_t* bf_local, const int32_t* face_cell,
                     const int32_t* face_lid, const double* nanson, const double* face_vals, const int32_t* cells,
                     const double* sol, const double* law_host, const int32_t* brow_ptr, const int32_t* bcol,
                     const uint8_t* bc_flag, double* data, void* stream);

/* ---- (3) Dirichlet operations (jax_fem/solver.py:290-363) on merged (last-wins) row lists    */
/* apply_bc_vec: res[row] = sol[row] - val*scale                                                  */
int fem_apply_bc_vec(int64_t n_bc, const int32_t* bc_rows, const double* bc_vals, double scale,
                     const double* sol, double* res, void* stream);
/* x0 = assign_bc(0) - copy_bc(dofs): x0 = 0, x0[row] = val - dofs[row]   (solver.py:402-409)     */
int fem_bc_initial_guess(int64_t n, int64_t n_bc, const int32_t* bc_rows, const double* bc_vals,
                         const double* dofs, double* x0, void* stream);

/* ---- (3) sparse kernels and Krylov solvers: jax_solve (jax_fem/solver.py:63-92)               */
/* Optional node-block structure (vec, brow_ptr, bcol) of the FE matrix: when given (vec = 2 or 3) the kernels walk
 * the neighbour list of a mesh node once for its vec scalar rows and read one column index per vec x vec block
 * instead of `indices` (same result up to summation order); pass vec = 1 / NULL for a general CSR matrix.       */
int fem_spmv(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, int vec,
             const int32_t* brow_ptr, const int32_t* bcol, const double* x, double* y, void* stream);
/* diag[i] = A[i,i] (jacobi = A.diagonal(), solver.py:68)                                          */
int fem_csr_diagonal(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data,
                     double* diag, void* stream);
/* transpose of a structurally symmetric node-block CSR (A.transpose(A_T), solver.py:1407-1408):
 * tperm (nnzb) maps block entry (n,m) -> entry (m,n).                                             */
int fem_csr_transpose_values(int vec, int64_t n_nodes, const int32_t* brow_ptr, const int32_t* bcol,
                             const int32_t* tperm, const double* data, double* data_t, void* stream);

/* workspace (in doubles) for the solvers below */
int64_t fem_krylov_workspace(int64_t n);

/* Jacobi-preconditioned CG / BiCGSTAB, whole solve on the device, same recurrences and stopping
 * rule as jax.scipy.sparse.linalg.cg / bicgstab: stop when ||r||^2 <= max(tol^2 ||b||^2, atol^2).
 * x holds x0 on entry and the solution on exit.  diag = NULL means no preconditioner.
 * info_host[0] = iterations (negative on breakdown as in JAX), info_host[1] = final ||r||^2,
 * info_host[2] = ||A x - b|| (the reference's post-check, solver.py:87).
 * check_every: iterations between host polls of the device-side convergence flag.               */
int fem_pcg(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, int vec,
            const int32_t* brow_ptr, const int32_t* bcol, const double* diag, const double* b, double* x,
            double tol, double atol, int maxiter, int check_every, double* workspace, double* info_host, void* stream);
int fem_pbicgstab(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, int vec,
                  const int32_t* brow_ptr, const int32_t* bcol, const double* diag, const double* b, double* x,
                  double tol, double atol, int maxiter, int check_every, double* workspace, double* info_host,
                  void* stream);

/* ---- (e) multi-GPU: one rank's share of the Jacobi-CG above (cells sharded across ranks).
 *      Vectors hold the rank's owned dofs first, then its ghosts.  Every kernel leaves rank-local partial sums in
 *      workspace[16..20); the caller all-reduces that slice (ncclAllReduce on the same stream) and exchanges the
 *      ghost entries of p (ncclSend/ncclRecv) between the steps:
 *        begin -> spmv_dot(x0, with_dot=0) -> init -> [allreduce] -> scalars(0)
 *        loop:  [halo p] -> spmv_dot(p, 1) -> [allreduce] -> update -> [allreduce] -> direction -> scalars(1)
 *      workspace[7] != 0 once converged (same stopping rule as fem_pcg); workspace[6] = iterations.            */
int fem_dcg_begin(double* workspace, double tol, double atol, int maxiter, void* stream);
int fem_dcg_spmv_dot(int64_t n_owned, int64_t n_local, const int32_t* indptr, const int32_t* indices,
                     const double* data, int vec, const int32_t* brow_ptr, const int32_t* bcol,
                     const double* p, double* q, int with_dot, double* workspace, void* stream);
int fem_dcg_init(int64_t n_owned, int64_t n_local, const double* b, const double* diag, const double* q,
                 double* r, double* p, double* workspace, void* stream);
int fem_dcg_scalars(int phase, double* workspace, void* stream);
int fem_dcg_update(int64_t n_owned, int64_t n_local, const double* diag, const double* p, const double* q,
                   double* x, double* r, double* workspace, void* stream);
int fem_dcg_direction(int64_t n_owned, const double* diag, const double* r, double* p, double* workspace,
                      void* stream);

/* ---- (e) multi-GPU, whole loop in the library: NCCL plumbing, halo exchange, distributed Jacobi-CG / BiCGSTAB.
 *      Replaces what the reference's MPI demo does through PETSc (applications/parallel/poisson_mpi.py:123-164,214-437)
 *      and, per rank, jax_solve (jax_fem/solver.py:63-92) / the adjoint solve (solver.py:1409).
 *
 *      libnccl.so.2 is resolved at run time (dlopen: the copy the process already loaded, e.g. PyTorch's; FEM_NCCL_LIB
 *      overrides), so the library has no link-time NCCL dependency.  `comm` arguments are ncclComm_t passed as void*.
 *      fem_nccl_unique_id / fem_nccl_comm_create wrap ncclGetUniqueId / ncclCommInitRank (current CUDA device) for callers
 *      that do not have a communicator yet; a communicator made elsewhere (any ncclComm_t) is accepted as well.
 *
 *      Halo plan of one rank: local vectors hold the owned dofs first, then the ghosts grouped by owner rank.  Neighbour k
 *      (rank peer_host[k]) receives the owned nodes send_idx[send_ptr_host[k] .. send_ptr_host[k+1]) (device array of
 *      local node ids) and fills the local nodes [recv_start_host[k], recv_start_host[k] + recv_count_host[k]).  sendbuf:
 *      device scratch of vec * send_ptr_host[n] doubles.  fem_halo_exchange = one pack kernel + ncclGroupStart /
 *      ncclSend / ncclRecv / ncclGroupEnd on `stream`.  fem_allreduce_sum: ncclAllReduce(sum, double) in place.          */
int fem_nccl_unique_id(void* id128_host);
int fem_nccl_comm_create(int world, int rank, const void* id128_host, void** comm_out);
int fem_nccl_comm_destroy(void* comm);
int fem_halo_create(void* nccl_comm, int vec, int n_neighbours, const int32_t* peer_host,
                    const int64_t* send_ptr_host, const int32_t* send_idx, const int64_t* recv_start_host,
                    const int64_t* recv_count_host, double* sendbuf, void** halo_out);
int fem_halo_destroy(void* halo);
int fem_halo_exchange(void* halo, double* x, void* stream);

/* Owned nodes [node_lo, node_hi) of the rank have no ghost neighbour: fem_dist_pcg / fem_dist_pbicgstab multiply their rows
 * while the halo exchange of the product's input is in flight on the plan's own stream (pack -> event -> ncclSend/ncclRecv on
 * that stream -> event), and the remaining rows once the ghosts have arrived; the rank-local dot of the product is completed
 * by the second launch.  Used when more than half of the owned nodes are in the range and the matrix has its node-block
 * structure (vec 2 or 3); an empty range (the default) keeps exchange and product in sequence.                          */
int fem_halo_set_interior(void* halo, int64_t node_lo, int64_t node_hi);

/* Peer-memory halo exchange for the distributed Krylov loops (ranks on one node, NVLink / NVSwitch): instead of ncclSend /
 * ncclRecv the pack kernel stores every interface value straight into the neighbour's mailbox and the last block raises a flag
 * there (st.release.sys); the receiver's unpack kernel waits for its flags (bounded spin: an error, never a hang) and copies
 * the ghost blocks into the vector.  The all-reduce that follows every product in CG / BiCGSTAB orders consecutive exchanges,
 * and the two receive buffers alternate.  Setup, once per halo plan:
 *   fem_halo_p2p_alloc    every rank: mailbox of 256 + 2 * n_ghost_nodes * vec * 8 bytes (cudaMalloc) and its 64-byte CUDA IPC
 *                         handle, to be handed to the neighbours by any host-side means;
 *   fem_halo_p2p_connect  per neighbour k: map its mailbox (cudaIpcOpenMemHandle); peer_ghost_offset_nodes = where this rank's
 *                         block starts inside the neighbour's ghost range, my_slot_in_peer = this rank's index in the
 *                         neighbour's peer list (its flag word);
 *   fem_halo_p2p_enable   after a barrier, on every rank or on none.
 * fem_halo_exchange itself keeps using NCCL (no ordering guarantee between two stand-alone exchanges).                    */
int fem_halo_p2p_alloc(void* halo, int64_t n_owned_nodes, int64_t n_ghost_nodes, void* ipc_handle64_out);
int fem_halo_p2p_connect(void* halo, int k, const void* peer_ipc_handle64, int64_t peer_n_ghost_nodes,
                         int64_t peer_ghost_offset_nodes, int my_slot_in_peer);
int fem_halo_p2p_enable(void* halo, int on);
int fem_allreduce_sum(void* halo, double* buf, int count, void* stream);
/* Distributed Jacobi-CG / BiCGSTAB on the rank's owned rows (CSR rows 0..n_owned-1 complete, columns index the local
 * vector of n_local entries).  Same recurrences, stopping rule and info_host as fem_pcg / fem_pbicgstab; per iteration
 * CG issues 1 halo exchange + 2 all-reduces (1 and 2 doubles), BiCGSTAB 2 halo exchanges + 4 all-reduces, all on
 * `stream`; the host only polls the (all-reduced, hence rank-uniform) convergence flag every check_every iterations.
 * b, x: local vectors (x: x0 on entry, solution with up-to-date ghosts on exit); workspace: fem_krylov_workspace(n_local). */
int fem_dist_pcg(void* halo, int64_t n_owned, int64_t n_local, const int32_t* indptr, const int32_t* indices,
                 const double* data, int vec, const int32_t* brow_ptr, const int32_t* bcol, const double* diag,
                 const double* b, double* x, double tol, double atol, int maxiter, int check_every,
                 double* workspace, double* info_host, void* stream);
int fem_dist_pbicgstab(void* halo, int64_t n_owned, int64_t n_local, const int32_t* indptr, const int32_t* indices,
                       const double* data, int vec, const int32_t* brow_ptr, const int32_t* bcol, const double* diag,
                       const double* b, double* x, double tol, double atol, int maxiter, int check_every,
                       double* workspace, double* info_host, void* stream);

/* ---- (4) implicit adjoint: -lambda^T dc/dtheta per quadrature point
 *      (jax_fem/solver.py:1386-1394,1414-1416) for per-quad parameters; lambda must already be
 *      zero on Dirichlet rows (BC rows of c do not depend on theta).  grad: (n_cells, NQ).       */
int fem_adjoint_param_grad(int ele_type, int vec, int law_id, const double* law_params_host,
                           const double* points, const int32_t* cells, int64_t n_cells,
                           const double* sol, const double* internal_var, const double* lam,
                           const double* ref_tables, double* grad, void* stream);

/* The same for HEX27 + SIMP (the 216- or 27-point rules of basis.py:58-65): grad (n_cells, n_quad). */
int fem_hex27_adjoint_param_grad(int law_id, const double* law_params_host, const double* points, const int32_t* cells,
                                 int64_t n_cells, const double* sol, const double* internal_var, const double* lam,
                                 const double* ref_tables, int n_quad, double* grad, void* stream);

/* small helpers used by the Newton loop */
int fem_dot(int64_t n, const double* x, const double* y, double* result_host, double* workspace, void* stream);
int fem_axpy(int64_t n, double alpha, const double* x, double* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEM_B200_H */
