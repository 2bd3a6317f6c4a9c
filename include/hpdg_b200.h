/*
 * hpdg_b200.h -- C ABI of the B200-native SIPG hot path (drop-in boundary).
 *
 * One opaque context per GPU / rank.  Vectors are plain `double` arrays in the layout of the
 * reference's DynamicBlockVector (dune/hpdg/common/dynamicbvector.hh:366-379): one contiguous
 * array, element block i of (p_i+1)^dim doubles at offset sum_{j<i} (p_j+1)^dim, local index
 * x-fastest (dune/hpdg/localfunctions/lagrange/qkgausslobatto/qkgllocalbasis.hh:69-78); elements
 * numbered x-fastest like YaspGrid's leaf index.
 *
 * Every entry point returns 0 on success, non-zero on failure; hpdg_last_error() then describes the
 * failure (the C++ shim include/hpdg_b200.hh rethrows it as an exception, mirroring DUNE_THROW).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Entry points come in two flavours: `*_device` take device pointers (vectors resident in HBM, the
 * normal mode inside a solver loop) and the plain names take HOST pointers and stage the copies
 * (the literal drop-in for code holding DUNE containers).  Calls on one context are serialised by
 * the caller, like the reference's non-reentrant local operators (ipdgoperator.hh:233,397-401).
 */
#ifndef HPDG_B200_H
#define HPDG_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hpdg_ctx hpdg_ctx;

#define HPDG_FINEST (-1)          /* level argument: finest level */
#define HPDG_JACOBI_DENSE 0       /* precomputed dense per-element inverses, batched by block size */
#define HPDG_JACOBI_FD 1          /* same inverse in Kronecker (fast diagonalisation) form */
#define HPDG_SMOOTHER_BLOCKGS 2    /* hpdg_vcycle only: the reference's DynamicBlockGS on assembled level matrices */
#define HPDG_SMOOTHER_BLOCKGS_MF 3 /* hpdg_vcycle only: the same DynamicBlockGS sweeps without a matrix (any mesh size) */
#define HPDG_PRECOND_NONE 0       /* hpdg_pcg: plain CG */
#define HPDG_PRECOND_JACOBI 1     /* hpdg_pcg: fd block-Jacobi preconditioner */
#define HPDG_PRECOND_VCYCLE 2     /* hpdg_pcg: one p-multigrid V-cycle with fd block-Jacobi smoothing */

/* -- problem description -------------------------------------------------------------------------
 * Replaces: DynamicDGQkGLBlockBasis(gridView, k | degree map) (functionspacebases/dynamicdgqkglbasis.hh:54-69),
 * YaspGrid extents (matrix-free/test/testdg.cc:144) and IPDGOperator(basis, penalty, dirichlet)
 * (matrix-free/localoperators/ipdgoperator.hh:40).  degree: ndegree == 1 -> uniform, else one entry per
 * element (ndegree == n[0]*..*n[dim-1]).  Domain is [0,L[0]] x .. ; dim = 2 or 3. */
int hpdg_create(hpdg_ctx** out, int dim, const int* n, const double* L, const int* degree, long ndegree,
                double sigma, int dirichlet, int device);

/* Rank-local brick of a structured mesh split over pgrid[0]*pgrid[1]*pgrid[2] ranks (rank = px + pgrid[0]*(py +
 * pgrid[1]*pz)); n and L describe the LOCAL brick.  Models the reference's element-wise owner/overlap
 * decomposition (parallel/communicationhpdg.hh:261-263) with a face-trace halo instead of whole ghost
 * blocks.  nccl_id: the 128-byte ncclUniqueId produced by hpdg_nccl_unique_id on rank 0 and broadcast by the
 * caller; NULL = no NCCL communicator (then the peer-memory halo below must be attached before the first apply, and the entry
 * points that need NCCL -- dot products over ranks, NCCL halo, distributed V-cycle -- fail).  Uniform degree only. */
int hpdg_create_distributed(hpdg_ctx** out, int dim, const int* n, const double* L, int degree, double sigma,
                            int dirichlet, int device, const int* pgrid, int rank, int nranks,
                            const void* nccl_id);
/* Non-conforming 2-D mesh (hanging nodes; reference: the non-conforming branch of SumFactIPDGOperator,
 * matrix-free/localoperators/sfipdg.hh:213-222,472-491): the n[0] x n[1] base grid with the cells flagged in refine[] (x fastest,
 * 0 / 1) split once into 2 x 2 children.  Leaf elements are numbered base cell by base cell (x fastest), the four children of a
 * refined cell x fastest; degree[] and every vector hold one entry / block per leaf in that order.  Available on such a context:
 * the operator apply (hpdg_op_apply*, accumulate forms, BLAS-1); smoothers, assembly and the p-hierarchy report an error. */
int hpdg_create_refined_2d(hpdg_ctx** ctx, const int* n, const double* L, const unsigned char* refine, const int* degree,
                           long nleaf, double sigma, int dirichlet, int device);

/* The same with a per-element degree map of the LOCAL brick (x-fastest element order): the hp mesh partitioned element-wise over the
 * ranks.  The degrees of the elements across every rank-boundary face are exchanged once (parallel/updatedegrees.hh:11-46); per
 * apply the rank-boundary face traces -- variable-size blocks per element, cf. parallel/communicationhpdg.hh:309-326,387-418 --
 * travel by one grouped ncclSend/ncclRecv.  nccl_id must not be NULL when nranks > 1. */
int hpdg_create_distributed_hp(hpdg_ctx** out, int dim, const int* n, const double* L, const int* degree, double sigma,
                               int dirichlet, int device, const int* pgrid, int rank, int nranks, const void* nccl_id);
int hpdg_nccl_unique_id(void* out128);
/* Optional NVLink peer-memory halo (one process per GPU, same node): every rank exports the 64-byte cudaIpcMemHandle_t of its
 * halo arena, the caller all-gathers them (indexed by rank) and every rank attaches.  After that the operator apply stores the
 * face traces straight into the neighbours' memory from the pack kernel and the tile kernel waits on per-face step flags: no
 * NCCL call in the apply loop.  NCCL stays in use for the dot-product all-reduce. */
int hpdg_halo_ipc_handle(hpdg_ctx* ctx, void* out64);
int hpdg_halo_ipc_attach(hpdg_ctx* ctx, const void* handles_by_rank);
void hpdg_destroy(hpdg_ctx* ctx);
const char* hpdg_last_error(const hpdg_ctx* ctx); /* ctx may be NULL after a failed create */
int hpdg_set_option(hpdg_ctx* ctx, const char* name, long value); /* "force_generic", "halo_p2p", "halo_timeout_ms", "variant" (40: no persistent kernels), "q3p_grid" */

/* -- sizes (DynamicBlockVector::dimension(), blockRows(i): dynamicbvector.hh:134-143,282) ---------- */
int hpdg_num_levels(const hpdg_ctx* ctx);
long hpdg_num_elements(const hpdg_ctx* ctx);
long hpdg_dimension(const hpdg_ctx* ctx, int level);
int hpdg_block_offsets(const hpdg_ctx* ctx, int level, long* offsets /* nelem+1 */);
int hpdg_level_degrees(const hpdg_ctx* ctx, int level, int* degree /* nelem */);

/* p-multigrid hierarchy as MultigridSetup::setupData builds it (iterationsteps/solversetup.hh:71-108):
 * pLevels = floor(log2(p_max)) coarse levels with order caps p_max/(2*(pLevels-idx)); coarse operators are the
 * Galerkin products (ordertransfer.hh:124-144), i.e. the same form with the FINE face penalties. */
int hpdg_build_p_hierarchy(hpdg_ctx* ctx);

/* -- device vectors ------------------------------------------------------------------------------- */
int hpdg_vec_alloc(hpdg_ctx* ctx, int level, double** d_vec);
int hpdg_vec_free(hpdg_ctx* ctx, double* d_vec);
int hpdg_vec_upload(hpdg_ctx* ctx, int level, const double* h_src, double* d_dst);
int hpdg_vec_download(hpdg_ctx* ctx, int level, const double* d_src, double* h_dst);
int hpdg_host_alloc(hpdg_ctx* ctx, size_t bytes, void** h_ptr); /* pinned */
int hpdg_host_free(hpdg_ctx* ctx, void* h_ptr);
int hpdg_sync(hpdg_ctx* ctx);
void* hpdg_stream(hpdg_ctx* ctx); /* cudaStream_t the context launches on */

/* -- operator: y = factor * A x ---------------------------------------------------------------------
 * Replaces Operator<V,GV,IPDGOperator>::apply(x, Ax) (matrix-free/operator.hh:41-56 with the local operator's
 * factor, ipdgoperator.hh:62-67) and the assembled matrix.mv(x, y) behind operatorFromMatrix
 * (iterationsteps/mg/multigrid.hh:137-155).  y is overwritten. */
int hpdg_op_apply(hpdg_ctx* ctx, int level, const double* h_x, double* h_y, double factor);
int hpdg_op_apply_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor);
int hpdg_op_apply_async(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor);

/* Accumulate mode: y += factor * A x.  Operator::apply over a TUPLE of local operators zeroes Ax once and every local operator
 * adds factor_k * (A_k x) (matrix-free/operator.hh:42-55; matrix-free/test/testoperator.cc:80-98: factors 1 and 2 give 3 A x):
 * the first operator of a tuple maps to hpdg_op_apply*, each further one to hpdg_op_apply_accum*. */
int hpdg_op_apply_accum(hpdg_ctx* ctx, int level, const double* h_x, double* h_y, double factor);
int hpdg_op_apply_accum_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor);
int hpdg_op_apply_accum_async(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor);

/* -- block Jacobi: c = damping * sum_e P_e^T D_e^-1 P_e r -------------------------------------------
 * Replaces IPDGBlockJacobi inside Operator::apply (matrix-free/localoperators/ipdgblockjacobi.hh:58-178) used
 * as Smoother<V>(c, r) (iterationsteps/mg/multigrid.hh:13-14); damping is the caller's `c *= 0.75`
 * (matrix-free/test/testdgblockjacobi.cc:103). */
int hpdg_jacobi_setup(hpdg_ctx* ctx, int level, int form);
int hpdg_jacobi_apply(hpdg_ctx* ctx, int level, int form, const double* h_r, double* h_c, double damping);
int hpdg_jacobi_apply_device(hpdg_ctx* ctx, int level, int form, const double* d_r, double* d_c, double damping);
/* enqueue only (no synchronisation): pair with hpdg_sync, like hpdg_op_apply_async */
int hpdg_jacobi_apply_async(hpdg_ctx* ctx, int level, int form, const double* d_r, double* d_c, double damping);
size_t hpdg_jacobi_bytes(const hpdg_ctx* ctx, int level, int form);
/* MatrixCreator protocol bind(e) + matrix() (matrix-free/localoperators/ipdgdiagonalblock.hh:29-360,
 * slowipdgdiag.hh:32-218): the diagonal block A_ee, n_e x n_e row-major, to host memory. */
int hpdg_diag_block(hpdg_ctx* ctx, int level, long element, double* h_out);

/* -- assembled matrix in the reference's DynamicBCRSMatrix layout (common/dynamicbcrs.hh:178-199: one contiguous array,
 * blocks in (block row, ascending block column) order, each dense row-major rowMap[i] x colMap[j]) -- replaces
 * BuildingBlocks::laplace / dynamicStiffnessMatrix (buildingblocks/matrices.hh:29-89, test/testobjects.hh:20-81).
 * Intended for small/medium meshes (3-D Q3 costs 229 KB per element).  hpdg_bcrs_sizes gives the array lengths;
 * hpdg_assemble_bcrs builds the matrix on the device (it stays resident for hpdg_bcrs_mv / hpdg_blockgs_iterate) and copies
 * pattern + entries to the host arrays (any of them may be NULL). */
int hpdg_bcrs_sizes(hpdg_ctx* ctx, int level, long* nblocks, long* nentries);
int hpdg_assemble_bcrs(hpdg_ctx* ctx, int level, long* h_block_row_ptr /* nelem+1 */, int* h_block_col /* nblocks */,
                       long* h_block_off /* nblocks+1 */, double* h_entries /* nentries */);
/* y = A x with the assembled matrix: BCRSMatrix<MatrixWindow>::mv (common/matrixwindow.hh:196-209) */
int hpdg_bcrs_mv(hpdg_ctx* ctx, int level, const double* h_x, double* h_y);
int hpdg_bcrs_mv_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y);
/* one DynamicBlockGS::iterate() with the GSCore local solver (iterationsteps/dynamicblockgs.hh:17-40,94-126): block rows in
 * ascending order, inner forward scalar GS sweep from zero.  x is updated in place. */
int hpdg_blockgs_iterate(hpdg_ctx* ctx, int level, const double* h_b, double* h_x);
int hpdg_blockgs_iterate_device(hpdg_ctx* ctx, int level, const double* d_b, double* d_x);
/* the same DynamicBlockGS::iterate() WITHOUT an assembled matrix: per hyperplane of independent block rows, (A x)_i comes from
 * the operator kernel's element pass (own block + the neighbours' current face traces) and the diagonal block A_ii from its
 * Kronecker factors, so the reference's default smoother (solversetup.hh:139-145) runs at any mesh size the vectors fit.
 * No hpdg_assemble_bcrs needed; single rank. */
int hpdg_blockgs_mf_iterate(hpdg_ctx* ctx, int level, const double* h_b, double* h_x);
int hpdg_blockgs_mf_iterate_device(hpdg_ctx* ctx, int level, const double* d_b, double* d_x);
/* L1Smoother (iterationsteps/l1smoother.hh:20-145), the reference's smoother for MPI runs: block Gauss-Seidel whose local solver
 * (a forward scalar GS sweep from zero, :127-145) divides by D_aa + reg_a, where reg is the l1 norm of the rows of the blocks
 * that couple an owned block row to a ghost block (preprocess(), :31-57).  hpdg_l1_setup = L1Smoother(ghosts) + preprocess():
 * `ghosts` are block (element) indices, duplicates count again as in the reference; hpdg_l1_iterate = iterate() (:63-113),
 * x updated in place.  Needs hpdg_assemble_bcrs first. */
int hpdg_l1_setup(hpdg_ctx* ctx, int level, const long* ghosts, long nghost);
int hpdg_l1_iterate(hpdg_ctx* ctx, int level, const double* h_b, double* h_x);
int hpdg_l1_iterate_device(hpdg_ctx* ctx, int level, const double* d_b, double* d_x);

/* -- p-transfer between level and level-1 (transferoperators/ordertransfer.hh:91-119) ------------- */
int hpdg_restrict(hpdg_ctx* ctx, int fine_level, const double* h_fine, double* h_coarse);
int hpdg_prolong(hpdg_ctx* ctx, int fine_level, const double* h_coarse, double* h_fine);
int hpdg_restrict_device(hpdg_ctx* ctx, int fine_level, const double* d_fine, double* d_coarse);
int hpdg_prolong_device(hpdg_ctx* ctx, int fine_level, const double* d_coarse, double* d_fine);

/* -- one multigrid cycle: Multigrid<Vector>::apply(x, b) (iterationsteps/mg/multigrid_impl.hh:16-117) with the
 * block-Jacobi smoother on every level and `coarse_its` damped Jacobi iterations as coarse solver, or -- with form =
 * HPDG_SMOOTHER_BLOCKGS -- the reference's own configuration: DynamicBlockGS on the assembled Galerkin level matrices as
 * smoother and `coarse_its` block-GS iterations as coarse solver (solversetup.hh:139-145,198-215; small/medium meshes), or --
 * form = HPDG_SMOOTHER_BLOCKGS_MF -- the same cycle with the matrix-free block-GS sweeps (no matrices).  On return x += correction and
 * b holds the residual (multigrid_impl.hh:60-61). */
int hpdg_vcycle(hpdg_ctx* ctx, int form, double damping, int pre, int post, int coarse_its, double* h_x, double* h_b);
int hpdg_vcycle_device(hpdg_ctx* ctx, int form, double damping, int pre, int post, int coarse_its, double* d_x,
                       double* d_b);

/* -- BLAS-1 used by the Krylov / MG drivers (DynamicBlockVector::operator*, two_norm:
 * common/dynamicbvector.hh:258-264,300-314); sums over all ranks of a distributed context. */
int hpdg_dot_device(hpdg_ctx* ctx, int level, const double* d_x, const double* d_y, double* h_result);
int hpdg_two_norm_device(hpdg_ctx* ctx, int level, const double* d_x, double* h_result);
int hpdg_axpy_device(hpdg_ctx* ctx, int level, double a, const double* d_x, double* d_y);   /* y += a x  (enqueue only) */
int hpdg_scale_device(hpdg_ctx* ctx, int level, double a, double* d_x);                     /* x *= a    (enqueue only) */
int hpdg_assign_device(hpdg_ctx* ctx, int level, const double* d_src, double* d_dst);       /* dst = src (enqueue only) */

/* -- solver loops around the hot path (finest level) ------------------------------------------------------------------------
 * hpdg_pcg: preconditioned conjugate gradients, everything resident on the device: per iteration one operator apply, one
 * preconditioner application (HPDG_PRECOND_*: none / fd block Jacobi with `damping` / one V-cycle with fd block-Jacobi smoothing,
 * pre = post = `smooth`, `coarse_its` coarse iterations), three dot products that each end in a 1-double ncclAllReduce on the
 * context stream of a distributed context, and two fused updates whose step lengths are read from device memory.  The host only
 * reads the residual norm every `check_every` iterations.  Stops at ||r|| <= tol ||r_0|| or after maxit iterations; x is the
 * initial iterate on entry.  iters / relres may be NULL.
 * hpdg_loop_solve: Dune::Solvers::LoopSolver around the multigrid step with the energy norm, as buildingblocks/solve.hh:150-166
 * wires it (MultigridWrapper::iterate copies the right-hand side, iterationsteps/mg/mgwrapper.hh:22-27): stops when
 * ||x_k - x_{k-1}||_A / ||x_{k-1}||_A < tol or after maxit iterations. */
int hpdg_pcg(hpdg_ctx* ctx, int precond, double damping, int smooth, int coarse_its, double* h_x, const double* h_b, double tol,
             int maxit, int check_every, int* iters, double* relres);
int hpdg_pcg_device(hpdg_ctx* ctx, int precond, double damping, int smooth, int coarse_its, double* d_x, const double* d_b,
                    double tol, int maxit, int check_every, int* iters, double* relres);
int hpdg_loop_solve_device(hpdg_ctx* ctx, int form, double damping, int pre, int post, int coarse_its, double* d_x,
                           const double* d_b, double tol, int maxit, int* iters, double* last_error);

/* -- introspection --------------------------------------------------------------------------------- */
/* Host-only (no device work): the 1-D tables of degree p the kernels are built from -- GL nodes on [0,1] ascending
 * (qkgllocalbasis.hh:222-234), exact mass int l_i l_j and stiffness int l_i' l_j' (n x n row-major, n = p + 1), l_i(s) and l_i'(s)
 * at the end points s = 0, 1 ([2][n]).  Any output may be NULL. */
int hpdg_tables_1d(int degree, double* nodes, double* mass, double* stiffness, double* end_values, double* end_derivatives);
/* Host-only: the tangential coupling (M^ee)^-1 int_I l^e_i(tau) l^o_j(tau_o(tau)) dtau of a face between elements of degree pe and po.
 * kind 0: conforming face (I = the whole side: the L2 projection of the neighbour's trace, variableipdg.hh:326-361); 1 / 2: e is the
 * coarse side of a hanging face, I = low / high half of its side; 3 / 4: e is the fine side on the low / high half of the neighbour's
 * side (sfipdg.hh:472-491).  out: (pe+1) x (po+1) row-major.  own (kinds 1, 2; may be NULL): (pe+1) x (pe+1), the same with l^e_j. */
int hpdg_tables_face(int pe, int po, int kind, double* out, double* own);
long hpdg_launch_count(const hpdg_ctx* ctx);      /* kernels launched so far by this context */
int hpdg_uses_uniform_kernel(const hpdg_ctx* ctx, int level);
/* time `reps` back-to-back operator applies with CUDA events on the context stream; ms per apply */
int hpdg_time_apply_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, int reps, float* ms_per_apply);
/* the same for `reps` back-to-back block-Jacobi applications c = damping * D^-1 r (hpdg_jacobi_setup first) */
int hpdg_time_jacobi_device(hpdg_ctx* ctx, int level, int form, const double* d_r, double* d_c, double damping, int reps,
                            float* ms_per_apply);

#ifdef __cplusplus
}
#endif
#endif
