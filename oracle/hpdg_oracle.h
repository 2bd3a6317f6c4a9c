/*
 * hpdg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference algorithms on the north-star path of
 * c1887/dune-hpdg: the SIPG Poisson operator on the hp Qk Gauss-Lobatto DG basis
 * (matrix-free quadrature loop and assembled DynamicBCRSMatrix), the block
 * Gauss-Seidel / block-Jacobi smoothers, the p-transfer and the multigrid cycle.
 * Each function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (CUDA) path never links or calls it.
 *
 * PARITY PIN STATUS: the reference cannot be compiled here (ten un-vendored DUNE
 * modules) and ships no golden vectors.  The oracle is pinned by re-running the
 * reference's own DIFFERENTIAL tests (two independent formulations restated here:
 * quadrature-loop matrix-free vs assembled matrix; see tests/test_oracle_pins.py) and
 * by closed-form properties.  3-D has no reference test at all: "parity unpinned by
 * reference tests" for 3-D; it is pinned by the dim-generic source only.
 */
#ifndef HPDG_ORACLE_H
#define HPDG_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct omesh omesh;   /* structured YaspGrid-like mesh + per-element degrees */
typedef struct obcrs obcrs;   /* DynamicBCRSMatrix-layout block CSR */

/* --- 1-D building blocks ------------------------------------------------ */
void   orc_gl_nodes(int p, double* x);                 /* p+1 GL nodes on [0,1], ascending */
void   orc_gauss_legendre(int m, double* x, double* w);/* m-point Gauss rule on [0,1] */
double orc_lagrange(int p, const double* nodes, int i, double x);
double orc_lagrange_prime(int p, const double* nodes, int i, double x);

/* --- mesh / basis -------------------------------------------------------- */
omesh* orc_mesh_create(int dim, const int* n, const double* L, const int* degree,
                       const int* pen_degree /* NULL => degree */, double sigma, int dirichlet);
void   orc_mesh_destroy(omesh* m);
long   orc_mesh_dimension(const omesh* m);
long   orc_mesh_nelem(const omesh* m);
void   orc_mesh_offsets(const omesh* m, long* off /* nelem+1 */);
/* interpolate f(x)=|x|^2 (testdg.cc:97) into the nodal basis */
void   orc_interpolate_normsq(const omesh* m, double* x);

/* --- matrix-free apply (operator.hh:41-56 + ipdgoperator.hh:80-390) ----- */
void   orc_apply_mf(const omesh* m, const double* x, double* y, double factor, int threads);

/* --- assembled matrix (matrices.hh:29-89 pattern, variableipdg.hh, LaplaceAssembler) */
obcrs* orc_assemble(const omesh* m);
void   orc_bcrs_destroy(obcrs* A);
long   orc_bcrs_nblocks(const obcrs* A);
long   orc_bcrs_nentries(const obcrs* A);
void   orc_bcrs_export(const obcrs* A, long* rowptr, int* col, long* boff, double* val);
void   orc_bcrs_mv(const obcrs* A, const double* x, double* y, int threads);      /* y = A x  (matrixwindow.hh:196) */
void   orc_bcrs_mmv(const obcrs* A, const double* x, double* y);                  /* y -= A x */
double orc_bcrs_frobenius_diff(const obcrs* A, const obcrs* B);
/* diagonal block e (n_e x n_e row-major) */
void   orc_bcrs_diag_block(const obcrs* A, long e, double* out);

/* --- smoothers ------------------------------------------------------------ */
/* one DynamicBlockGS::iterate with GSCore inner solver (dynamicblockgs.hh:17-40,94-126) */
void   orc_blockgs_iterate(const obcrs* A, const double* b, double* x);
/* L1Smoother (iterationsteps/l1smoother.hh:20-145): preprocess -> reg (one double per DoF), then iterate */
void   orc_l1_regularization(const obcrs* A, const long* ghosts, long nghost, double* reg);
void   orc_l1_iterate(const obcrs* A, const double* reg, const double* b, double* x);
/* matrix-free block Jacobi c = sum_e P_e^T solve(D_e, P_e r)  (ipdgblockjacobi.hh:58-178)
 * local_solver: 0 = exact (dense Cholesky), 1 = one scalar GS sweep from zero (testdgblockjacobi.cc:63-76) */
void   orc_blockjacobi_apply(const omesh* m, const double* r, double* c, double factor, int local_solver);
/* the diagonal block as ipdgblockjacobi.hh assembles it (bulk + own-side face terms) */
void   orc_diag_block_mf(const omesh* m, long e, double* out);

/* --- p-transfer (dynamicordertransfer.hh:48-73, ordertransfer.hh:45-144) - */
void   orc_transfer_matrix(int dim, int p_coarse, int p_fine, double* T /* nf^d x nc^d row-major */);
/* coarse mesh = same mesh, degrees min(p, max_order), penalty degrees inherited from m */
omesh* orc_mesh_coarsen(const omesh* m, int max_order);
void   orc_restrict(const omesh* fine, const omesh* coarse, const double* xf, double* xc);
void   orc_prolong(const omesh* fine, const omesh* coarse, const double* xc, double* xf);
obcrs* orc_galerkin_restrict(const omesh* fine, const omesh* coarse, const obcrs* Af);

/* --- multigrid V-cycle (mg/multigrid_impl.hh:16-117) ---------------------- */
/* levels[0] = coarsest ... levels[nlev-1] = finest; smoother: 0 = DynamicBlockGS on the
 * assembled level matrices (reference default, solversetup.hh:139-145), 1 = damped exact
 * block Jacobi (matrix-free level operators).  x += correction, b := residual (as :60-61). */
void   orc_vcycle(int nlev, omesh* const* levels, obcrs* const* mats /* NULL for smoother 1 */,
                  int smoother, double damping, int pre, int post, int coarse_its,
                  double* x, double* b);

/* --- helpers --------------------------------------------------------------- */
/* libstdc++-compatible fill: mt19937(seed) + normal_distribution<>(0,1) (test/randomvector.hh:11-21) */
void   orc_fill_random(double* v, long n, unsigned seed);
int    orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
