// Internal context of the B200 SIPG hot path (host structures + device pointers).
// The public surface is the C ABI in include/hpdg_b200.h; nothing here is exported.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "tables.hpp"

namespace hpdg {

// ---- ghost layer of a rank-local brick (multi-GPU) -------------------------------
// For each of the 2*dim brick faces that touches another rank, the neighbour rank's boundary
// elements contribute, per face node, the pair (der, val) = (g_{1-s}.u_line, t_{1-s}.u_line) of
// their DoF lines normal to the face: "face traces".  Element degree is uniform across ranks in
// the distributed path (checked at create time).
struct Ghost {
  bool active[6] = {false, false, false, false, false, false};
  double* d_recv[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [nfaceelem][N^(dim-1)][2]
  double* d_send[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t count[6] = {0, 0, 0, 0, 0, 0};  // doubles per face buffer
  int peer[6] = {-1, -1, -1, -1, -1, -1};
  // peer-to-peer mode: the pack kernel stores the traces straight into the neighbour's arena over NVLink and
  // publishes a per-face step flag; the tile kernel's rank-boundary tiles wait on the local flags.
  bool p2p = false;
  bool peer_attached = false;
  char* arena = nullptr;            // local: recv[f][parity] buffers, then int flags[6][2], then int err
  size_t arena_bytes = 0;
  size_t recv_off[6][2] = {};       // byte offsets into the arena (identical on every rank: same brick shape)
  size_t flag_off = 0;
  char* peer_arena[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int step = 0;
};

// Distributed hp: ghost layer of one level.  The degrees of the remote elements across every rank-boundary face are exchanged once
// (the analogue of parallel/updatedegrees.hh:11-46); their face traces -- variable-size blocks per element, like the reference's
// ghost copy (parallel/communicationhpdg.hh:309-326,387-418), but N^(dim-1) (der, val) pairs instead of whole blocks -- per apply.
struct HpGhost {
  bool ready = false;
  int maxp = 0;                  // largest ghost degree
  long nface[6] = {};            // face elements of brick face f
  int* d_deg[6] = {};            // [nface] degree of the remote element on this level
  int* d_pdeg[6] = {};           // [nface] its degree on the finest level (enters the face penalty on Galerkin coarse levels)
  long* d_troff[6] = {};         // [nface] offset (pairs) of the remote element's traces in d_recv[f]
  double* d_recv[6] = {};        // received traces (pairs)
  double* d_send[6] = {};        // packed traces of this rank's boundary elements on face f
  long* d_send_src[6] = {};      // [nface] offset (pairs) of the element's face-f traces in the level's trace array
  long* d_send_dst[6] = {};      // [nface + 1] offsets (pairs) in d_send[f]
  long send_pairs[6] = {}, recv_pairs[6] = {};
  std::vector<int> h_deg[6], h_pdeg[6];   // host copies of the ghost degrees (block-Jacobi setup)
  std::vector<long> h_troff[6];           // host copy of d_troff
};

// Per-face metadata of an element in the generic hp kernel (apply_generic.cu); built once per level on the host.
struct FaceInfo {
  double nuk;    // nu * kappa_d: outward normal sign of the face times prod_{d' != d} h_d' / h_d
  double cpen;   // sigma * max(p-,p+)^2 (ipdgoperator.hh:129-131) or sigma p^2 on a Dirichlet face (:310)
  long tro;      // offset (pairs) of the neighbour's traces on the shared face
  short mode;    // -1 unused slot; 0 natural boundary: no face term (ipdgoperator.hh:97-105); 1 Dirichlet boundary (weight 1,
                 // :357); 2 neighbour of the same degree; 3 neighbour of another degree (tangential L2 projection)
  short po;      // neighbour degree
  short ghost;   // neighbour lives on another rank: traces from the ghost layer of brick face f
  short kind;    // non-conforming meshes: 0 conforming intersection; 1 / 2 this element is the coarse side and the
                 // intersection covers the low / high half of its side; 3 / 4 it is the fine side on the low / high half of the
                 // neighbour's side (sfipdg.hh:472-491)
};
static_assert(sizeof(FaceInfo) == 32, "FaceInfo layout");

struct JacobiDense {
  bool ready = false;
  double* d_inv = nullptr;      // all inverses, bucket by bucket, element-major inside a bucket
  std::vector<size_t> bucket_off;  // start (in doubles) of each degree bucket's inverses
  size_t bytes = 0;
};

struct JacobiFD {
  bool ready = false;
  // Fast-diagonalisation form of D_e^-1 = (Vx x Vy x Vz) diag(1/(lx+ly+lz)) (Vx x Vy x Vz)^T.
  // 1-D factors are deduplicated: each entry is {V (kMaxN*kMaxN, row-major n x n), lambda (kMaxN)}.
  double* d_fac = nullptr;      // [nfac][kMaxN*kMaxN + kMaxN]
  int* d_idx = nullptr;         // [nelem][3] factor index per direction
  int nfac = 0;
};

// Assembled matrix in DynamicBCRSMatrix layout (common/dynamicbcrs.hh:178-199) + hyperplane lists for block-GS
struct Bcrs {
  bool ready = false;
  std::vector<long> rowptr, boff, wave_begin;
  std::vector<int> col;
  long* d_rowptr = nullptr; int* d_col = nullptr; int* d_brow = nullptr; long* d_boff = nullptr;
  double* d_val = nullptr; int* d_wave = nullptr; double* d_res = nullptr;
  double* d_l1reg = nullptr; bool l1_ready = false;  // L1Smoother's diagonal regularisation (iterationsteps/l1smoother.hh:31-57)
};

struct Level {
  int dim = 0;
  int n[3] = {1, 1, 1};
  double h[3] = {1, 1, 1};
  long nelem = 0, ndof = 0;
  std::vector<int> deg, pdeg;
  std::vector<long> off;
  int* d_deg = nullptr;
  int* d_pdeg = nullptr;
  long* d_off = nullptr;
  bool uniform = false;   // all deg equal and all pdeg equal
  int p_uni = -1, pen_uni = -1;
  // degree buckets: elements sorted by degree (stable), for launches with a uniform block size
  std::vector<int> bucket_p;
  std::vector<long> bucket_begin;  // size buckets+1
  int* d_elist = nullptr;
  int maxp = 0;
  int cap = -1;   // p-hierarchy: degree cap of this level (-1: finest level, no cap)
  HpGhost hpg;    // distributed hp: ghost degrees / trace buffers of this level
  // hp apply: face traces (der, val) of every element, written once per apply by k_face_traces (apply_generic.cu)
  long* d_troff = nullptr;      // [nelem+1] offset (in pairs) of an element's traces; face f at troff[e] + f * N_e^(dim-1)
  std::vector<long> troff_h;    // host copy
  FaceInfo* d_finfo = nullptr;  // [nelem][2 dim][fslots] face metadata of the generic kernel (built on first use)
  int fslots = 1;               // intersections per element side: 1, or 2 on a non-conforming mesh
  // non-conforming 2-D mesh (one level of refinement, hpdg_create_refined_2d): leaf elements in base-cell order, children x-fastest
  bool nc = false;
  std::vector<FaceInfo> nc_faces;   // host: [nelem][4][2], tro still relative (filled in by generic_face_table)
  std::vector<long> nc_nbr;         // host: [nelem][4][2] neighbour element of the intersection (-1: none)
  double* d_tr = nullptr;       // [2 * tr_pairs]
  long tr_pairs = 0;
  // matrix-free block Gauss-Seidel: elements sorted by (hyperplane ix+iy+iz, degree bucket); segment k = w * buckets + b
  int* d_gs_elist = nullptr; std::vector<long> gs_seg; int gs_nw = 0;
  JacobiDense jd;
  JacobiFD jf;
  Bcrs bcrs;
  Ghost cg;   // halo buffers of a coarse level (NCCL transport); the finest level uses Ctx::ghost
  // compact tile lists (interior / rank-boundary tiles) of the distributed apply, per tile shape
  int *d_tiles_int = nullptr, *d_tiles_bnd = nullptr, *d_tiles_all = nullptr;  // all = interior first, boundary last
  long n_tiles_int = 0, n_tiles_bnd = 0, tile_key = -1;
  double* d_jinv = nullptr;     // persistent Q3 block Jacobi: damping / eigenvalue sums, [vx][vy][vz][j][k][i]
  double jinv_damping = 0;      // the damping d_jinv was built with (0: not built)
  int q3j_state = 0;            // persistent Q3 block Jacobi: 0 = not set up, 1 = usable, -1 = factor lacks the mirror structure
  double q3j_V[3][3][16], q3j_lam[3][3][4];  // its 1-D eigenvectors / eigenvalues, [direction][boundary variant]
  double q4j_V[3][3][25], q4j_lam[3][3][5];  // the same for the experimental persistent Q4 kernel (variant 50)
  void* d_tile_desc = nullptr;  // persistent Q3 kernel: int4 per 4x4x4 tile {first element, packed tile coordinates, brick-face bits, 0}
  // scratch vectors for the V-cycle (device, ndof each)
  double *mg_x = nullptr, *mg_r = nullptr, *mg_t1 = nullptr, *mg_t2 = nullptr;
};

constexpr int kBucketStreams = 8;

struct Ctx {
  int device = 0;
  int dim = 0;
  double sigma = 2.0;
  int dirichlet = 1;
  cudaStream_t stream = nullptr, stream_comm = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  DegTable* d_tab = nullptr;
  double* d_P = nullptr;
  double* d_T = nullptr;
  double* d_Mab = nullptr;
  double *d_Pnc_eo = nullptr, *d_Pnc_ee = nullptr;   // non-conforming face couplings (tables.hpp), uploaded on first use
  std::vector<Level> levels;  // [0] coarsest ... back() finest (reference: multigrid_impl.hh:19-20)
  std::string err;
  // distributed brick
  int rank = 0, nranks = 1;
  bool create_nc = false;        // set by hpdg_create_refined_2d before the first level is set up
  bool hp_distributed = false;   // created with a per-element degree array: every level takes the generic hp path
  int pgrid[3] = {1, 1, 1}, pcoord[3] = {0, 0, 0};
  bool bnd_is_rank[6] = {false, false, false, false, false, false};
  Ghost ghost;           // finest level only
  void* nccl = nullptr;  // ncclComm_t
  // device staging for the host-pointer entry points
  double *d_in = nullptr, *d_out = nullptr;
  size_t stage_cap = 0;
  int force_generic = 0;
  int variant = 0;  // kernel variant selector for tuning experiments
  int q3p_grid = 0; // persistent Q3 kernel: CTA count override (0 = one per SM slot)
  int* d_sched = nullptr;  // persistent Q3 kernel: tile counters of its dynamic scheduler
  int q3p_tune = 0; // persistent Q3 kernel: tuning switches (bit 0: L2 prefetch two tiles ahead)
  int slab_z0 = 0, slab_nz = 0;  // restrict the next uniform launch to element layers [z0, z0+nz) (chunked host apply)
  cudaStream_t stream_h2d = nullptr, stream_d2h = nullptr;
  cudaStream_t bucket_stream[kBucketStreams] = {};  // hp apply: degree buckets run concurrently (one side stream per bucket, up to 8)
  cudaEvent_t bucket_ev[kBucketStreams + 1] = {};
  cudaEvent_t ev_chunk[3][32] = {};
  // fusion requests of the V-cycle driver, consumed by the next launch:
  int fuse_accum = 0;          // apply: y = y_old + factor * A x   (r -= A c without a separate axpy)
  double* fuse_xacc = nullptr; // block Jacobi: additionally x += c
  double* fuse_xin = nullptr;  // uniform operator apply: additionally fuse_xin += (input vector); else an axpy in front of it
  int xacc_in_apply = 1;       // V-cycle: fuse x += c into the apply of c (uniform levels) instead of into the Jacobi kernel (option)
  std::map<const void*, int> kattr;  // kernel_slots(): kernels whose attributes are set on this context's device -> resident CTA slots
  double* d_scalar = nullptr;        // device scratch of the BLAS-1 reductions (one context = one device)
  double* d_partial = nullptr;       // per-CTA partial sums of the two-stage dot product
  volatile int* h_ghost_err = nullptr;  // p2p halo: time-out flag of the tile kernels (mapped pinned host memory)
  int* d_ghost_err = nullptr;          // its device alias
  long long halo_timeout_cycles = 40000000000LL;  // ~20 s of SM clock; option "halo_timeout_ms"
  double *cg_p = nullptr, *cg_q = nullptr, *cg_r = nullptr, *cg_z = nullptr;  // work vectors of the solver loops (finest level)
  bool end_tables_set = false;  // apply_generic.cu: end-point tables of all degrees copied to this device's constant memory
  long launches = 0;  // kernels launched by this context (bench.py's gpu_launches)
};

#define HPDG_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

// ---- kernel launchers (defined in the .cu files) -----------------------------------------------
int launch_apply_generic(Ctx* ctx, Level& L, const double* x, double* y, double factor);
int launch_face_traces(Ctx* ctx, Level& L, const double* x, cudaStream_t stream);  // every element's own face traces -> L.d_tr
int launch_hp_pack(Ctx* ctx, Level& L, cudaStream_t stream);   // rank-boundary traces of L.d_tr -> L.hpg.d_send[f]
int generic_trace_setup(Ctx* ctx, Level& L);                   // allocates L.d_troff / L.d_tr on first use
int generic_face_table(Ctx* ctx, Level& L);                    // builds L.d_finfo on first use (after the ghost setup of a distributed level)
int hp_ghost_setup(Ctx* ctx, Level& L);                        // (api.cu) one-time exchange of the neighbour degrees across rank boundaries
int hp_halo_exchange(Ctx* ctx, Level& L);                      // (api.cu) pack + NCCL send/recv of the rank-boundary face traces
// returns -1 if (dim, degree) has no specialised kernel
int launch_apply_uniform(Ctx* ctx, Level& L, const double* x, double* y, double factor, int part, cudaStream_t stream = nullptr);
int uniform_supported(const Ctx* ctx, const Level& L);
int uniform_persistent(const Ctx* ctx, const Level& L, const double* x);  // the level's apply of x runs the persistent Q3 tile kernel
// one-time, per-context (= per-device) kernel attributes: opts the kernel in to `smem` bytes of dynamic shared memory and, if
// slots != nullptr, returns the number of CTAs of `threads` threads that are resident on the whole device at once
int kernel_slots(Ctx* ctx, const void* func, int threads, size_t smem, int* slots);
int q3p_level_setup(Ctx* ctx, Level& L, int tile_h = 4);  // descriptors of the 4 x 4 x tile_h tiles + scheduler counters of the persistent kernels
int uniform_tile_height(const Level& L);
int uniform_tile_lists(Ctx* ctx, Level& L, int TX, int TY, int TZ, const int* bmode);
int launch_pack_traces(Ctx* ctx, Level& L, const double* x, cudaStream_t stream);
int level_ghost(Ctx* ctx, Level& L, Ghost** out);  // halo buffers of this level (allocated on first use for coarse levels)
int launch_halo_flags(Ctx* ctx, cudaStream_t stream);

int jacobi_setup_dense(Ctx* ctx, Level& L);
int jacobi_apply_dense(Ctx* ctx, Level& L, const double* r, double* c, double damping);
int jacobi_setup_fd(Ctx* ctx, Level& L);
int jacobi_apply_fd(Ctx* ctx, Level& L, const double* r, double* c, double damping);
int jacobi_apply_fd_uniform(Ctx* ctx, Level& L, const double* r, double* c, double damping);  // -1: no specialised kernel
int diag_block_device(Ctx* ctx, Level& L, long e, double* d_out);

int bcrs_build(Ctx* ctx, Level& L);
int bcrs_mv(Ctx* ctx, Level& L, const double* x, double* y);
int blockgs_iterate(Ctx* ctx, Level& L, const double* b, double* x, int l1 = 0);
int blockgs_mf_iterate(Ctx* ctx, Level& L, const double* b, double* x);   // (apply_generic.cu) the same sweep without a matrix
int l1_setup(Ctx* ctx, Level& L, const long* ghosts, long nghost);

int launch_restrict(Ctx* ctx, Level& fine, Level& coarse, const double* xf, double* xc);
int launch_prolong(Ctx* ctx, Level& fine, Level& coarse, const double* xc, double* xf);
int launch_axpy(Ctx* ctx, long n, double a, const double* x, double* y);          // y += a x
int launch_xpay_sub(Ctx* ctx, long n, const double* b, const double* ax, double* r);  // r = b - ax
int launch_dot(Ctx* ctx, long n, const double* x, const double* y, double* d_result);    // *d_result = x . y (this rank's part)
int launch_scale(Ctx* ctx, long n, double a, double* x);                                  // x *= a
constexpr int kScalarSlots = 16;   // Ctx::d_scalar: device-resident scalars of the BLAS-1 / Krylov drivers
int blas_scratch(Ctx* ctx);        // allocates Ctx::d_scalar / d_partial on first use
// preconditioned CG updates with device-resident scalars s = Ctx::d_scalar (no host round trip inside an iteration)
int launch_cg_update(Ctx* ctx, long n, int num, int den, const double* p, const double* q, double* x, double* r);  // a = s[num]/s[den]; x += a p; r -= a q
int launch_cg_update_rr(Ctx* ctx, long n, int num, int den, const double* p, const double* q, double* x, double* r, double* d_rr);  // ... and *d_rr = r . r
int launch_cg_direction(Ctx* ctx, long n, int num, int den, const double* z, double* p);                          // b = s[num]/s[den]; p = z + b p

}  // namespace hpdg
