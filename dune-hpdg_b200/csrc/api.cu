// C ABI (include/hpdg_b200.h) over the CUDA kernels: context, levels, staging, V-cycle driver, NCCL halo.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/hpdg_b200.h"
#include "ctx.hpp"

using namespace hpdg;

struct hpdg_ctx : public Ctx {};

static std::string g_create_err;

// ---- NCCL through dlopen: the library keeps no link-time dependency on NCCL, and inside a torch
// process it binds to the libnccl torch already loaded. -------------------------------------------
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }
  if (!h) for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { err = "missing NCCL symbol " name; return false; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.lib = h;
  return true;
}
#define HPDG_NCCL(call)                                                                       \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != ncclSuccess) { ctx->err = std::string(#call) + ": " + g_nccl.GetErrorString(r__); return 1; } \
  } while (0)

int ipow_h(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

// Every extern "C" entry point that touches the device starts here: the context's device becomes the calling thread's current
// device (a process may hold contexts on several devices, and the caller's current device is its own business).
#define HPDG_ENTER(ctx) HPDG_CUDA(cudaSetDevice((ctx)->device))

// Synchronise the context's streams and report a halo time-out of the NVLink peer-memory path: a rank-boundary tile that gave up
// waiting for a neighbour's traces has computed from stale data, so EVERY synchronising entry point fails (and clears the flag, so
// that the next exchange starts clean).  The flag lives in mapped pinned host memory: reading it costs nothing.
int sync_check(Ctx* ctx) {
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream));
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream_comm));
  if (ctx->h_ghost_err && *ctx->h_ghost_err) {
    *ctx->h_ghost_err = 0;
    cudaMemset(ctx->ghost.arena + ctx->ghost.flag_off + 12 * sizeof(int), 0, sizeof(int));
    ctx->err = "halo exchange timed out waiting for a neighbour rank's face traces (results of this step are invalid)";
    return 1;
  }
  return 0;
}

int setup_level(Ctx* ctx, Level& L, int dim, const int* n, const double* h, const std::vector<int>& deg,
                const std::vector<int>& pdeg) {
  L.dim = dim;
  L.nelem = 1;
  for (int d = 0; d < 3; d++) { L.n[d] = d < dim ? n[d] : 1; L.h[d] = d < dim ? h[d] : 1.0; L.nelem *= L.n[d]; }
  if (L.nc) L.nelem = (long)deg.size();   // non-conforming mesh: the leaf elements of the refined base grid
  L.deg = deg; L.pdeg = pdeg;
  L.off.assign(L.nelem + 1, 0);
  L.maxp = 0;
  bool uni = true;
  for (long e = 0; e < L.nelem; e++) {
    L.off[e + 1] = L.off[e] + ipow_h(deg[e] + 1, dim);
    L.maxp = std::max(L.maxp, deg[e]);
    if (deg[e] != deg[0] || pdeg[e] != pdeg[0]) uni = false;
  }
  L.ndof = L.off[L.nelem];
  if (L.nc) uni = false;   // only the generic element pass knows hanging faces
  L.uniform = uni; L.p_uni = uni ? deg[0] : -1; L.pen_uni = uni ? pdeg[0] : -1;
  // buckets by degree (stable counting sort)
  std::vector<long> cnt(kMaxP + 2, 0);
  for (long e = 0; e < L.nelem; e++) cnt[deg[e] + 1]++;
  for (int p = 0; p <= kMaxP; p++) cnt[p + 1] += cnt[p];
  std::vector<int> elist(L.nelem);
  std::vector<long> pos(cnt.begin(), cnt.end() - 1);
  for (long e = 0; e < L.nelem; e++) elist[pos[deg[e]]++] = (int)e;
  L.bucket_p.clear(); L.bucket_begin.clear();
  for (int p = 0; p <= kMaxP; p++)
    if (cnt[p + 1] > cnt[p]) { L.bucket_p.push_back(p); L.bucket_begin.push_back(cnt[p]); }
  L.bucket_begin.push_back(L.nelem);
  HPDG_CUDA(cudaMalloc(&L.d_deg, sizeof(int) * L.nelem));
  HPDG_CUDA(cudaMalloc(&L.d_pdeg, sizeof(int) * L.nelem));
  HPDG_CUDA(cudaMalloc(&L.d_off, sizeof(long) * (L.nelem + 1)));
  HPDG_CUDA(cudaMalloc(&L.d_elist, sizeof(int) * L.nelem));
  HPDG_CUDA(cudaMemcpy(L.d_deg, deg.data(), sizeof(int) * L.nelem, cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMemcpy(L.d_pdeg, pdeg.data(), sizeof(int) * L.nelem, cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMemcpy(L.d_off, L.off.data(), sizeof(long) * (L.nelem + 1), cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMemcpy(L.d_elist, elist.data(), sizeof(int) * L.nelem, cudaMemcpyHostToDevice));
  return 0;
}

void free_level(Level& L) {
  cudaFree(L.d_deg); cudaFree(L.d_pdeg); cudaFree(L.d_off); cudaFree(L.d_elist); cudaFree(L.d_troff); cudaFree(L.d_tr); cudaFree(L.d_gs_elist); cudaFree(L.d_finfo);
  cudaFree(L.jd.d_inv); cudaFree(L.jf.d_fac); cudaFree(L.jf.d_idx);
  cudaFree(L.mg_x); cudaFree(L.mg_r); cudaFree(L.mg_t1); cudaFree(L.mg_t2);
  cudaFree(L.d_tiles_int); cudaFree(L.d_tiles_bnd); cudaFree(L.d_tiles_all); cudaFree(L.d_tile_desc); cudaFree(L.d_jinv);
  cudaFree(L.bcrs.d_rowptr); cudaFree(L.bcrs.d_col); cudaFree(L.bcrs.d_brow); cudaFree(L.bcrs.d_boff); cudaFree(L.bcrs.d_val);
  cudaFree(L.bcrs.d_wave); cudaFree(L.bcrs.d_res); cudaFree(L.bcrs.d_l1reg);
  for (int f = 0; f < 6; f++) {
    cudaFree(L.cg.d_send[f]); cudaFree(L.cg.d_recv[f]);
    cudaFree(L.hpg.d_deg[f]); cudaFree(L.hpg.d_pdeg[f]); cudaFree(L.hpg.d_troff[f]); cudaFree(L.hpg.d_recv[f]); cudaFree(L.hpg.d_send[f]);
    cudaFree(L.hpg.d_send_src[f]); cudaFree(L.hpg.d_send_dst[f]);
  }
}

int create_common(Ctx* ctx, int dim, const int* n, const double* Lx, const std::vector<int>& deg, double sigma,
                  int dirichlet, int device) {
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    ctx->err = "no CUDA device available (this library has no CPU fallback)";
    return 1;
  }
  ctx->device = device; ctx->dim = dim; ctx->sigma = sigma; ctx->dirichlet = dirichlet;
  HPDG_CUDA(cudaSetDevice(device));
  HPDG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  {
    // the halo stream gets the highest priority so that the NCCL kernel is scheduled as soon as any tile CTA retires
    // instead of queueing behind the interior-tile kernel (which would serialise the exchange after it)
    int lo = 0, hi = 0;
    HPDG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    HPDG_CUDA(cudaStreamCreateWithPriority(&ctx->stream_comm, cudaStreamNonBlocking, hi));
  }
  HPDG_CUDA(cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming));
  HPDG_CUDA(cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming));
  const HostTables& H = host_tables();
  HPDG_CUDA(cudaMalloc(&ctx->d_tab, sizeof(DegTable) * H.deg.size()));
  HPDG_CUDA(cudaMemcpy(ctx->d_tab, H.deg.data(), sizeof(DegTable) * H.deg.size(), cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMalloc(&ctx->d_P, sizeof(double) * H.P.size()));
  HPDG_CUDA(cudaMemcpy(ctx->d_P, H.P.data(), sizeof(double) * H.P.size(), cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMalloc(&ctx->d_T, sizeof(double) * H.T.size()));
  HPDG_CUDA(cudaMemcpy(ctx->d_T, H.T.data(), sizeof(double) * H.T.size(), cudaMemcpyHostToDevice));
  double h[3] = {1, 1, 1};
  for (int d = 0; d < dim; d++) h[d] = Lx[d] / n[d];
  ctx->levels.resize(1);
  ctx->levels[0].nc = ctx->create_nc;
  return setup_level(ctx, ctx->levels[0], dim, n, h, deg, deg);
}

Level* get_level(Ctx* ctx, int level) {
  int nl = (int)ctx->levels.size();
  if (level == HPDG_FINEST) level = nl - 1;
  if (level < 0 || level >= nl) { ctx->err = "level index out of range"; return nullptr; }
  return &ctx->levels[level];
}

int ensure_stage(Ctx* ctx, size_t ndof) {
  if (ctx->stage_cap >= ndof) return 0;
  cudaFree(ctx->d_in); cudaFree(ctx->d_out);
  ctx->d_in = ctx->d_out = nullptr; ctx->stage_cap = 0;
  HPDG_CUDA(cudaMalloc(&ctx->d_in, ndof * sizeof(double)));
  HPDG_CUDA(cudaMalloc(&ctx->d_out, ndof * sizeof(double)));
  ctx->stage_cap = ndof;
  return 0;
}

// Distributed apply.  Compute stream: interior tiles (need no ghost data).  Halo stream (highest priority, so its
// kernels are scheduled as soon as any tile CTA retires): pack the brick's face traces -> grouped ncclSend/ncclRecv with the
// <= 6 face neighbours (the copyFromMaster analogue, parallel/communicationhpdg.hh:411-418) -> rank-boundary tiles.
// The two streams write disjoint tiles of y; the compute stream joins the halo stream at the end.
int op_apply_distributed(Ctx* ctx, Level& L, const double* d_x, double* d_y, double factor) {
  const bool finest = (&L == &ctx->levels.back());
  Ghost* Gp = nullptr;
  if (level_ghost(ctx, L, &Gp)) return 1;
  Ghost& G = *Gp;
  if (finest && ctx->ghost.p2p) {
    // NVLink peer-memory halo, no NCCL call in the loop.  Halo stream (highest priority): pack kernel storing this rank's face
    // traces straight into the neighbours' arenas -> flag kernel.  Compute stream, concurrently: ONE tile kernel over all tiles,
    // interior tiles first; its rank-boundary tiles wait on the flags the NEIGHBOURS raise (the local pack is not a dependency
    // of the local apply).  The compute stream joins the halo stream at the end so that x may be overwritten afterwards.
    ctx->ghost.step++;
    if (uniform_persistent(ctx, L, d_x)) {
      // persistent Q3 kernel: packing the face traces, publishing the flags and all tiles are ONE launch (its CTAs fill every
      // SM for the whole run, so a separate pack kernel would not be scheduled next to it)
      return launch_apply_uniform(ctx, L, d_x, d_y, factor, 3, ctx->stream);
    }
    HPDG_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));                  // x is ready
    HPDG_CUDA(cudaStreamWaitEvent(ctx->stream_comm, ctx->ev_a, 0));
    if (launch_pack_traces(ctx, L, d_x, ctx->stream_comm)) return 1;
    if (launch_halo_flags(ctx, ctx->stream_comm)) return 1;
    HPDG_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream_comm));
    if (launch_apply_uniform(ctx, L, d_x, d_y, factor, 3, ctx->stream)) return 1;
    HPDG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    return 0;
  }
  if (!ctx->nccl) { ctx->err = "this distributed context has no NCCL communicator (created with nccl_id = NULL) and the peer-memory halo is not attached for this level"; return 1; }
  HPDG_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));                    // x is ready
  HPDG_CUDA(cudaStreamWaitEvent(ctx->stream_comm, ctx->ev_a, 0));
  if (launch_pack_traces(ctx, L, d_x, ctx->stream_comm)) return 1;
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  HPDG_NCCL(g_nccl.GroupStart());
  for (int f = 0; f < 6; f++) {
    if (!G.active[f]) continue;
    HPDG_NCCL(g_nccl.Send(G.d_send[f], G.count[f], ncclDouble, G.peer[f], comm, ctx->stream_comm));
    HPDG_NCCL(g_nccl.Recv(G.d_recv[f], G.count[f], ncclDouble, G.peer[f], comm, ctx->stream_comm));
  }
  HPDG_NCCL(g_nccl.GroupEnd());
  if (launch_apply_uniform(ctx, L, d_x, d_y, factor, 2, ctx->stream_comm)) return 1;   // rank-boundary tiles
  HPDG_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream_comm));
  if (launch_apply_uniform(ctx, L, d_x, d_y, factor, 1, ctx->stream)) return 1;        // interior tiles, overlaps the exchange
  HPDG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
  return 0;
}

int op_apply_async(Ctx* ctx, Level& L, const double* d_x, double* d_y, double factor) {
  if (ctx->nranks > 1 && uniform_supported(ctx, L) && !ctx->hp_distributed) return op_apply_distributed(ctx, L, d_x, d_y, factor);
  if (ctx->nranks == 1 && uniform_supported(ctx, L)) return launch_apply_uniform(ctx, L, d_x, d_y, factor, 0);
  return launch_apply_generic(ctx, L, d_x, d_y, factor);   // any degree map; distributed: hp face-trace halo over NCCL
}

int jacobi_async(Ctx* ctx, Level& L, int form, const double* r, double* c, double damping) {
  if (form == HPDG_JACOBI_DENSE) return jacobi_apply_dense(ctx, L, r, c, damping);
  if (form == HPDG_JACOBI_FD) return jacobi_apply_fd(ctx, L, r, c, damping);
  ctx->err = "unknown block-Jacobi form"; return 1;
}

// ---- distributed hp: ghost degrees (once per level) and face-trace halo (per apply) over NCCL -------------------------------------
// boundary elements of brick face f in face-element order (lower tangential direction fastest)
std::vector<long> face_elements(const Level& L, int f) {
  const int d = f / 2, s = f % 2, ta = d == 0 ? 1 : 0, tb = d == 2 ? 1 : 2;
  const long stride[3] = {1, L.n[0], (long)L.n[0] * L.n[1]};
  std::vector<long> out;
  for (int b = 0; b < (L.dim == 3 ? L.n[tb] : 1); b++)
    for (int a = 0; a < L.n[ta]; a++)
      out.push_back((s ? L.n[d] - 1 : 0) * stride[d] + a * stride[ta] + (L.dim == 3 ? b * stride[tb] : 0));
  return out;
}

}  // namespace
namespace hpdg {
int hp_ghost_setup(Ctx* ctx, Level& L) {
  HpGhost& G = L.hpg;
  if (G.ready) return 0;
  if (!ctx->nccl) { ctx->err = "the distributed hp apply needs the NCCL communicator (context created with nccl_id = NULL)"; return 1; }
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  if (generic_trace_setup(ctx, L)) return 1;
  // 1. exchange the finest-level degrees of the boundary elements (the analogue of parallel/updatedegrees.hh:11-46)
  int* d_sdeg[6] = {};
  std::vector<long> fel[6];
  for (int f = 0; f < 6; f++) {
    if (!ctx->ghost.active[f]) continue;
    fel[f] = face_elements(L, f);
    G.nface[f] = (long)fel[f].size();
    std::vector<int> pd(fel[f].size());
    for (size_t i = 0; i < fel[f].size(); i++) pd[i] = L.pdeg[fel[f][i]];
    HPDG_CUDA(cudaMalloc(&d_sdeg[f], sizeof(int) * pd.size()));
    HPDG_CUDA(cudaMemcpy(d_sdeg[f], pd.data(), sizeof(int) * pd.size(), cudaMemcpyHostToDevice));
    HPDG_CUDA(cudaMalloc(&G.d_pdeg[f], sizeof(int) * pd.size()));
    HPDG_CUDA(cudaMalloc(&G.d_deg[f], sizeof(int) * pd.size()));
  }
  HPDG_NCCL(g_nccl.GroupStart());
  for (int f = 0; f < 6; f++) {
    if (!ctx->ghost.active[f]) continue;
    HPDG_NCCL(g_nccl.Send(d_sdeg[f], (size_t)G.nface[f], ncclInt32, ctx->ghost.peer[f], comm, ctx->stream));
    HPDG_NCCL(g_nccl.Recv(G.d_pdeg[f], (size_t)G.nface[f], ncclInt32, ctx->ghost.peer[f], comm, ctx->stream));
  }
  HPDG_NCCL(g_nccl.GroupEnd());
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream));
  // 2. this level's ghost degrees (the level's cap applied, ordertransfer.hh:62-67), receive offsets, send gather lists
  G.maxp = 0;
  for (int f = 0; f < 6; f++) {
    if (!ctx->ghost.active[f]) continue;
    cudaFree(d_sdeg[f]);
    const long nfc = G.nface[f];
    G.h_pdeg[f].resize(nfc); G.h_deg[f].resize(nfc);
    HPDG_CUDA(cudaMemcpy(G.h_pdeg[f].data(), G.d_pdeg[f], sizeof(int) * nfc, cudaMemcpyDeviceToHost));
    std::vector<long> roff(nfc), src(nfc), dst(nfc + 1, 0);
    // (roff is kept as G.h_troff[f] for the level's face table)
    long rp = 0;
    for (long i = 0; i < nfc; i++) {
      if (G.h_pdeg[f][i] < 0 || G.h_pdeg[f][i] > kMaxP) { ctx->err = "ghost degree out of range (neighbour rank sent garbage?)"; return 1; }
      G.h_deg[f][i] = L.cap >= 0 ? std::min(G.h_pdeg[f][i], L.cap) : G.h_pdeg[f][i];
      G.maxp = std::max(G.maxp, G.h_deg[f][i]);
      roff[i] = rp; rp += ipow_h(G.h_deg[f][i] + 1, L.dim - 1);
      const long e = fel[f][i];
      const long nfe = ipow_h(L.deg[e] + 1, L.dim - 1);
      long tro = 0;  // offset of element e's traces: prefix over the elements' 2 dim N^(dim-1) pairs
      (void)tro;
      src[i] = -1; dst[i + 1] = dst[i] + nfe;
    }
    // element trace offsets (host prefix, as generic_trace_setup builds them)
    {
      std::vector<long> troff(L.nelem + 1, 0);
      for (long e = 0; e < L.nelem; e++) troff[e + 1] = troff[e] + 2 * L.dim * ipow_h(L.deg[e] + 1, L.dim - 1);
      for (long i = 0; i < nfc; i++) { const long e = fel[f][i]; src[i] = troff[e] + (long)f * ipow_h(L.deg[e] + 1, L.dim - 1); }
    }
    G.recv_pairs[f] = rp; G.send_pairs[f] = dst[nfc];
    HPDG_CUDA(cudaMemcpy(G.d_deg[f], G.h_deg[f].data(), sizeof(int) * nfc, cudaMemcpyHostToDevice));
    HPDG_CUDA(cudaMalloc(&G.d_troff[f], sizeof(long) * nfc));
    HPDG_CUDA(cudaMemcpy(G.d_troff[f], roff.data(), sizeof(long) * nfc, cudaMemcpyHostToDevice));
    G.h_troff[f] = roff;
    HPDG_CUDA(cudaMalloc(&G.d_send_src[f], sizeof(long) * nfc));
    HPDG_CUDA(cudaMemcpy(G.d_send_src[f], src.data(), sizeof(long) * nfc, cudaMemcpyHostToDevice));
    HPDG_CUDA(cudaMalloc(&G.d_send_dst[f], sizeof(long) * (nfc + 1)));
    HPDG_CUDA(cudaMemcpy(G.d_send_dst[f], dst.data(), sizeof(long) * (nfc + 1), cudaMemcpyHostToDevice));
    HPDG_CUDA(cudaMalloc(&G.d_recv[f], sizeof(double) * 2 * std::max<long>(rp, 1)));
    HPDG_CUDA(cudaMalloc(&G.d_send[f], sizeof(double) * 2 * std::max<long>(dst[nfc], 1)));
  }
  G.ready = true;
  return 0;
}
// Per apply (after k_face_traces): gather this rank's boundary traces, one grouped ncclSend/ncclRecv with the <= 6 face neighbours
// (variable-size blocks per element: parallel/communicationhpdg.hh:387-418), all on the context stream.
int hp_halo_exchange(Ctx* ctx, Level& L) {
  if (hp_ghost_setup(ctx, L)) return 1;
  if (launch_hp_pack(ctx, L, ctx->stream)) return 1;
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  HPDG_NCCL(g_nccl.GroupStart());
  for (int f = 0; f < 6; f++) {
    if (!ctx->ghost.active[f]) continue;
    HPDG_NCCL(g_nccl.Send(L.hpg.d_send[f], (size_t)L.hpg.send_pairs[f] * 2, ncclDouble, ctx->ghost.peer[f], comm, ctx->stream));
    HPDG_NCCL(g_nccl.Recv(L.hpg.d_recv[f], (size_t)L.hpg.recv_pairs[f] * 2, ncclDouble, ctx->ghost.peer[f], comm, ctx->stream));
  }
  HPDG_NCCL(g_nccl.GroupEnd());
  return 0;
}
}  // namespace hpdg

namespace {
// ---- V-cycle (iterationsteps/mg/multigrid_impl.hh:16-117) -----------------------------------------
struct VC { int form; double damping; int pre, post, coarse_its; };

int mg_smooth(Ctx* ctx, int l, const VC& v, int steps, double* x, double* r) {
  Level& L = ctx->levels[l];
  if (v.form == HPDG_SMOOTHER_BLOCKGS || v.form == HPDG_SMOOTHER_BLOCKGS_MF) {
    // the reference's default: smootherFromIterationStep2 around DynamicBlockGS on the level's assembled (Galerkin) matrix
    // (solversetup.hh:139-145, multigrid.hh:96-107).  tmp1 is zeroed once per applySmoother (multigrid_impl.hh:73) and NOT between
    // steps: iterate() continues from the previous tmp1, exactly as the reference does.
    HPDG_CUDA(cudaMemsetAsync(L.mg_t1, 0, sizeof(double) * L.ndof, ctx->stream));
    for (int i = 0; i < steps; i++) {
      if (v.form == HPDG_SMOOTHER_BLOCKGS ? blockgs_iterate(ctx, L, r, L.mg_t1) : blockgs_mf_iterate(ctx, L, r, L.mg_t1)) return 1;
      if (v.damping != 1.0) { if (launch_axpy(ctx, L.ndof, v.damping - 1.0, L.mg_t1, L.mg_t1)) return 1; }
      if (launch_axpy(ctx, L.ndof, 1.0, L.mg_t1, x)) return 1;
      ctx->fuse_accum = 1;
      const int rc = op_apply_async(ctx, L, L.mg_t1, r, -1.0);
      ctx->fuse_accum = 0;
      if (rc) return 1;
    }
    return 0;
  }
  for (int i = 0; i < steps; i++) {                                   // multigrid_impl.hh:76-81
    // smoother(tmp1, r); x += tmp1 -- no separate axpy: on uniform levels the update rides on the operator kernel that applies
    // tmp1 next (those kernels are issue bound, the extra 16 B/DoF are free: measured 10.4 -> 8.1 -> ... ms per 64^3 cycle,
    // DESIGN.md section 6a), otherwise it is fused into the Jacobi kernel
    const bool in_apply = ctx->xacc_in_apply && uniform_supported(ctx, L);
    if (!in_apply) ctx->fuse_xacc = x;
    int rc = jacobi_async(ctx, L, v.form, r, L.mg_t1, v.damping);
    ctx->fuse_xacc = nullptr;
    if (rc) return 1;
    // tmp2 = A tmp1; r -= tmp2 -- fused: r = r + (-1) * A tmp1 in the operator kernel's store
    ctx->fuse_accum = 1;
    if (in_apply) ctx->fuse_xin = x;
    rc = op_apply_async(ctx, L, L.mg_t1, r, -1.0);
    ctx->fuse_accum = 0; ctx->fuse_xin = nullptr;
    if (rc) return 1;
  }
  return 0;
}

int mg_level(Ctx* ctx, int l, const VC& v) {
  Level& L = ctx->levels[l];
  double* x = L.mg_x; double* r = L.mg_r;
  if (l == 0 && (v.form == HPDG_SMOOTHER_BLOCKGS || v.form == HPDG_SMOOTHER_BLOCKGS_MF)) {  // coarse solver of the reference: coarse_its block-GS iterations (solversetup.hh:198-215)
    for (int i = 0; i < v.coarse_its; i++)
      if (v.form == HPDG_SMOOTHER_BLOCKGS ? blockgs_iterate(ctx, L, r, x) : blockgs_mf_iterate(ctx, L, r, x)) return 1;
    return 0;
  }
  if (l == 0) {  // coarse solver: coarse_its damped block-Jacobi iterations from x = 0 (cf. solversetup.hh:198-215)
    for (int i = 0; i < v.coarse_its; i++) {
      if (op_apply_async(ctx, L, x, L.mg_t1, 1.0)) return 1;
      if (launch_xpay_sub(ctx, L.ndof, r, L.mg_t1, L.mg_t2)) return 1;
      if (jacobi_async(ctx, L, v.form, L.mg_t2, L.mg_t1, v.damping)) return 1;
      if (launch_axpy(ctx, L.ndof, 1.0, L.mg_t1, x)) return 1;
    }
    return 0;
  }
  if (mg_smooth(ctx, l, v, v.pre, x, r)) return 1;                      // :99
  Level& C = ctx->levels[l - 1];
  if (launch_restrict(ctx, L, C, r, C.mg_r)) return 1;                  // :103
  HPDG_CUDA(cudaMemsetAsync(C.mg_x, 0, sizeof(double) * C.ndof, ctx->stream));
  if (mg_level(ctx, l - 1, v)) return 1;                                // mu_ = 1
  if (launch_prolong(ctx, L, C, C.mg_x, L.mg_t1)) return 1;             // :108
  // x += tmp1 and r -= A tmp1 (:110-112): both ride on the operator kernel on uniform levels (no axpy pass)
  const bool in_apply = ctx->xacc_in_apply && uniform_supported(ctx, L);
  if (!in_apply && launch_axpy(ctx, L.ndof, 1.0, L.mg_t1, x)) return 1;
  ctx->fuse_accum = 1;
  if (in_apply) ctx->fuse_xin = x;
  const int rcc = op_apply_async(ctx, L, L.mg_t1, r, -1.0);
  ctx->fuse_accum = 0; ctx->fuse_xin = nullptr;
  if (rcc) return 1;
  return mg_smooth(ctx, l, v, v.post, x, r);                            // :116
}

int vcycle_device(Ctx* ctx, const VC& v, double* d_x, double* d_b) {
  if (ctx->nranks > 1 && v.form != HPDG_JACOBI_FD) { ctx->err = "the distributed V-cycle supports the fd block-Jacobi smoother only"; return 1; }
  const int nl = (int)ctx->levels.size();
  for (int l = 0; l < nl; l++) {
    Level& L = ctx->levels[l];
    if (!L.mg_x) {
      HPDG_CUDA(cudaMalloc(&L.mg_x, sizeof(double) * L.ndof));
      HPDG_CUDA(cudaMalloc(&L.mg_r, sizeof(double) * L.ndof));
      HPDG_CUDA(cudaMalloc(&L.mg_t1, sizeof(double) * L.ndof));
      HPDG_CUDA(cudaMalloc(&L.mg_t2, sizeof(double) * L.ndof));
    }
    if (v.form == HPDG_SMOOTHER_BLOCKGS) { if (bcrs_build(ctx, L)) return 1; }
    else if (v.form == HPDG_SMOOTHER_BLOCKGS_MF) {}
    else if (v.form == HPDG_JACOBI_DENSE ? !L.jd.ready : !L.jf.ready) {
      if (v.form == HPDG_JACOBI_DENSE ? jacobi_setup_dense(ctx, L) : jacobi_setup_fd(ctx, L)) return 1;
    }
  }
  Level& F = ctx->levels[nl - 1];
  if (nl > 1) {
    // On the finest level of a hierarchy the cycle only ever ADDS to its iterate (smoothing steps, prolongated corrections) and
    // subtracts A * (what it added) from its residual, so it works directly on the caller's vectors: b <- b - A x in one
    // accumulating apply (:30-36), then x and b play state.x / state.r of the finest level.  On return x += correction and b is
    // the residual (:60-61) without the final axpy / copy (64 B/DoF of vector traffic per cycle less).
    ctx->fuse_accum = 1;
    const int r0 = op_apply_async(ctx, F, d_x, d_b, -1.0);
    ctx->fuse_accum = 0;
    if (r0) return 1;
    double *sx = F.mg_x, *sr = F.mg_r;
    F.mg_x = d_x; F.mg_r = d_b;
    const int r1 = mg_level(ctx, nl - 1, v);
    F.mg_x = sx; F.mg_r = sr;
    return r1;
  }
  HPDG_CUDA(cudaMemsetAsync(F.mg_x, 0, sizeof(double) * F.ndof, ctx->stream));
  if (op_apply_async(ctx, F, d_x, F.mg_t1, 1.0)) return 1;             // r = b - A x (:30-36)
  if (launch_xpay_sub(ctx, F.ndof, d_b, F.mg_t1, F.mg_r)) return 1;
  if (mg_level(ctx, nl - 1, v)) return 1;
  if (launch_axpy(ctx, F.ndof, 1.0, F.mg_x, d_x)) return 1;            // x += state.x[fine] (:60)
  HPDG_CUDA(cudaMemcpyAsync(d_b, F.mg_r, sizeof(double) * F.ndof, cudaMemcpyDeviceToDevice, ctx->stream));  // b = r (:61)
  return 0;
}
}  // namespace

namespace hpdg {
int kernel_slots(Ctx* ctx, const void* func, int threads, size_t smem, int* slots) {
  auto it = ctx->kattr.find(func);
  if (it == ctx->kattr.end()) {
    if (smem > 48 * 1024) HPDG_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nsm = 0, occ = 0;
    HPDG_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device));
    HPDG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, threads, smem));
    it = ctx->kattr.emplace(func, nsm * std::max(occ, 1)).first;
  }
  if (slots) *slots = it->second;
  return 0;
}
}  // namespace hpdg

// CUDA events around `reps` asynchronous launches on the context stream (events are destroyed on every path)
template <class Launch>
static int time_launches(Ctx* ctx, int reps, float* ms_per_launch, Launch launch) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = 0;
  float ms = 0;
  auto ck = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && !rc) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); rc = 1; } };
  ck(cudaEventCreate(&e0), "cudaEventCreate"); ck(cudaEventCreate(&e1), "cudaEventCreate");
  if (!rc) ck(cudaEventRecord(e0, ctx->stream), "cudaEventRecord");
  for (int i = 0; i < reps && !rc; i++) rc = launch();
  if (!rc) ck(cudaEventRecord(e1, ctx->stream), "cudaEventRecord");
  if (!rc) ck(cudaEventSynchronize(e1), "cudaEventSynchronize");
  if (!rc) ck(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (rc) return 1;
  *ms_per_launch = ms / std::max(reps, 1);
  return sync_check(ctx);
}

// =================================================================================================
extern "C" {

int hpdg_create(hpdg_ctx** out, int dim, const int* n, const double* L, const int* degree, long ndegree,
                double sigma, int dirichlet, int device) {
  *out = nullptr;
  if (dim != 2 && dim != 3) { g_create_err = "dim must be 2 or 3"; return 1; }
  long nelem = 1;
  for (int d = 0; d < dim; d++) { if (n[d] < 1) { g_create_err = "mesh extents must be positive"; return 1; } nelem *= n[d]; }
  if (ndegree != 1 && ndegree != nelem) { g_create_err = "degree array must have 1 or nelem entries"; return 1; }
  std::vector<int> deg(nelem);
  for (long e = 0; e < nelem; e++) {
    deg[e] = degree[ndegree == 1 ? 0 : e];
    if (deg[e] < 0 || deg[e] > kMaxP) { g_create_err = "polynomial degree out of range 0..13"; return 1; }
  }
  hpdg_ctx* ctx = new hpdg_ctx();
  if (create_common(ctx, dim, n, L, deg, sigma, dirichlet, device)) { g_create_err = ctx->err; delete ctx; return 1; }
  *out = ctx;
  return 0;
}

// Non-conforming 2-D mesh (reference: the hanging-node branch of SumFactIPDGOperator, sfipdg.hh:213-222,472-491): the
// base grid with the flagged cells split once into 2 x 2 children.  Builds the leaf numbering and, per (leaf, side), its one or
// two intersections: neighbour leaf, kind (conforming / coarse side half / fine side), penalty max(p-, p+)^2, normal sign * kappa.
int hpdg_create_refined_2d(hpdg_ctx** out, const int* n, const double* L, const unsigned char* refine, const int* degree,
                           long nleaf, double sigma, int dirichlet, int device) {
  *out = nullptr;
  if (!n || !L || !refine || !degree) { g_create_err = "null argument"; return 1; }
  if (n[0] < 1 || n[1] < 1) { g_create_err = "mesh extents must be positive"; return 1; }
  const long ncell = (long)n[0] * n[1];
  std::vector<long> first(ncell + 1, 0);
  for (long c = 0; c < ncell; c++) first[c + 1] = first[c] + (refine[c] ? 4 : 1);
  if (nleaf != first[ncell]) { g_create_err = "degree array must have one entry per leaf element (unrefined cells + 4 per refined cell)"; return 1; }
  std::vector<int> deg(nleaf);
  for (long e = 0; e < nleaf; e++) {
    deg[e] = degree[e];
    if (deg[e] < 0 || deg[e] > kMaxP) { g_create_err = "polynomial degree out of range 0..13"; return 1; }
  }
  hpdg_ctx* ctx = new hpdg_ctx();
  ctx->create_nc = true;
  if (create_common(ctx, 2, n, L, deg, sigma, dirichlet, device)) { g_create_err = ctx->err; delete ctx; return 1; }
  Level& Lv = ctx->levels[0];
  const double h[2] = {L[0] / n[0], L[1] / n[1]};
  Lv.nc_faces.assign((size_t)nleaf * 8, FaceInfo());
  Lv.nc_nbr.assign((size_t)nleaf * 8, -1);
  auto set_face = [&](long e, int f, int slot, long o, int kind) {
    const int d = f / 2, s = f % 2;
    FaceInfo F;
    const double kappa = d == 0 ? h[1] / h[0] : h[0] / h[1];   // h_t / h_n: the same for a cell and its children
    F.nuk = s ? kappa : -kappa; F.tro = 0; F.ghost = 0; F.kind = (short)kind; F.po = (short)deg[e];
    if (o >= 0) {
      F.po = (short)deg[o];
      const int pm = std::max(deg[e], deg[o]);
      F.cpen = sigma * (double)pm * pm;
      F.mode = (kind == 0 && deg[o] == deg[e]) ? 2 : 3;
    } else if (o == -1) {
      F.cpen = sigma * (double)deg[e] * deg[e];
      F.mode = dirichlet ? 1 : 0;
    } else { F.cpen = 0; F.mode = -1; }   // o == -2: unused second slot
    Lv.nc_faces[(size_t)e * 8 + f * 2 + slot] = F;
    Lv.nc_nbr[(size_t)e * 8 + f * 2 + slot] = o;
  };
  for (int cy = 0; cy < n[1]; cy++) for (int cx = 0; cx < n[0]; cx++) {
    const long c = cx + (long)n[0] * cy;
    for (int f = 0; f < 4; f++) {
      const int d = f / 2, s = f % 2;
      const int ox = cx + (d == 0 ? (s ? 1 : -1) : 0), oy = cy + (d == 1 ? (s ? 1 : -1) : 0);
      const bool inside = ox >= 0 && ox < n[0] && oy >= 0 && oy < n[1];
      const long oc = inside ? ox + (long)n[0] * oy : -1;
      // child (a, b) of a refined cell is leaf first[cell] + a + 2 b; the children of the neighbour cell that touch the shared face
      // have normal index (s ? 0 : 1); their tangential index t selects the half of the face
      auto nb_child = [&](int t) { return first[oc] + (d == 0 ? (s ? 0 : 1) + 2 * t : t + 2 * (s ? 0 : 1)); };
      if (!refine[c]) {
        const long e = first[c];
        if (!inside) { set_face(e, f, 0, -1, 0); set_face(e, f, 1, -2, 0); }
        else if (!refine[oc]) { set_face(e, f, 0, first[oc], 0); set_face(e, f, 1, -2, 0); }
        else { set_face(e, f, 0, nb_child(0), 1); set_face(e, f, 1, nb_child(1), 2); }   // coarse side: low / high half
      } else {
        for (int b = 0; b < 2; b++) for (int a = 0; a < 2; a++) {
          const long e = first[c] + a + 2 * b;
          const int nrm = d == 0 ? a : b, tng = d == 0 ? b : a;   // the child's index normal / tangential to the face
          set_face(e, f, 1, -2, 0);
          if (nrm != s) set_face(e, f, 0, first[c] + (d == 0 ? (1 - a) + 2 * b : a + 2 * (1 - b)), 0);   // sibling
          else if (!inside) set_face(e, f, 0, -1, 0);
          else if (refine[oc]) set_face(e, f, 0, nb_child(tng), 0);
          else set_face(e, f, 0, first[oc], 3 + tng);   // fine side on the low / high half of the coarse neighbour's side
        }
      }
    }
  }
  *out = ctx;
  return 0;
}

int hpdg_nccl_unique_id(void* out128) {
  std::string err;
  if (!load_nccl(err)) { g_create_err = err; return 1; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_err = "ncclGetUniqueId failed"; return 1; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out128, &id, 128);
  return 0;
}

// common part of the distributed creators: `deg` holds one degree per local element; hp = per-element degree map (generic path)
static int create_distributed_common(hpdg_ctx** out, int dim, const int* n, const double* L, const std::vector<int>& deg, bool hp,
                                     double sigma, int dirichlet, int device, const int* pgrid, int rank, int nranks,
                                     const void* nccl_id) {
  const long nelem = (long)n[0] * n[1] * n[2];
  hpdg_ctx* ctx = new hpdg_ctx();
  if (create_common(ctx, dim, n, L, deg, sigma, dirichlet, device)) { g_create_err = ctx->err; delete ctx; return 1; }
  ctx->rank = rank; ctx->nranks = nranks; ctx->hp_distributed = hp;
  for (int d = 0; d < 3; d++) ctx->pgrid[d] = pgrid[d];
  ctx->pcoord[0] = rank % pgrid[0]; ctx->pcoord[1] = (rank / pgrid[0]) % pgrid[1]; ctx->pcoord[2] = rank / (pgrid[0] * pgrid[1]);
  auto fail = [&](const std::string& e) { g_create_err = e; hpdg_destroy(ctx); return 1; };
  if (nranks > 1) {
    if (nccl_id) {  // nccl_id == NULL: no NCCL communicator (peer-memory halo only; entry points that need NCCL then fail)
      std::string err;
      if (!load_nccl(err)) return fail(err);
      ncclUniqueId id; memcpy(&id, nccl_id, 128);
      ncclComm_t comm;
      if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
      ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
      if (r != ncclSuccess) return fail(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
      ctx->nccl = comm;
    }
    const int pstride[3] = {1, pgrid[0], pgrid[0] * pgrid[1]};
    for (int f = 0; f < 6; f++) {
      const int d = f / 2, s = f % 2;
      const int c = ctx->pcoord[d] + (s ? 1 : -1);
      if (c < 0 || c >= pgrid[d]) continue;
      ctx->bnd_is_rank[f] = true;
      ctx->ghost.active[f] = true;
      ctx->ghost.peer[f] = rank + (s ? pstride[d] : -pstride[d]);
    }
    if (!hp) {
      // uniform degree: fixed-size trace buffers for the NCCL transport and the arena of the peer-memory transport
      const int N2 = (deg[0] + 1) * (deg[0] + 1);
      for (int f = 0; f < 6; f++) {
        if (!ctx->ghost.active[f]) continue;
        size_t felems = (size_t)nelem / n[f / 2];
        ctx->ghost.count[f] = felems * N2 * 2;
        if (cudaMalloc(&ctx->ghost.d_send[f], ctx->ghost.count[f] * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&ctx->ghost.d_recv[f], ctx->ghost.count[f] * sizeof(double)) != cudaSuccess)
          return fail("cudaMalloc of halo buffers failed");
      }
      // arena for the peer-to-peer mode: identical layout on every rank (all six faces, sized from the brick shape)
      size_t offb = 0;
      for (int f = 0; f < 6; f++) {
        const size_t bytes = ((size_t)nelem / n[f / 2]) * N2 * 2 * sizeof(double);
        for (int par = 0; par < 2; par++) { ctx->ghost.recv_off[f][par] = offb; offb += (bytes + 255) / 256 * 256; }
      }
      ctx->ghost.flag_off = offb; offb += 256;
      ctx->ghost.arena_bytes = offb;
      if (cudaMalloc(&ctx->ghost.arena, offb) != cudaSuccess || cudaMemset(ctx->ghost.arena, 0, offb) != cudaSuccess)
        return fail("cudaMalloc of the halo arena failed");
      int* herr = nullptr;
      if (cudaHostAlloc(&herr, sizeof(int), cudaHostAllocMapped) != cudaSuccess) return fail("cudaHostAlloc of the halo time-out flag failed");
      *herr = 0;
      ctx->h_ghost_err = herr;
      if (cudaHostGetDevicePointer(&ctx->d_ghost_err, herr, 0) != cudaSuccess) return fail("cudaHostGetDevicePointer failed");
    }
  }
  *out = ctx;
  return 0;
}

static int check_distributed_args(int dim, const int* n, const double* L, const int* pgrid, int rank, int nranks) {
  if (dim != 3) { g_create_err = "distributed bricks are 3-D"; return 1; }
  if (!n || !L || !pgrid) { g_create_err = "null argument"; return 1; }
  for (int d = 0; d < 3; d++) {
    if (n[d] < 1) { g_create_err = "mesh extents must be positive"; return 1; }
    if (pgrid[d] < 1) { g_create_err = "pgrid entries must be positive"; return 1; }
  }
  if (nranks < 1 || pgrid[0] * pgrid[1] * pgrid[2] != nranks) { g_create_err = "pgrid does not match nranks"; return 1; }
  if (rank < 0 || rank >= nranks) { g_create_err = "rank out of range"; return 1; }
  return 0;
}

int hpdg_create_distributed(hpdg_ctx** out, int dim, const int* n, const double* L, int degree, double sigma,
                            int dirichlet, int device, const int* pgrid, int rank, int nranks, const void* nccl_id) {
  *out = nullptr;
  if (check_distributed_args(dim, n, L, pgrid, rank, nranks)) return 1;
  if (degree < 0 || degree > kMaxP) { g_create_err = "polynomial degree out of range 0..13"; return 1; }
  if (nranks > 1 && (degree < 1 || degree > 5)) { g_create_err = "the uniform distributed path needs a degree with a specialised kernel (1..5); use hpdg_create_distributed_hp"; return 1; }
  std::vector<int> deg((size_t)n[0] * n[1] * n[2], degree);
  return create_distributed_common(out, dim, n, L, deg, false, sigma, dirichlet, device, pgrid, rank, nranks, nccl_id);
}

int hpdg_create_distributed_hp(hpdg_ctx** out, int dim, const int* n, const double* L, const int* degree, double sigma,
                               int dirichlet, int device, const int* pgrid, int rank, int nranks, const void* nccl_id) {
  *out = nullptr;
  if (check_distributed_args(dim, n, L, pgrid, rank, nranks)) return 1;
  if (!degree) { g_create_err = "degree array is null"; return 1; }
  if (nranks > 1 && !nccl_id) { g_create_err = "the distributed hp path exchanges its halo over NCCL: nccl_id must not be null"; return 1; }
  const long nelem = (long)n[0] * n[1] * n[2];
  std::vector<int> deg(nelem);
  for (long e = 0; e < nelem; e++) {
    deg[e] = degree[e];
    if (deg[e] < 0 || deg[e] > kMaxP) { g_create_err = "polynomial degree out of range 0..13"; return 1; }
  }
  return create_distributed_common(out, dim, n, L, deg, true, sigma, dirichlet, device, pgrid, rank, nranks, nccl_id);
}

void hpdg_destroy(hpdg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->stream_comm) cudaStreamSynchronize(ctx->stream_comm);
  if (ctx->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl);
  for (auto& L : ctx->levels) free_level(L);
  for (int f = 0; f < 6; f++) { cudaFree(ctx->ghost.d_send[f]); cudaFree(ctx->ghost.d_recv[f]); if (ctx->ghost.peer_arena[f]) cudaIpcCloseMemHandle(ctx->ghost.peer_arena[f]); }
  cudaFree(ctx->ghost.arena);
  if (ctx->h_ghost_err) cudaFreeHost(const_cast<int*>(ctx->h_ghost_err));
  cudaFree(ctx->d_scalar); cudaFree(ctx->d_partial);
  if (ctx->cg_p) { cudaFree(ctx->cg_p); cudaFree(ctx->cg_q); cudaFree(ctx->cg_r); cudaFree(ctx->cg_z); }
  cudaFree(ctx->d_tab); cudaFree(ctx->d_sched); cudaFree(ctx->d_P); cudaFree(ctx->d_T); cudaFree(ctx->d_Mab); cudaFree(ctx->d_Pnc_eo); cudaFree(ctx->d_Pnc_ee); cudaFree(ctx->d_in); cudaFree(ctx->d_out);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->stream_comm) cudaStreamDestroy(ctx->stream_comm);
  if (ctx->bucket_stream[0]) { for (int k = 0; k < kBucketStreams; k++) cudaStreamDestroy(ctx->bucket_stream[k]); for (int k = 0; k <= kBucketStreams; k++) cudaEventDestroy(ctx->bucket_ev[k]); }
  if (ctx->stream_h2d) { cudaStreamDestroy(ctx->stream_h2d); cudaStreamDestroy(ctx->stream_d2h); for (int k = 0; k < 3; k++) for (int c = 0; c < 32; c++) cudaEventDestroy(ctx->ev_chunk[k][c]); }
  delete ctx;
}

const char* hpdg_last_error(const hpdg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int hpdg_set_option(hpdg_ctx* ctx, const char* name, long value) {
  HPDG_ENTER(ctx);
  if (!strcmp(name, "force_generic")) { ctx->force_generic = (int)value; return 0; }
  if (!strcmp(name, "variant")) { ctx->variant = (int)value; return 0; }
  if (!strcmp(name, "xacc_in_apply")) { ctx->xacc_in_apply = (int)value; return 0; }
  if (!strcmp(name, "q3p_grid")) { ctx->q3p_grid = (int)value; return 0; }
  if (!strcmp(name, "q3p_tune")) { ctx->q3p_tune = (int)value; return 0; }
  if (!strcmp(name, "halo_timeout_ms")) { ctx->halo_timeout_cycles = (long long)value * 2000000LL; return 0; }  // ~2 GHz SM clock
  if (!strcmp(name, "halo_p2p")) {  // switch between the NVLink peer-memory halo and NCCL send/recv (attach must have succeeded for 1)
    if (value && !ctx->ghost.peer_attached) { ctx->err = "halo_p2p: hpdg_halo_ipc_attach has not succeeded on this context"; return 1; }
    ctx->ghost.p2p = value != 0; return 0;
  }
  ctx->err = std::string("unknown option ") + name; return 1;
}

int hpdg_num_levels(const hpdg_ctx* ctx) { return (int)ctx->levels.size(); }
long hpdg_num_elements(const hpdg_ctx* ctx) { return ctx->levels.back().nelem; }
long hpdg_dimension(const hpdg_ctx* ctx, int level) {
  Level* L = get_level(const_cast<hpdg_ctx*>(ctx), level);
  return L ? L->ndof : -1;
}
int hpdg_block_offsets(const hpdg_ctx* ctx, int level, long* offsets) {
  Level* L = get_level(const_cast<hpdg_ctx*>(ctx), level);
  if (!L) return 1;
  memcpy(offsets, L->off.data(), sizeof(long) * (L->nelem + 1));
  return 0;
}
int hpdg_level_degrees(const hpdg_ctx* ctx, int level, int* degree) {
  Level* L = get_level(const_cast<hpdg_ctx*>(ctx), level);
  if (!L) return 1;
  memcpy(degree, L->deg.data(), sizeof(int) * L->nelem);
  return 0;
}

int hpdg_build_p_hierarchy(hpdg_ctx* ctx) {
  HPDG_ENTER(ctx);
  if (ctx->levels.back().nc) { ctx->err = "p-hierarchy: not available on non-conforming meshes (operator apply only)"; return 1; }
  if (ctx->levels.size() != 1) { ctx->err = "hierarchy already built"; return 1; }
  Level fine = ctx->levels[0];
  const int pmax = fine.maxp;
  const int pLevels = pmax >= 1 ? (int)std::log2((double)pmax) : 0;   // solversetup.hh:77
  std::vector<Level> lv(pLevels + 1);
  lv[pLevels] = fine;
  ctx->levels.clear();
  for (int idx = pLevels - 1; idx >= 0; idx--) {
    const int cap = pmax / ((pLevels - idx) * 2);                     // solversetup.hh:94
    std::vector<int> deg(fine.nelem);
    for (long e = 0; e < fine.nelem; e++) deg[e] = std::min(lv[idx + 1].deg[e], cap);  // ordertransfer.hh:62-67
    Level L;
    L.cap = cap;
    if (setup_level(ctx, L, fine.dim, fine.n, fine.h, deg, fine.pdeg)) {  // leave the context as it was: one level
      free_level(L);
      for (int k = idx + 1; k < pLevels; k++) free_level(lv[k]);
      ctx->levels.assign(1, fine);
      return 1;
    }
    lv[idx] = L;
  }
  ctx->levels = lv;
  return 0;
}

int hpdg_vec_alloc(hpdg_ctx* ctx, int level, double** d_vec) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  HPDG_CUDA(cudaMalloc(d_vec, sizeof(double) * std::max<long>(L->ndof, 1)));
  HPDG_CUDA(cudaMemsetAsync(*d_vec, 0, sizeof(double) * L->ndof, ctx->stream));
  return sync_check(ctx);
}
int hpdg_vec_free(hpdg_ctx* ctx, double* d_vec) { HPDG_ENTER(ctx); HPDG_CUDA(cudaFree(d_vec)); return 0; }
int hpdg_vec_upload(hpdg_ctx* ctx, int level, const double* h_src, double* d_dst) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  HPDG_CUDA(cudaMemcpyAsync(d_dst, h_src, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  return sync_check(ctx);
}
int hpdg_vec_download(hpdg_ctx* ctx, int level, const double* d_src, double* h_dst) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}
int hpdg_host_alloc(hpdg_ctx* ctx, size_t bytes, void** h_ptr) { HPDG_ENTER(ctx); HPDG_CUDA(cudaMallocHost(h_ptr, bytes)); return 0; }
int hpdg_host_free(hpdg_ctx* ctx, void* h_ptr) { HPDG_ENTER(ctx); HPDG_CUDA(cudaFreeHost(h_ptr)); return 0; }
int hpdg_sync(hpdg_ctx* ctx) { HPDG_ENTER(ctx); return sync_check(ctx); }

int hpdg_halo_ipc_handle(hpdg_ctx* ctx, void* out64) {
  HPDG_ENTER(ctx);
  if (!ctx->ghost.arena) { ctx->err = "not a distributed context"; return 1; }
  cudaIpcMemHandle_t h;
  HPDG_CUDA(cudaIpcGetMemHandle(&h, ctx->ghost.arena));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
  memcpy(out64, &h, 64);
  return 0;
}
int hpdg_halo_ipc_attach(hpdg_ctx* ctx, const void* handles /* nranks x 64 bytes, indexed by rank */) {
  HPDG_ENTER(ctx);
  if (!ctx->ghost.arena) { ctx->err = "not a distributed context"; return 1; }
  for (int f = 0; f < 6; f++) {
    if (!ctx->ghost.active[f]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)ctx->ghost.peer[f] * 64, 64);
    void* p = nullptr;
    HPDG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ghost.peer_arena[f] = static_cast<char*>(p);
  }
  ctx->ghost.peer_attached = true;
  ctx->ghost.p2p = true;
  return 0;
}
void* hpdg_stream(hpdg_ctx* ctx) { return (void*)ctx->stream; }

int hpdg_op_apply_async(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  return op_apply_async(ctx, *L, d_x, d_y, factor);
}
int hpdg_op_apply_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor) {
  HPDG_ENTER(ctx);
  if (hpdg_op_apply_async(ctx, level, d_x, d_y, factor)) return 1;
  return sync_check(ctx);
}
// Host-pointer apply as a 3-stage pipeline over z-slabs: H2D of slab c+1, tile kernel on slab c and D2H of slab c-1 run on
// three streams, so the PCIe copies in both directions overlap each other and the compute (the kernel on slab c needs slabs
// c-1..c+1 resident because of the z-neighbour traces).
static int op_apply_host_chunked(hpdg_ctx* ctx, Level& L, const double* h_x, double* h_y, double factor) {
  HPDG_ENTER(ctx);
  const int th = uniform_tile_height(L);
  const int layers = L.n[2];
  int nchunk = std::min(16, layers / th);
  const int per = ((layers + nchunk - 1) / nchunk + th - 1) / th * th;   // layers per chunk, multiple of the tile height
  nchunk = (layers + per - 1) / per;
  if (!ctx->stream_h2d) {
    HPDG_CUDA(cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking));
    HPDG_CUDA(cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking));
    for (int k = 0; k < 3; k++) for (int c = 0; c < 32; c++) HPDG_CUDA(cudaEventCreateWithFlags(&ctx->ev_chunk[k][c], cudaEventDisableTiming));
  }
  const size_t layer_dofs = (size_t)L.ndof / layers;
  HPDG_CUDA(cudaEventRecord(ctx->ev_chunk[2][0], ctx->stream));            // earlier work on the context stream is done
  HPDG_CUDA(cudaStreamWaitEvent(ctx->stream_h2d, ctx->ev_chunk[2][0], 0));
  HPDG_CUDA(cudaStreamWaitEvent(ctx->stream_d2h, ctx->ev_chunk[2][0], 0));
  for (int c = 0; c < nchunk; c++) {
    const size_t z0 = (size_t)c * per, nz = std::min<size_t>(per, layers - z0);
    HPDG_CUDA(cudaMemcpyAsync(ctx->d_in + z0 * layer_dofs, h_x + z0 * layer_dofs, nz * layer_dofs * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream_h2d));
    HPDG_CUDA(cudaEventRecord(ctx->ev_chunk[0][c], ctx->stream_h2d));
  }
  for (int c = 0; c < nchunk; c++) {
    const int z0 = c * per, nz = std::min(per, layers - z0);
    HPDG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[0][std::min(c + 1, nchunk - 1)], 0));
    ctx->slab_z0 = z0; ctx->slab_nz = nz;
    const int rc = launch_apply_uniform(ctx, L, ctx->d_in, ctx->d_out, factor, 0);
    ctx->slab_z0 = ctx->slab_nz = 0;
    if (rc) return 1;
    HPDG_CUDA(cudaEventRecord(ctx->ev_chunk[1][c], ctx->stream));
    HPDG_CUDA(cudaStreamWaitEvent(ctx->stream_d2h, ctx->ev_chunk[1][c], 0));
    HPDG_CUDA(cudaMemcpyAsync(h_y + (size_t)z0 * layer_dofs, ctx->d_out + (size_t)z0 * layer_dofs, (size_t)nz * layer_dofs * sizeof(double),
                              cudaMemcpyDeviceToHost, ctx->stream_d2h));
  }
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream_d2h));
  return sync_check(ctx);
}

int hpdg_op_apply(hpdg_ctx* ctx, int level, const double* h_x, double* h_y, double factor) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ensure_stage(ctx, L->ndof)) return 1;
  if (ctx->nranks == 1 && uniform_supported(ctx, *L) && L->n[2] >= 4 * uniform_tile_height(*L) && L->ndof >= (1 << 20))
    return op_apply_host_chunked(ctx, *L, h_x, h_y, factor);
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_x, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  if (op_apply_async(ctx, *L, ctx->d_in, ctx->d_out, factor)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_y, ctx->d_out, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}

int hpdg_jacobi_setup(hpdg_ctx* ctx, int level, int form) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (form == HPDG_JACOBI_DENSE) return jacobi_setup_dense(ctx, *L);
  if (form == HPDG_JACOBI_FD) return jacobi_setup_fd(ctx, *L);
  ctx->err = "unknown block-Jacobi form"; return 1;
}
int hpdg_jacobi_apply_async(hpdg_ctx* ctx, int level, int form, const double* d_r, double* d_c, double damping) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  return jacobi_async(ctx, *L, form, d_r, d_c, damping);
}
int hpdg_jacobi_apply_device(hpdg_ctx* ctx, int level, int form, const double* d_r, double* d_c, double damping) {
  HPDG_ENTER(ctx);
  if (hpdg_jacobi_apply_async(ctx, level, form, d_r, d_c, damping)) return 1;
  return sync_check(ctx);
}
int hpdg_jacobi_apply(hpdg_ctx* ctx, int level, int form, const double* h_r, double* h_c, double damping) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ensure_stage(ctx, L->ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_r, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  if (jacobi_async(ctx, *L, form, ctx->d_in, ctx->d_out, damping)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_c, ctx->d_out, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}
size_t hpdg_jacobi_bytes(const hpdg_ctx* ctx, int level, int form) {
  Level* L = get_level(const_cast<hpdg_ctx*>(ctx), level); if (!L) return 0;
  if (form == HPDG_JACOBI_DENSE) return L->jd.bytes;
  return (size_t)L->jf.nfac * (kMaxN * kMaxN + kMaxN) * sizeof(double) + (size_t)L->nelem * 3 * sizeof(int);
}
int hpdg_diag_block(hpdg_ctx* ctx, int level, long element, double* h_out) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (element < 0 || element >= L->nelem) { ctx->err = "element index out of range"; return 1; }
  int n1 = L->deg[element] + 1, ne = ipow_h(n1, L->dim);
  double* d = nullptr;
  HPDG_CUDA(cudaMalloc(&d, sizeof(double) * ne * ne));
  int rc = diag_block_device(ctx, *L, element, d);
  if (!rc) { cudaError_t e = cudaMemcpy(h_out, d, sizeof(double) * ne * ne, cudaMemcpyDeviceToHost); if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = 1; } }
  cudaFree(d);
  return rc;
}

int hpdg_bcrs_sizes(hpdg_ctx* ctx, int level, long* nblocks, long* nentries) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (L->nc) { ctx->err = "not available on non-conforming meshes (operator apply only)"; return 1; }
  long nb = 0, ne = 0;
  const long stride[3] = {1, L->n[0], (long)L->n[0] * L->n[1]};
  for (long e = 0; e < L->nelem; e++) {
    long r = e; int ijk[3];
    ijk[0] = (int)(r % L->n[0]); r /= L->n[0]; ijk[1] = (int)(r % L->n[1]); r /= L->n[1]; ijk[2] = (int)r;
    const long re = L->off[e + 1] - L->off[e];
    nb++; ne += re * re;
    for (int d = 0; d < L->dim; d++) for (int s = 0; s < 2; s++) {
      const int cc = ijk[d] + (s ? 1 : -1);
      if (cc < 0 || cc >= L->n[d]) continue;
      const long o = e + (s ? stride[d] : -stride[d]);
      nb++; ne += re * (L->off[o + 1] - L->off[o]);
    }
  }
  if (nblocks) *nblocks = nb;
  if (nentries) *nentries = ne;
  return 0;
}
int hpdg_assemble_bcrs(hpdg_ctx* ctx, int level, long* h_rowptr, int* h_col, long* h_boff, double* h_entries) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ctx->nranks > 1) { ctx->err = "assembled export is single-rank"; return 1; }
  if (bcrs_build(ctx, *L)) return 1;
  Bcrs& A = L->bcrs;
  if (h_rowptr) memcpy(h_rowptr, A.rowptr.data(), sizeof(long) * A.rowptr.size());
  if (h_col) memcpy(h_col, A.col.data(), sizeof(int) * A.col.size());
  if (h_boff) memcpy(h_boff, A.boff.data(), sizeof(long) * A.boff.size());
  if (h_entries) HPDG_CUDA(cudaMemcpy(h_entries, A.d_val, sizeof(double) * (size_t)A.boff.back(), cudaMemcpyDeviceToHost));
  return 0;
}
int hpdg_bcrs_mv_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (bcrs_mv(ctx, *L, d_x, d_y)) return 1;
  return sync_check(ctx);
}
int hpdg_bcrs_mv(hpdg_ctx* ctx, int level, const double* h_x, double* h_y) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ensure_stage(ctx, L->ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_x, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  if (bcrs_mv(ctx, *L, ctx->d_in, ctx->d_out)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_y, ctx->d_out, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}
int hpdg_blockgs_iterate_device(hpdg_ctx* ctx, int level, const double* d_b, double* d_x) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (blockgs_iterate(ctx, *L, d_b, d_x)) return 1;
  return sync_check(ctx);
}
int hpdg_blockgs_iterate(hpdg_ctx* ctx, int level, const double* h_b, double* h_x) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ensure_stage(ctx, L->ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_b, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_out, h_x, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  if (blockgs_iterate(ctx, *L, ctx->d_in, ctx->d_out)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_x, ctx->d_out, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}

int hpdg_blockgs_mf_iterate_device(hpdg_ctx* ctx, int level, const double* d_b, double* d_x) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (blockgs_mf_iterate(ctx, *L, d_b, d_x)) return 1;
  return sync_check(ctx);
}
int hpdg_blockgs_mf_iterate(hpdg_ctx* ctx, int level, const double* h_b, double* h_x) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ensure_stage(ctx, L->ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_b, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_out, h_x, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  if (blockgs_mf_iterate(ctx, *L, ctx->d_in, ctx->d_out)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_x, ctx->d_out, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}

int hpdg_l1_setup(hpdg_ctx* ctx, int level, const long* ghosts, long nghost) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (nghost > 0 && !ghosts) { ctx->err = "ghost index list is null"; return 1; }
  return l1_setup(ctx, *L, ghosts, nghost);
}
int hpdg_l1_iterate_device(hpdg_ctx* ctx, int level, const double* d_b, double* d_x) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (blockgs_iterate(ctx, *L, d_b, d_x, 1)) return 1;
  return sync_check(ctx);
}
int hpdg_l1_iterate(hpdg_ctx* ctx, int level, const double* h_b, double* h_x) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ensure_stage(ctx, L->ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_b, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_out, h_x, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  if (blockgs_iterate(ctx, *L, ctx->d_in, ctx->d_out, 1)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_x, ctx->d_out, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}

static int xfer_host(hpdg_ctx* ctx, int fine_level, const double* h_in, double* h_out, bool restrict_) {
  HPDG_ENTER(ctx);
  Level* F = get_level(ctx, fine_level); if (!F) return 1;
  int fl = (int)(F - &ctx->levels[0]);
  if (fl < 1) { ctx->err = "no coarser level below this one"; return 1; }
  Level& C = ctx->levels[fl - 1];
  if (ensure_stage(ctx, F->ndof)) return 1;
  long nin = restrict_ ? F->ndof : C.ndof, nout = restrict_ ? C.ndof : F->ndof;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_in, sizeof(double) * nin, cudaMemcpyHostToDevice, ctx->stream));
  if (restrict_ ? launch_restrict(ctx, *F, C, ctx->d_in, ctx->d_out) : launch_prolong(ctx, *F, C, ctx->d_in, ctx->d_out)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_out, ctx->d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}
int hpdg_restrict(hpdg_ctx* ctx, int fine_level, const double* h_fine, double* h_coarse) { return xfer_host(ctx, fine_level, h_fine, h_coarse, true); }
int hpdg_prolong(hpdg_ctx* ctx, int fine_level, const double* h_coarse, double* h_fine) { return xfer_host(ctx, fine_level, h_coarse, h_fine, false); }
int hpdg_restrict_device(hpdg_ctx* ctx, int fine_level, const double* d_fine, double* d_coarse) {
  HPDG_ENTER(ctx);
  Level* F = get_level(ctx, fine_level); if (!F) return 1;
  int fl = (int)(F - &ctx->levels[0]);
  if (fl < 1) { ctx->err = "no coarser level below this one"; return 1; }
  if (launch_restrict(ctx, *F, ctx->levels[fl - 1], d_fine, d_coarse)) return 1;
  return sync_check(ctx);
}
int hpdg_prolong_device(hpdg_ctx* ctx, int fine_level, const double* d_coarse, double* d_fine) {
  HPDG_ENTER(ctx);
  Level* F = get_level(ctx, fine_level); if (!F) return 1;
  int fl = (int)(F - &ctx->levels[0]);
  if (fl < 1) { ctx->err = "no coarser level below this one"; return 1; }
  if (launch_prolong(ctx, *F, ctx->levels[fl - 1], d_coarse, d_fine)) return 1;
  return sync_check(ctx);
}

int hpdg_vcycle_device(hpdg_ctx* ctx, int form, double damping, int pre, int post, int coarse_its, double* d_x, double* d_b) {
  HPDG_ENTER(ctx);
  VC v = {form, damping, pre, post, coarse_its};
  if (vcycle_device(ctx, v, d_x, d_b)) return 1;
  return sync_check(ctx);
}
int hpdg_vcycle(hpdg_ctx* ctx, int form, double damping, int pre, int post, int coarse_its, double* h_x, double* h_b) {
  HPDG_ENTER(ctx);
  Level& F = ctx->levels.back();
  if (ensure_stage(ctx, F.ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_x, sizeof(double) * F.ndof, cudaMemcpyHostToDevice, ctx->stream));
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_out, h_b, sizeof(double) * F.ndof, cudaMemcpyHostToDevice, ctx->stream));
  VC v = {form, damping, pre, post, coarse_its};
  if (vcycle_device(ctx, v, ctx->d_in, ctx->d_out)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_x, ctx->d_in, sizeof(double) * F.ndof, cudaMemcpyDeviceToHost, ctx->stream));
  HPDG_CUDA(cudaMemcpyAsync(h_b, ctx->d_out, sizeof(double) * F.ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}

// ---- accumulate mode: y += factor * A x -------------------------------------------------------------
// Operator::apply over a TUPLE of local operators zeroes Ax once and lets every local operator add its own factor * (A x)
// (matrix-free/operator.hh:42-55, test: matrix-free/test/testoperator.cc:80-98).  The first operator of a tuple maps to
// hpdg_op_apply*, every further one to hpdg_op_apply_accum*.
int hpdg_op_apply_accum_async(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  ctx->fuse_accum = 1;
  const int rc = op_apply_async(ctx, *L, d_x, d_y, factor);
  ctx->fuse_accum = 0;
  return rc;
}
int hpdg_op_apply_accum_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, double factor) {
  if (hpdg_op_apply_accum_async(ctx, level, d_x, d_y, factor)) return 1;
  return sync_check(ctx);
}
int hpdg_op_apply_accum(hpdg_ctx* ctx, int level, const double* h_x, double* h_y, double factor) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (ensure_stage(ctx, L->ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_x, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_out, h_y, sizeof(double) * L->ndof, cudaMemcpyHostToDevice, ctx->stream));
  ctx->fuse_accum = 1;
  const int rc = op_apply_async(ctx, *L, ctx->d_in, ctx->d_out, factor);
  ctx->fuse_accum = 0;
  if (rc) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_y, ctx->d_out, sizeof(double) * L->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}

// ---- BLAS-1 (DynamicBlockVector: common/dynamicbvector.hh:185-314) -----------------------------------
// this rank's dot product into device scalar `slot`, summed over all ranks of a distributed context (NCCL on the same stream:
// no host round trip)
static int dot_to_slot(hpdg_ctx* ctx, long n, const double* d_x, const double* d_y, int slot) {
  if (blas_scratch(ctx)) return 1;
  if (launch_dot(ctx, n, d_x, d_y, ctx->d_scalar + slot)) return 1;
  if (ctx->nranks > 1 && !ctx->nccl) { ctx->err = "dot product over ranks needs the NCCL communicator (context created with nccl_id = NULL)"; return 1; }
  if (ctx->nranks > 1)
    HPDG_NCCL(g_nccl.AllReduce(ctx->d_scalar + slot, ctx->d_scalar + slot, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  return 0;
}
int hpdg_dot_device(hpdg_ctx* ctx, int level, const double* d_x, const double* d_y, double* h_result) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  if (dot_to_slot(ctx, L->ndof, d_x, d_y, 15)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_result, ctx->d_scalar + 15, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}
int hpdg_two_norm_device(hpdg_ctx* ctx, int level, const double* d_x, double* h_result) {
  if (hpdg_dot_device(ctx, level, d_x, d_x, h_result)) return 1;
  *h_result = std::sqrt(*h_result);
  return 0;
}
int hpdg_axpy_device(hpdg_ctx* ctx, int level, double a, const double* d_x, double* d_y) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  return launch_axpy(ctx, L->ndof, a, d_x, d_y);
}
int hpdg_scale_device(hpdg_ctx* ctx, int level, double a, double* d_x) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  return launch_scale(ctx, L->ndof, a, d_x);
}
int hpdg_assign_device(hpdg_ctx* ctx, int level, const double* d_src, double* d_dst) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  HPDG_CUDA(cudaMemcpyAsync(d_dst, d_src, sizeof(double) * L->ndof, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

// ---- solver loops around the hot path ------------------------------------------------------------------
// Preconditioned CG on the finest level, vectors and scalars resident on the device: per iteration one operator apply, one
// preconditioner application, three dot products (each ending in a 1-double ncclAllReduce on the context stream when the
// context is distributed) and two fused vector updates whose step lengths are read from device memory -- nothing returns
// to the host except, every `check_every` iterations, the residual norm.
//   precond: HPDG_PRECOND_NONE, HPDG_PRECOND_JACOBI (fd block Jacobi, `damping`), HPDG_PRECOND_VCYCLE (one V-cycle with fd
//   block-Jacobi smoothing: `damping`, pre = post = `smooth`, `coarse_its` coarse iterations).
// Stops when ||r||_2 <= tol * ||r_0||_2 or after maxit iterations.
static int pcg_device(hpdg_ctx* ctx, int precond, double damping, int smooth, int coarse_its, double* d_x, const double* d_b,
                      double tol, int maxit, int check_every, int* iters, double* relres) {
  Level& F = ctx->levels.back();
  const long n = F.ndof;
  if (blas_scratch(ctx)) return 1;
  if (!ctx->cg_p) {
    HPDG_CUDA(cudaMalloc(&ctx->cg_p, sizeof(double) * n)); HPDG_CUDA(cudaMalloc(&ctx->cg_q, sizeof(double) * n));
    HPDG_CUDA(cudaMalloc(&ctx->cg_r, sizeof(double) * n)); HPDG_CUDA(cudaMalloc(&ctx->cg_z, sizeof(double) * n));
  }
  double *p = ctx->cg_p, *q = ctx->cg_q, *r = ctx->cg_r, *z = ctx->cg_z;
  if (precond == HPDG_PRECOND_JACOBI && !F.jf.ready) { if (jacobi_setup_fd(ctx, F)) return 1; }
  if (check_every < 1) check_every = 1;
  const VC vc = {HPDG_JACOBI_FD, damping, smooth, smooth, coarse_its};
  auto apply_precond = [&](const double* rin, double* zout) -> int {
    if (precond == HPDG_PRECOND_NONE) { HPDG_CUDA(cudaMemcpyAsync(zout, rin, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream)); return 0; }
    if (precond == HPDG_PRECOND_JACOBI) return jacobi_async(ctx, F, HPDG_JACOBI_FD, rin, zout, damping);
    // V-cycle: z = 0; z += MG(r); the cycle overwrites its right-hand side with the residual, so it works on a copy (q is free here)
    HPDG_CUDA(cudaMemsetAsync(zout, 0, sizeof(double) * n, ctx->stream));
    HPDG_CUDA(cudaMemcpyAsync(q, rin, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    return vcycle_device(ctx, vc, zout, q);
  };
  enum { RZ0 = 0, RZ1 = 1, PQ = 2, RR = 3 };
  double rr0 = 0, rr = 0;
  if (op_apply_async(ctx, F, d_x, q, 1.0)) return 1;
  if (launch_xpay_sub(ctx, n, d_b, q, r)) return 1;                       // r = b - A x
  if (dot_to_slot(ctx, n, r, r, RR)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(&rr0, ctx->d_scalar + RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (sync_check(ctx)) return 1;
  rr = rr0;
  int it = 0;
  if (rr0 > 0) {
    if (apply_precond(r, z)) return 1;
    HPDG_CUDA(cudaMemcpyAsync(p, z, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (dot_to_slot(ctx, n, r, z, RZ0)) return 1;
    for (it = 1; it <= maxit; it++) {
      const int cur = (it - 1) & 1, nxt = it & 1;                         // rz of this / the next iteration
      if (op_apply_async(ctx, F, p, q, 1.0)) return 1;                    // q = A p
      if (dot_to_slot(ctx, n, p, q, PQ)) return 1;
      // x += (rz/pq) p ; r -= (rz/pq) q ; rr = r . r in the same pass, summed over the ranks
      if (launch_cg_update_rr(ctx, n, cur, PQ, p, q, d_x, r, ctx->d_scalar + RR)) return 1;
      if (ctx->nranks > 1) {
        if (!ctx->nccl) { ctx->err = "dot product over ranks needs the NCCL communicator (context created with nccl_id = NULL)"; return 1; }
        HPDG_NCCL(g_nccl.AllReduce(ctx->d_scalar + RR, ctx->d_scalar + RR, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
      }
      if (it % check_every == 0 || it == maxit) {
        HPDG_CUDA(cudaMemcpyAsync(&rr, ctx->d_scalar + RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (sync_check(ctx)) return 1;
        if (!(rr == rr)) { ctx->err = "pcg: residual is not a number (operator or preconditioner not SPD?)"; return 1; }
        if (std::sqrt(rr) <= tol * std::sqrt(rr0)) break;
      }
      if (apply_precond(r, z)) return 1;
      if (dot_to_slot(ctx, n, r, z, nxt)) return 1;
      if (launch_cg_direction(ctx, n, nxt, cur, z, p)) return 1;          // p = z + (rz_new/rz) p
    }
    if (it > maxit) it = maxit;
  }
  if (iters) *iters = it;
  if (relres) *relres = rr0 > 0 ? std::sqrt(rr / rr0) : 0.0;
  return sync_check(ctx);
}
int hpdg_pcg_device(hpdg_ctx* ctx, int precond, double damping, int smooth, int coarse_its, double* d_x, const double* d_b,
                    double tol, int maxit, int check_every, int* iters, double* relres) {
  HPDG_ENTER(ctx);
  if (precond < HPDG_PRECOND_NONE || precond > HPDG_PRECOND_VCYCLE) { ctx->err = "unknown preconditioner"; return 1; }
  return pcg_device(ctx, precond, damping, smooth, coarse_its, d_x, d_b, tol, maxit, check_every, iters, relres);
}
int hpdg_pcg(hpdg_ctx* ctx, int precond, double damping, int smooth, int coarse_its, double* h_x, const double* h_b, double tol,
             int maxit, int check_every, int* iters, double* relres) {
  HPDG_ENTER(ctx);
  Level& F = ctx->levels.back();
  if (ensure_stage(ctx, F.ndof)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_in, h_x, sizeof(double) * F.ndof, cudaMemcpyHostToDevice, ctx->stream));
  HPDG_CUDA(cudaMemcpyAsync(ctx->d_out, h_b, sizeof(double) * F.ndof, cudaMemcpyHostToDevice, ctx->stream));
  if (hpdg_pcg_device(ctx, precond, damping, smooth, coarse_its, ctx->d_in, ctx->d_out, tol, maxit, check_every, iters, relres)) return 1;
  HPDG_CUDA(cudaMemcpyAsync(h_x, ctx->d_in, sizeof(double) * F.ndof, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_check(ctx);
}

// Dune::Solvers::LoopSolver around the multigrid step with the energy norm as the reference wires it (buildingblocks/solve.hh:150-166,
// iterationsteps/mg/mgwrapper.hh:22-27): per iteration the step works on a COPY of the right-hand side (the cycle overwrites it),
// the error measure is the energy norm of the correction ||x_k - x_{k-1}||_A divided by ||x_{k-1}||_A (useRelativeError = true;
// dune-solvers' LoopSolver is un-vendored, this is its documented behaviour: with a zero old iterate the quotient is infinite and the
// loop goes on); the loop stops when the measure drops below tol or after maxit iterations.  Each energy norm costs one operator
// apply + one dot product.
int hpdg_loop_solve_device(hpdg_ctx* ctx, int form, double damping, int pre, int post, int coarse_its, double* d_x,
                           const double* d_b, double tol, int maxit, int* iters, double* last_error) {
  HPDG_ENTER(ctx);
  Level& F = ctx->levels.back();
  const long n = F.ndof;
  if (blas_scratch(ctx)) return 1;
  if (!ctx->cg_p) {
    HPDG_CUDA(cudaMalloc(&ctx->cg_p, sizeof(double) * n)); HPDG_CUDA(cudaMalloc(&ctx->cg_q, sizeof(double) * n));
    HPDG_CUDA(cudaMalloc(&ctx->cg_r, sizeof(double) * n)); HPDG_CUDA(cudaMalloc(&ctx->cg_z, sizeof(double) * n));
  }
  double *old = ctx->cg_p, *rhs = ctx->cg_q, *corr = ctx->cg_r, *acorr = ctx->cg_z;
  const VC vc = {form, damping, pre, post, coarse_its};
  int it = 0;
  double err = 0;
  for (it = 1; it <= maxit; it++) {
    HPDG_CUDA(cudaMemcpyAsync(old, d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    HPDG_CUDA(cudaMemcpyAsync(rhs, d_b, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));   // mgwrapper.hh:23
    if (vcycle_device(ctx, vc, d_x, rhs)) return 1;
    if (launch_xpay_sub(ctx, n, d_x, old, corr)) return 1;                                               // correction
    if (op_apply_async(ctx, F, corr, acorr, 1.0)) return 1;
    if (dot_to_slot(ctx, n, corr, acorr, 4)) return 1;                                                   // ||c||_A^2
    if (op_apply_async(ctx, F, old, acorr, 1.0)) return 1;
    if (dot_to_slot(ctx, n, old, acorr, 5)) return 1;                                                    // ||x_old||_A^2
    double h[2] = {0, 0};
    HPDG_CUDA(cudaMemcpyAsync(h, ctx->d_scalar + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sync_check(ctx)) return 1;
    const double nc = std::sqrt(std::max(h[0], 0.0)), no = std::sqrt(std::max(h[1], 0.0));
    err = no > 0 ? nc / no : (nc > 0 ? INFINITY : 0.0);
    if (err < tol) break;
  }
  if (it > maxit) it = maxit;
  if (iters) *iters = it;
  if (last_error) *last_error = err;
  return 0;
}

// host-only introspection (no device work): the 1-D tables of degree p the kernels are built from, n = p + 1, row-major n x n
int hpdg_tables_1d(int degree, double* nodes, double* mass, double* stiffness, double* end_values, double* end_derivatives) {
  if (degree < 0 || degree > kMaxP) { g_create_err = "polynomial degree out of range 0..13"; return 1; }
  const DegTable& T = host_tables().deg[degree];
  const int n = degree + 1;
  for (int i = 0; i < n; i++) {
    if (nodes) nodes[i] = T.nodes[i];
    for (int j = 0; j < n; j++) {
      if (mass) mass[i * n + j] = T.M[i * kMaxN + j];
      if (stiffness) stiffness[i * n + j] = T.S[i * kMaxN + j];
    }
    for (int s = 0; s < 2; s++) {
      if (end_values) end_values[s * n + i] = T.t[s][i];
      if (end_derivatives) end_derivatives[s * n + i] = T.g[s][i];
    }
  }
  return 0;
}

// host-only introspection: the tangential couplings of the mixed-degree and non-conforming faces (tables.hpp).  kind 0: conforming
// face, (M^ee)^-1 M^eo; 1 / 2: e is the coarse side, low / high half of its side; 3 / 4: e is the fine side on the low / high half of the
// neighbour's side.  out is (pe+1) x (po+1) row-major; own (kinds 1, 2 only, may be NULL) is the (pe+1) x (pe+1) own-side coupling.
int hpdg_tables_face(int pe, int po, int kind, double* out, double* own) {
  if (pe < 0 || pe > kMaxP || po < 0 || po > kMaxP || kind < 0 || kind > 4) { g_create_err = "degree or kind out of range"; return 1; }
  const HostTables& H = host_tables();
  const int ND = kMaxP + 1, ne = pe + 1, no = po + 1;
  const double* P = kind == 0 ? &H.P[((size_t)pe * ND + po) * kMaxN * kMaxN]
                              : &H.Pnc_eo[(((size_t)(kind - 1) * ND + pe) * ND + po) * kMaxN * kMaxN];
  for (int i = 0; i < ne; i++) for (int j = 0; j < no; j++) out[i * no + j] = P[i * kMaxN + j];
  if (own) {
    if (kind != 1 && kind != 2) { g_create_err = "the own-side coupling exists for the coarse side (kinds 1, 2) only"; return 1; }
    const double* Q = &H.Pnc_ee[((size_t)(kind - 1) * ND + pe) * kMaxN * kMaxN];
    for (int i = 0; i < ne; i++) for (int j = 0; j < ne; j++) own[i * ne + j] = Q[i * kMaxN + j];
  }
  return 0;
}

long hpdg_launch_count(const hpdg_ctx* ctx) { return ctx->launches; }
int hpdg_uses_uniform_kernel(const hpdg_ctx* ctx, int level) {
  Level* L = get_level(const_cast<hpdg_ctx*>(ctx), level);
  return L ? uniform_supported(ctx, *L) : 0;
}
int hpdg_time_apply_device(hpdg_ctx* ctx, int level, const double* d_x, double* d_y, int reps, float* ms_per_apply) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  return time_launches(ctx, reps, ms_per_apply, [&]() { return op_apply_async(ctx, *L, d_x, d_y, 1.0); });
}
int hpdg_time_jacobi_device(hpdg_ctx* ctx, int level, int form, const double* d_r, double* d_c, double damping, int reps,
                            float* ms_per_apply) {
  HPDG_ENTER(ctx);
  Level* L = get_level(ctx, level); if (!L) return 1;
  return time_launches(ctx, reps, ms_per_apply, [&]() { return jacobi_async(ctx, *L, form, d_r, d_c, damping); });
}

}  // extern "C"
