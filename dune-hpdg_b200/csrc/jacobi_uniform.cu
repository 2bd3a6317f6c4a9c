// Block-Jacobi in fast-diagonalisation form for uniform degree on a 3-D brick: the tiled counterpart of k_jacobi_fd.
//
//   c_e = damping * (Vx x Vy x Vz) diag(1/(lx_i + ly_j + lz_k)) (Vx x Vy x Vz)^T r_e          (see jacobi.cu)
//
// Replaces IPDGBlockJacobi driven by Operator::apply (matrix-free/localoperators/ipdgblockjacobi.hh:58-178) with an exact
// local solver.  On a uniform-degree brick the 1-D factor of direction d depends only on whether the element is the
// first and/or last of its grid line (Dirichlet / natural boundary face instead of an interior face), so there are at
// most 4 variants per direction; they travel in the kernel parameter block.  Same five-pass pencil structure as the
// operator kernel (apply_uniform.cu) but without any neighbour coupling:
//   P1 z-pencils: global r -> Vz^T      P2 x-pencils: Vx^T      P3 y-pencils: Vy^T, scale by 1/(sum of eigenvalues), Vy
//   P4 x-pencils: Vx                     P5 z-pencils: damping * Vz -> global c
// Algorithmic HBM traffic: 16 B/DoF.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "jacobi_uniform_q3p.cuh"
#include "jacobi_uniform_q4p.cuh"

namespace hpdg {

template <int N>
struct FDUniParams {
  double V[3][4][N * N];  // [direction][variant: bit0 first-in-line at a domain boundary, bit1 last][row-major, eigenvectors in columns]
  double lam[3][4][N];
  double damping;
  int n[3];
  int ntile[3];
  int bnd[6];  // 1 if brick face f is a domain boundary (as opposed to a rank boundary)
  const double* r;
  double* c;
  double* xacc;  // optional: x += c
};

template <int N> struct PitchJ {
  static constexpr int PP = (N % 2 == 0) ? N * N + 1 : N * N;
  static constexpr int EP0 = N * PP;
  static constexpr int EP = EP0 + ((N - EP0 % 16) % 16 + 16) % 16;
};

// out = V^T a (TRANS) or V a, V = P.V[D][VAR]
template <int N, int D, int VAR, bool TRANS>
__device__ __forceinline__ void fd_line(const FDUniParams<N>& P, double (&a)[N]) {
  double o[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < N; m++) s = fma(TRANS ? P.V[D][VAR][m * N + i] : P.V[D][VAR][i * N + m], a[m], s);
    o[i] = s;
  }
#pragma unroll
  for (int i = 0; i < N; i++) a[i] = o[i];
}
template <int N, int D, bool TRANS>
__device__ __forceinline__ void fd_line_v(const FDUniParams<N>& P, int var, double (&a)[N]) {
  switch (var) {  // uniform across the warp: the variant depends only on the element's position along the pencil
    case 0: fd_line<N, D, 0, TRANS>(P, a); break;
    case 1: fd_line<N, D, 1, TRANS>(P, a); break;
    case 2: fd_line<N, D, 2, TRANS>(P, a); break;
    default: fd_line<N, D, 3, TRANS>(P, a); break;
  }
}

template <int N, int TX, int TY, int TZ>
constexpr int fdu_threads() {
  int a = N * N * TY * TZ, b = N * N * TX * TZ, c = N * N * TX * TY;
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

template <int N, int TX, int TY, int TZ, int MINB>
__global__ void __launch_bounds__(fdu_threads<N, TX, TY, TZ>(), MINB)
k_jacobi_fd_uniform(const __grid_constant__ FDUniParams<N> P) {
  constexpr int PP = PitchJ<N>::PP, EP = PitchJ<N>::EP;
  constexpr int N2 = N * N, N3 = N * N * N;
  extern __shared__ double sw[];
  int tb = blockIdx.x;
  const int tx = tb % P.ntile[0]; tb /= P.ntile[0];
  const int ty = tb % P.ntile[1]; const int tz = tb / P.ntile[1];
  const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
  const int lenx = min(TX, P.n[0] - x0), leny = min(TY, P.n[1] - y0), lenz = min(TZ, P.n[2] - z0);
  const long sx = N3, sy = (long)P.n[0] * N3, sz = (long)P.n[0] * P.n[1] * N3;
  const int tid = threadIdx.x;
  auto variant = [&](int d, int pos) {
    return ((pos == 0 && P.bnd[2 * d]) ? 1 : 0) | ((pos == P.n[d] - 1 && P.bnd[2 * d + 1]) ? 2 : 0);
  };
  const int zi = tid % N, zj = (tid / N) % N, zex = (tid / N2) % TX, zey = tid / (N2 * TX);
  const bool zact = (tid < N2 * TX * TY) && zex < lenx && zey < leny;
  const int xj = tid % N, xk = (tid / N) % N, xey = (tid / N2) % TY, xez = tid / (N2 * TY);
  const bool xact = (tid < N2 * TY * TZ) && xey < leny && xez < lenz;
  const int yi = tid % N, yex = (tid / N) % TX, yk = (tid / (N * TX)) % N, yez = tid / (N2 * TX);
  const bool yact = (tid < N2 * TX * TZ) && yex < lenx && yez < lenz;
  const int znode = zi + N * zj;
  const long zcol = (long)(x0 + zex) * sx + (long)(y0 + zey) * sy + znode;
  const int zbase = (zex + TX * zey) * EP + znode;
  const int xbase = TX * (xey + TY * xez) * EP + N * xj + PP * xk;
  const int ybase = (yex + TX * TY * yez) * EP + yi + PP * yk;

  if (zact) {  // P1
    double a[TZ][N];
#pragma unroll
    for (int e = 0; e < TZ; e++)
#pragma unroll
      for (int k = 0; k < N; k++) a[e][k] = (e < lenz) ? __ldg(P.r + zcol + (long)(z0 + e) * sz + N2 * k) : 0.0;
#pragma unroll
    for (int e = 0; e < TZ; e++)
      if (e < lenz) {
        fd_line_v<N, 2, true>(P, variant(2, z0 + e), a[e]);
#pragma unroll
        for (int k = 0; k < N; k++) sw[zbase + TX * TY * EP * e + PP * k] = a[e][k];
      }
  }
  __syncthreads();
  if (xact) {  // P2
#pragma unroll
    for (int e = 0; e < TX; e++)
      if (e < lenx) {
        double a[N];
#pragma unroll
        for (int i = 0; i < N; i++) a[i] = sw[xbase + e * EP + i];
        fd_line_v<N, 0, true>(P, variant(0, x0 + e), a);
#pragma unroll
        for (int i = 0; i < N; i++) sw[xbase + e * EP + i] = a[i];
      }
  }
  __syncthreads();
  if (yact) {  // P3
    const int vx = variant(0, x0 + yex), vz = variant(2, z0 + yez);
    const double lxz = P.lam[0][vx][yi] + P.lam[2][vz][yk];
#pragma unroll
    for (int e = 0; e < TY; e++)
      if (e < leny) {
        double a[N];
#pragma unroll
        for (int j = 0; j < N; j++) a[j] = sw[ybase + TX * EP * e + N * j];
        const int vy = variant(1, y0 + e);
        fd_line_v<N, 1, true>(P, vy, a);
#pragma unroll
        for (int j = 0; j < N; j++) a[j] = a[j] / (lxz + P.lam[1][vy][j]);
        fd_line_v<N, 1, false>(P, vy, a);
#pragma unroll
        for (int j = 0; j < N; j++) sw[ybase + TX * EP * e + N * j] = a[j];
      }
  }
  __syncthreads();
  if (xact) {  // P4
#pragma unroll
    for (int e = 0; e < TX; e++)
      if (e < lenx) {
        double a[N];
#pragma unroll
        for (int i = 0; i < N; i++) a[i] = sw[xbase + e * EP + i];
        fd_line_v<N, 0, false>(P, variant(0, x0 + e), a);
#pragma unroll
        for (int i = 0; i < N; i++) sw[xbase + e * EP + i] = a[i];
      }
  }
  __syncthreads();
  if (zact) {  // P5
#pragma unroll
    for (int e = 0; e < TZ; e++)
      if (e < lenz) {
        double a[N];
#pragma unroll
        for (int k = 0; k < N; k++) a[k] = sw[zbase + TX * TY * EP * e + PP * k];
        fd_line_v<N, 2, false>(P, variant(2, z0 + e), a);
        double* co = P.c + zcol + (long)(z0 + e) * sz;
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double cv = P.damping * a[k];
          co[N2 * k] = cv;
          if (P.xacc) { double* xo = P.xacc + zcol + (long)(z0 + e) * sz; xo[N2 * k] += cv; }
        }
      }
  }
}

template <int N, int TX, int TY, int TZ, int MINB>
static int launch_fdu(Ctx* ctx, Level& L, const double* r, double* c, double damping) {
  static thread_local FDUniParams<N> P;
  const DegTable& T = host_tables().deg[N - 1];
  const double cpen = ctx->sigma * (double)L.pen_uni * L.pen_uni;
  for (int d = 0; d < 3; d++) {
    double kappa = 1.0 / L.h[d];
    for (int dd = 0; dd < 3; dd++) if (dd != d) kappa *= L.h[dd];
    for (int s = 0; s < 2; s++) P.bnd[2 * d + s] = ctx->bnd_is_rank[2 * d + s] ? 0 : 1;
    for (int var = 0; var < 4; var++) {
      double w[2], cc[2];
      for (int s = 0; s < 2; s++) {
        const bool at_bnd = (var >> s) & 1;
        if (!at_bnd) { w[s] = 0.5; cc[s] = cpen; }                 // interior face (ipdgblockjacobi.hh:69,80-86)
        else if (ctx->dirichlet) { w[s] = 1.0; cc[s] = cpen; }     // Dirichlet face (:74)
        else { w[s] = 0.0; cc[s] = 0.0; }                          // natural boundary (:72-73)
      }
      double D[N * N], M[N * N], V[N * N], lam[N];
      for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
        double v = kappa * T.S[i * kMaxN + j];
        for (int s = 0; s < 2; s++) {
          const double nu = s ? 1.0 : -1.0;
          v += -w[s] * nu * kappa * (T.t[s][i] * T.g[s][j] + T.g[s][i] * T.t[s][j]) + cc[s] * T.t[s][i] * T.t[s][j];
        }
        D[i * N + j] = v; M[i * N + j] = T.M[i * kMaxN + j];
      }
      gen_eig(N, D, M, V, lam);
      for (int i = 0; i < N * N; i++) P.V[d][var][i] = V[i];
      for (int i = 0; i < N; i++) P.lam[d][var][i] = lam[i];
    }
  }
  const int tdim[3] = {TX, TY, TZ};
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.ntile[d] = (L.n[d] + tdim[d] - 1) / tdim[d]; }
  P.damping = damping; P.r = r; P.c = c; P.xacc = ctx->fuse_xacc;
  constexpr int threads = fdu_threads<N, TX, TY, TZ>();
  constexpr size_t smem = sizeof(double) * TX * TY * TZ * PitchJ<N>::EP;
  if (kernel_slots(ctx, reinterpret_cast<const void*>(k_jacobi_fd_uniform<N, TX, TY, TZ, MINB>), threads, smem, nullptr)) return 1;
  const long ntiles = (long)P.ntile[0] * P.ntile[1] * P.ntile[2];
  k_jacobi_fd_uniform<N, TX, TY, TZ, MINB><<<(unsigned)ntiles, threads, smem, ctx->stream>>>(P);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

// 1-D generalised eigen-decomposition of the diagonal block's factor in direction d for a boundary variant
// (bit 0: first element of its line at a domain boundary, bit 1: last); see launch_fdu
static void fd_factor(const Ctx* ctx, const Level& L, int N, int d, int var, double* V, double* lam) {
  const DegTable& T = host_tables().deg[N - 1];
  const double cpen = ctx->sigma * (double)L.pen_uni * L.pen_uni;
  double kappa = 1.0 / L.h[d];
  for (int dd = 0; dd < 3; dd++) if (dd != d) kappa *= L.h[dd];
  double w[2], cc[2];
  for (int s = 0; s < 2; s++) {
    const bool at_bnd = (var >> s) & 1;
    if (!at_bnd) { w[s] = 0.5; cc[s] = cpen; }                 // interior face (ipdgblockjacobi.hh:69,80-86)
    else if (ctx->dirichlet) { w[s] = 1.0; cc[s] = cpen; }     // Dirichlet face (:74)
    else { w[s] = 0.0; cc[s] = 0.0; }                          // natural boundary (:72-73)
  }
  std::vector<double> D((size_t)N * N), M((size_t)N * N);
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
    double v = kappa * T.S[i * kMaxN + j];
    for (int s = 0; s < 2; s++) {
      const double nu = s ? 1.0 : -1.0;
      v += -w[s] * nu * kappa * (T.t[s][i] * T.g[s][j] + T.g[s][i] * T.t[s][j]) + cc[s] * T.t[s][i] * T.t[s][j];
    }
    D[i * N + j] = v; M[i * N + j] = T.M[i * kMaxN + j];
  }
  gen_eig(N, D.data(), M.data(), V, lam);
}

// persistent Q3 kernel (jacobi_uniform_q3p.cuh): same conditions as the persistent operator kernel.
// Returns -1 if the interior factor does not have the mirror structure the kernel relies on (then the tile kernel runs).
static int launch_q3j(Ctx* ctx, Level& L, const double* r, double* c, double damping) {
  static thread_local Q3jParams P;
  if (L.q3j_state < 0) return -1;
  double (&lam)[3][3][4] = L.q3j_lam;
  if (L.q3j_state == 0) {  // once per level: the 1-D factors (the context's sigma / boundary type are fixed at creation)
    L.q3j_state = -1;
    for (int d = 0; d < 3; d++)
      for (int var = 0; var < 3; var++) fd_factor(ctx, L, 4, d, var, L.q3j_V[d][var], lam[d][var]);
    // interior factor: reorder the eigenpairs to even, odd, even, odd under the node reflection i -> 3 - i
    for (int d = 0; d < 3; d++) {
      double* V = L.q3j_V[d][0];
      int even[4], odd[4], ne = 0, no = 0;
      for (int k = 0; k < 4; k++) {
        double se = 0, so = 0, nn = 0;
        for (int i = 0; i < 4; i++) {
          se += std::fabs(V[i * 4 + k] - V[(3 - i) * 4 + k]); so += std::fabs(V[i * 4 + k] + V[(3 - i) * 4 + k]); nn += std::fabs(V[i * 4 + k]);
        }
        if (se <= 1e-13 * nn) even[ne++] = k; else if (so <= 1e-13 * nn) odd[no++] = k; else return -1;
      }
      if (ne != 2 || no != 2) return -1;
      const int order[4] = {even[0], odd[0], even[1], odd[1]};
      double Vn[16], ln[4];
      for (int k = 0; k < 4; k++) { ln[k] = lam[d][0][order[k]]; for (int i = 0; i < 4; i++) Vn[i * 4 + k] = V[i * 4 + order[k]]; }
      for (int k = 0; k < 4; k++) { lam[d][0][k] = ln[k]; for (int i = 0; i < 4; i++) V[i * 4 + k] = Vn[i * 4 + k]; }
    }
    L.q3j_state = 1;
  }
  std::memcpy(P.V, L.q3j_V, sizeof(P.V));
  if (q3p_level_setup(ctx, L)) return 1;
  if (!L.d_jinv || L.jinv_damping != damping) {
    std::vector<double> inv(27 * 64);
    for (int vx = 0; vx < 3; vx++) for (int vy = 0; vy < 3; vy++) for (int vz = 0; vz < 3; vz++)
      for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) for (int i = 0; i < 4; i++)
        inv[(size_t)((vx * 3 + vy) * 3 + vz) * 64 + (j * 4 + k) * 4 + i] = damping / (lam[0][vx][i] + lam[1][vy][j] + lam[2][vz][k]);
    if (!L.d_jinv) HPDG_CUDA(cudaMalloc(&L.d_jinv, sizeof(double) * inv.size()));
    HPDG_CUDA(cudaMemcpyAsync(L.d_jinv, inv.data(), sizeof(double) * inv.size(), cudaMemcpyHostToDevice, ctx->stream));
    HPDG_CUDA(cudaStreamSynchronize(ctx->stream));  // inv is a local
    L.jinv_damping = damping;
  }
  for (int f = 0; f < 6; f++) P.bnd[f] = ctx->bnd_is_rank[f] ? 0 : 1;
  for (int d = 0; d < 3; d++) P.n[d] = L.n[d];
  P.r = r; P.c = c; P.xacc = ctx->fuse_xacc; P.inv = L.d_jinv;
  P.tile_desc = static_cast<const int4*>(L.d_tile_desc);
  P.sched = ctx->d_sched + 10;
  P.ntiles = (L.n[0] / 4) * (L.n[1] / 4) * (L.n[2] / 4);
  int slots = 0;
  if (kernel_slots(ctx, reinterpret_cast<const void*>(hpdg_k_jacobi_fd_q3_persist), 256, kQ3jSmemBytes, &slots)) return 1;
  const int grid = std::min(P.ntiles, ctx->q3p_grid > 0 ? ctx->q3p_grid : slots);
  hpdg_k_jacobi_fd_q3_persist<<<grid, 256, kQ3jSmemBytes, ctx->stream>>>(P);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

// persistent Q4 kernel (jacobi_uniform_q4p.cuh): uniform Q4 bricks with extents multiple of (4, 4, 2); option "variant" = 40
// switches it off.  Returns -1 when it does not apply.
static int launch_q4j(Ctx* ctx, Level& L, const double* r, double* c, double damping) {
  static thread_local Q4jParams P;
  if (L.q3j_state < 0) return -1;
  double (&lam)[3][3][5] = L.q4j_lam;
  if (L.q3j_state == 0) {  // once per level
    L.q3j_state = -1;
    for (int d = 0; d < 3; d++)
      for (int var = 0; var < 3; var++) fd_factor(ctx, L, 5, d, var, L.q4j_V[d][var], lam[d][var]);
    for (int d = 0; d < 3; d++) {  // interior factor: eigenpairs in the order even, odd, even, odd, even under i -> 4 - i
      double* V = L.q4j_V[d][0];
      int even[5], odd[5], ne = 0, no = 0;
      for (int k = 0; k < 5; k++) {
        double se = 0, so = 0, nn = 0;
        for (int i = 0; i < 5; i++) {
          se += std::fabs(V[i * 5 + k] - V[(4 - i) * 5 + k]); so += std::fabs(V[i * 5 + k] + V[(4 - i) * 5 + k]); nn += std::fabs(V[i * 5 + k]);
        }
        if (se <= 1e-13 * nn) even[ne++] = k; else if (so <= 1e-13 * nn) odd[no++] = k; else return -1;
      }
      if (ne != 3 || no != 2) return -1;
      const int order[5] = {even[0], odd[0], even[1], odd[1], even[2]};
      double Vn[25], ln[5];
      for (int k = 0; k < 5; k++) { ln[k] = lam[d][0][order[k]]; for (int i = 0; i < 5; i++) Vn[i * 5 + k] = V[i * 5 + order[k]]; }
      for (int k = 0; k < 5; k++) { lam[d][0][k] = ln[k]; for (int i = 0; i < 5; i++) V[i * 5 + k] = Vn[i * 5 + k]; }
    }
    // tile descriptors of the 4x4x2 tiles + the scheduler counters
    if (q3p_level_setup(ctx, L, 2)) return 1;
    L.q3j_state = 1;
  }
  std::memcpy(P.V, L.q4j_V, sizeof(P.V));
  if (!L.d_jinv || L.jinv_damping != damping) {
    std::vector<double> inv(27 * 125);
    for (int vx = 0; vx < 3; vx++) for (int vy = 0; vy < 3; vy++) for (int vz = 0; vz < 3; vz++)
      for (int k = 0; k < 5; k++) for (int j = 0; j < 5; j++) for (int i = 0; i < 5; i++)
        inv[(size_t)((vx * 3 + vy) * 3 + vz) * 125 + (k * 5 + j) * 5 + i] = damping / (lam[0][vx][i] + lam[1][vy][j] + lam[2][vz][k]);
    if (!L.d_jinv) HPDG_CUDA(cudaMalloc(&L.d_jinv, sizeof(double) * inv.size()));
    HPDG_CUDA(cudaMemcpyAsync(L.d_jinv, inv.data(), sizeof(double) * inv.size(), cudaMemcpyHostToDevice, ctx->stream));
    HPDG_CUDA(cudaStreamSynchronize(ctx->stream));  // inv is a local
    L.jinv_damping = damping;
  }
  for (int f = 0; f < 6; f++) P.bnd[f] = ctx->bnd_is_rank[f] ? 0 : 1;
  for (int d = 0; d < 3; d++) P.n[d] = L.n[d];
  P.r = r; P.c = c; P.xacc = ctx->fuse_xacc; P.inv = L.d_jinv;
  P.tile_desc = static_cast<const int4*>(L.d_tile_desc);
  P.sched = ctx->d_sched + 10;
  P.ntiles = (L.n[0] / 4) * (L.n[1] / 4) * (L.n[2] / 2);
  int slots = 0;
  if (kernel_slots(ctx, reinterpret_cast<const void*>(hpdg_k_jacobi_fd_q4_persist), 160, kQ4jSmemBytes, &slots)) return 1;
  const int grid = std::min(P.ntiles, ctx->q3p_grid > 0 ? ctx->q3p_grid : slots);
  hpdg_k_jacobi_fd_q4_persist<<<grid, 160, kQ4jSmemBytes, ctx->stream>>>(P);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

// returns -1 when there is no specialised kernel for this level
int jacobi_apply_fd_uniform(Ctx* ctx, Level& L, const double* r, double* c, double damping) {
  if (!uniform_supported(ctx, L)) return -1;
  // the bulk copies / bulk stores of the persistent kernel need 16-byte aligned vectors
  const bool xacc_ok = (reinterpret_cast<uintptr_t>(ctx->fuse_xacc) & 15) == 0;   // bulk reduce-add of x += c
  if (uniform_persistent(ctx, L, r) && (reinterpret_cast<uintptr_t>(c) & 15) == 0 && xacc_ok) {
    const int rc = launch_q3j(ctx, L, r, c, damping);
    if (rc >= 0) return rc;
  }
  if (ctx->variant != 40 && L.p_uni == 4 && L.n[0] % 4 == 0 && L.n[1] % 4 == 0 && L.n[2] % 2 == 0 && L.ndof < (1L << 31) &&
      (reinterpret_cast<uintptr_t>(r) & 15) == 0 && (reinterpret_cast<uintptr_t>(c) & 15) == 0 && xacc_ok) {
    const int rc = launch_q4j(ctx, L, r, c, damping);
    if (rc >= 0) return rc;
  }
  switch (L.p_uni) {
    case 1: return launch_fdu<2, 4, 4, 4, 4>(ctx, L, r, c, damping);
    case 2: return launch_fdu<3, 4, 4, 4, 4>(ctx, L, r, c, damping);
    case 3: return launch_fdu<4, 4, 4, 4, 4>(ctx, L, r, c, damping);
    case 4: return launch_fdu<5, 4, 4, 4, 2>(ctx, L, r, c, damping);
    case 5: return launch_fdu<6, 2, 2, 2, 3>(ctx, L, r, c, damping);
    default: return -1;
  }
}

}  // namespace hpdg
