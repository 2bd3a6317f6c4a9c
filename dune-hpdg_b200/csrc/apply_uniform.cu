// Uniform-degree 3-D SIPG operator apply on a structured brick: the headline kernel.
//
// Replaces Operator::apply over IPDGOperator (reference: matrix-free/operator.hh:41-56,
// matrix-free/localoperators/ipdgoperator.hh:80-390) / DynamicBCRSMatrix::mv
// (common/matrixwindow.hh:196-209) when every element has the same degree.
//
// Formulation: on an axis-parallel mesh the operator is a Kronecker sum, so with
// Tt_d = M^-1 T_d (T_d = 1-D SIPG line operator in direction d: block tridiagonal over the elements
// of a grid line, off-diagonal blocks of rank 2)
//        y = factor * (M_x M_y M_z) (Tt_x + Tt_y + Tt_z) u .
// A CTA owns a TX x TY x TZ tile of elements held in shared memory.  Five passes, each with one
// thread per DoF *line pencil* of the tile (a thread walks the tile's elements along the pass
// direction, so in-tile neighbour traces stay in registers and no trace exchange is needed):
//   P1 z-pencils : load u from global (coalesced), w  = Tt_z u        -> smem u, w
//   P2 x-pencils : w += Tt_x u
//   P3 y-pencils : w += Tt_y u ; w = M_y w
//   P4 x-pencils : w  = M_x w
//   P5 z-pencils : y  = factor * M_z w                                  -> global (coalesced)
// Traces of the elements just outside the tile come straight from global/L2 (or, on a rank
// boundary, from the ghost trace buffer filled by the NCCL halo exchange); domain boundaries are
// folded in as synthetic neighbour traces (Dirichlet / natural).
// Algorithmic HBM traffic: read u once + write y once = 16 B/DoF.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "ctx.hpp"
#include "uniform_common.cuh"
#include "apply_uniform_q3p.cuh"

namespace hpdg {

template <int N, int TX, int TY, int TZ>
constexpr int uni_threads() {
  int a = N * N * TY * TZ, b = N * N * TX * TZ, c = N * N * TX * TY;
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// One tile.  FULL: the tile has the full TX x TY x TZ extent (all masks compile away).
// The outside traces of every pass are loaded in two steps (uniform_common.cuh: halo_issue / halo_reduce): the loads of a
// pass are issued before the barrier that opens it and reduced to (der, val) pairs after it, so no warp waits on L2 in
// front of a barrier.
template <int N, int TX, int TY, int TZ, bool FULL>
__device__ __forceinline__ void tile_body(const UniParams<N>& P, double* __restrict__ su, double* __restrict__ sw,
                                          const int x0, const int y0, const int z0, const int lenx_rt, const int leny_rt,
                                          const int lenz_rt) {
  constexpr int PP = Pitch<N>::PP, EP = Pitch<N>::EP;
  constexpr int N2 = N * N, N3 = N * N * N;
  const int lenx = FULL ? TX : lenx_rt, leny = FULL ? TY : leny_rt, lenz = FULL ? TZ : lenz_rt;
  const long sx = N3, sy = (long)P.n[0] * N3, sz = (long)P.n[0] * P.n[1] * N3;  // element strides in doubles
  const double* __restrict__ X = P.x;
  const int tid = threadIdx.x;

  // ---------------- roles ----------------
  // z-role: (i, j, ex, ey)
  const int zi = tid % N, zj = (tid / N) % N, zex = (tid / N2) % TX, zey = tid / (N2 * TX);
  const bool zact = (tid < N2 * TX * TY) && zex < lenx && zey < leny;
  // x-role: (j, k, ey, ez)
  const int xj = tid % N, xk = (tid / N) % N, xey = (tid / N2) % TY, xez = tid / (N2 * TY);
  const bool xact = (tid < N2 * TY * TZ) && xey < leny && xez < lenz;
  // y-role: (i, ex, k, ez)
  const int yi = tid % N, yex = (tid / N) % TX, yk = (tid / (N * TX)) % N, yez = tid / (N2 * TX);
  const bool yact = (tid < N2 * TX * TZ) && yex < lenx && yez < lenz;

  // ---------------- P1: z-pencils, global -> registers -> smem ----------------
  if (zact) {
    const long ecol = (long)(x0 + zex) * sx + (long)(y0 + zey) * sy;  // element (x, y, 0)
    const int node = zi + N * zj;
    double v[TZ][N];
    const int zbase = (zex + TX * zey) * EP + node;
    if (P.xin_acc) {   // V-cycle: x += c fused into the apply of c (same coalesced access pattern as the loads of this pass)
#pragma unroll
      for (int e = 0; e < TZ; e++)
#pragma unroll
        for (int k = 0; k < N; k++)
          if (FULL || e < lenz) { const long g = ecol + (long)(z0 + e) * sz + node + N2 * k; P.xin_acc[g] += __ldg(X + g); }
    }
#pragma unroll
    for (int e = 0; e < TZ; e++)
#pragma unroll
      for (int k = 0; k < N; k++) v[e][k] = (FULL || e < lenz) ? __ldg(X + ecol + (long)(z0 + e) * sz + node + N2 * k) : 0.0;
    const long fe = (((long)(x0 + zex) + (long)P.n[0] * (y0 + zey)) * N2 + node) * 2;
    const HaloRaw<N> hr = halo_issue<N>(z0 == 0 ? P.bmode[4] : 0, z0 + lenz == P.n[2] ? P.bmode[5] : 0,
                                        X + ecol + (long)(z0 - 1) * sz + node, X + ecol + (long)(z0 + lenz) * sz + node, N2,
                                        P.ghost[4] + fe, P.ghost[5] + fe);
    const HaloTrace h = halo_reduce<N>(P, hr);
    pencil_apply<N, TZ, 2, FULL>(P, v, lenz, h.pd, h.pv, h.pm, h.nd, h.nv, h.nm,
      [](int, int) { return 0.0; },
      [&](int e, const double (&a)[N]) {
        const int base = zbase + TX * TY * EP * e;
#pragma unroll
        for (int k = 0; k < N; k++) { su[base + PP * k] = v[e][k]; sw[base + PP * k] = a[k]; }
      });
  }
  HaloRaw<N> hxr = HaloRaw<N>();
  if (xact) {
    const long erow = (long)(y0 + xey) * sy + (long)(z0 + xez) * sz;  // element (0, y, z)
    const long fe = (((long)(y0 + xey) + (long)P.n[1] * (z0 + xez)) * N2 + xj + N * xk) * 2;
    hxr = halo_issue<N>(x0 == 0 ? P.bmode[0] : 0, x0 + lenx == P.n[0] ? P.bmode[1] : 0,
                        X + erow + (long)(x0 - 1) * sx + N * xj + N2 * xk, X + erow + (long)(x0 + lenx) * sx + N * xj + N2 * xk, 1,
                        P.ghost[0] + fe, P.ghost[1] + fe);
  }
  __syncthreads();

  // ---------------- P2: x-pencils ----------------
  if (xact) {
    const HaloTrace h = halo_reduce<N>(P, hxr);
    double v[TX][N];
    const int xbase = TX * (xey + TY * xez) * EP + N * xj + PP * xk;
#pragma unroll
    for (int e = 0; e < TX; e++)
#pragma unroll
      for (int i = 0; i < N; i++) v[e][i] = (FULL || e < lenx) ? su[xbase + e * EP + i] : 0.0;
    pencil_apply<N, TX, 0, FULL>(P, v, lenx, h.pd, h.pv, h.pm, h.nd, h.nv, h.nm,
      [&](int e, int i) { return sw[xbase + e * EP + i]; },
      [&](int e, const double (&a)[N]) {
#pragma unroll
        for (int i = 0; i < N; i++) sw[xbase + e * EP + i] = a[i];
      });
  }
  HaloRaw<N> hyr = HaloRaw<N>();
  if (yact) {
    const long ecol = (long)(x0 + yex) * sx + (long)(z0 + yez) * sz;  // element (x, 0, z)
    const long fe = (((long)(x0 + yex) + (long)P.n[0] * (z0 + yez)) * N2 + yi + N * yk) * 2;
    hyr = halo_issue<N>(y0 == 0 ? P.bmode[2] : 0, y0 + leny == P.n[1] ? P.bmode[3] : 0,
                        X + ecol + (long)(y0 - 1) * sy + yi + N2 * yk, X + ecol + (long)(y0 + leny) * sy + yi + N2 * yk, N,
                        P.ghost[2] + fe, P.ghost[3] + fe);
  }
  __syncthreads();

  // ---------------- P3: y-pencils, then M_y ----------------
  if (yact) {
    const HaloTrace h = halo_reduce<N>(P, hyr);
    double v[TY][N];
    const int ybase = (yex + TX * TY * yez) * EP + yi + PP * yk;
#pragma unroll
    for (int e = 0; e < TY; e++)
#pragma unroll
      for (int j = 0; j < N; j++) v[e][j] = (FULL || e < leny) ? su[ybase + TX * EP * e + N * j] : 0.0;
    pencil_apply<N, TY, 1, FULL>(P, v, leny, h.pd, h.pv, h.pm, h.nd, h.nv, h.nm,
      [&](int e, int j) { return sw[ybase + TX * EP * e + N * j]; },
      [&](int e, const double (&a)[N]) {
        double b[N];
#pragma unroll
        for (int j = 0; j < N; j++) b[j] = a[j];
        mass_line<N>(P, b);
#pragma unroll
        for (int j = 0; j < N; j++) sw[ybase + TX * EP * e + N * j] = b[j];
      });
  }
  __syncthreads();

  // ---------------- P4: M_x ----------------
  if (xact) {
#pragma unroll
    for (int e = 0; e < TX; e++)
      if (FULL || e < lenx) {
        const int base = (e + TX * (xey + TY * xez)) * EP + N * xj + PP * xk;
        double a[N];
#pragma unroll
        for (int i = 0; i < N; i++) a[i] = sw[base + i];
        mass_line<N>(P, a);
#pragma unroll
        for (int i = 0; i < N; i++) sw[base + i] = a[i];
      }
  }
  __syncthreads();

  // ---------------- P5: M_z, write ----------------
  if (zact) {
    const long ecol = (long)(x0 + zex) * sx + (long)(y0 + zey) * sy;
    const int node = zi + N * zj;
#pragma unroll
    for (int e = 0; e < TZ; e++)
      if (FULL || e < lenz) {
        const int base = (zex + TX * (zey + TY * e)) * EP + node;
        double a[N];
#pragma unroll
        for (int k = 0; k < N; k++) a[k] = sw[base + PP * k];
        mass_line<N, true>(P, a);
        double* yo = P.y + ecol + (long)(z0 + e) * sz + node;
#pragma unroll
        for (int k = 0; k < N; k++) {
          yo[N2 * k] = P.accum ? yo[N2 * k] + a[k] : a[k];
        }
      }
  }
}

template <int N, int TX, int TY, int TZ, int MINB>
__global__ void __launch_bounds__(uni_threads<N, TX, TY, TZ>(), MINB)
k_apply_uniform(const __grid_constant__ UniParams<N> P) {
  constexpr int EP = Pitch<N>::EP;
  extern __shared__ double sm[];
  double* su = sm;
  double* sw = sm + TX * TY * TZ * EP;
  int tb = P.tile_list ? P.tile_list[blockIdx.x] : blockIdx.x + P.tile_offset;
  if (P.tile_rot) { tb += P.tile_rot; if (tb >= (int)gridDim.x) tb -= gridDim.x; }
  const int tx = tb % P.ntile[0]; tb /= P.ntile[0];
  const int ty = tb % P.ntile[1]; const int tz = tb / P.ntile[1];
  const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
  const int lenx = min(TX, P.n[0] - x0), leny = min(TY, P.n[1] - y0), lenz = min(TZ, P.n[2] - z0);
  if (P.ghost_step > 0) {  // p2p halo: tiles on a rank boundary wait until the neighbour's traces for this step have arrived
    const bool t0 = x0 == 0 && P.bmode[0] == 3, t1 = x0 + lenx == P.n[0] && P.bmode[1] == 3;
    const bool t2 = y0 == 0 && P.bmode[2] == 3, t3 = y0 + leny == P.n[1] && P.bmode[3] == 3;
    const bool t4 = z0 == 0 && P.bmode[4] == 3, t5 = z0 + lenz == P.n[2] && P.bmode[5] == 3;
    if (t0 || t1 || t2 || t3 || t4 || t5) {
      if (threadIdx.x == 0) {
        const bool touch[6] = {t0, t1, t2, t3, t4, t5};
        const long long tstart = clock64();
        for (int f = 0; f < 6; f++) {
          if (!touch[f]) continue;
          const volatile int* fl = P.ghost_flag[f];
          while (*fl < P.ghost_step) {
            __nanosleep(200);
            if (*reinterpret_cast<volatile int*>(P.ghost_err)) break;  // another tile already timed out: do not wait again
            if (clock64() - tstart > P.ghost_timeout) {  // give up, never hang the GPU; every synchronising entry point reports it
              atomicExch(P.ghost_err, 1); *reinterpret_cast<volatile int*>(P.ghost_err_host) = 1; __threadfence_system(); break;
            }
          }
        }
        __threadfence();
      }
      __syncthreads();
    }
  }
  if (lenx == TX && leny == TY && lenz == TZ) tile_body<N, TX, TY, TZ, true>(P, su, sw, x0, y0, z0, TX, TY, TZ);
  else tile_body<N, TX, TY, TZ, false>(P, su, sw, x0, y0, z0, lenx, leny, lenz);
}

// ---- ghost trace packing (sender side of the halo exchange) -------------------------
// For brick face f = (d,s): for every boundary element and face node, (der,val) of the element's
// DoF line normal to the face at side s.  Layout matches UniParams::ghost of the receiving rank.
struct PackParams { double* out[6]; const double* g[6]; };
template <int N>
__global__ void k_pack_traces(const double* __restrict__ x, PackParams PK, int n0, int n1, int n2) {
  constexpr int N2 = N * N, N3 = N2 * N;
  const int f = blockIdx.y, d = f / 2, s = f % 2;
  double* __restrict__ out = PK.out[f];
  if (!out) return;
  const double* __restrict__ g = PK.g[f];
  const int na = d == 0 ? n1 : n0, nb = d == 2 ? n1 : n2;  // face element extents (low dim fastest)
  const long total = (long)na * nb * N2;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int node = (int)(t % N2);
    const long fe = t / N2;
    const int a = (int)(fe % na), b = (int)(fe / na);
    const int p = node % N, q = node / N;
    long e; int off, stride;
    if (d == 0) { e = (s ? n0 - 1 : 0) + (long)n0 * (a + (long)n1 * b); off = N * p + N2 * q; stride = 1; }
    else if (d == 1) { e = a + (long)n0 * ((s ? n1 - 1 : 0) + (long)n1 * b); off = p + N2 * q; stride = N; }
    else { e = a + (long)n0 * (b + (long)n1 * (s ? n2 - 1 : 0)); off = p + N * q; stride = N2; }
    const double* line = x + e * N3 + off;
    double der = 0, val = 0;
#pragma unroll
    for (int m = 0; m < N; m++) {
      double u = line[m * stride];
      der = fma(g[m], u, der);
      if (m == (s ? N - 1 : 0)) val = u;
    }
    out[2 * t] = der; out[2 * t + 1] = val;
  }
}

// compact tile lists (tiles that touch no rank boundary / tiles that do), built once per (level, tile shape)
int uniform_tile_lists(Ctx* ctx, Level& L, int TX, int TY, int TZ, const int* bmode) {
  const long key = ((long)TX << 16) | ((long)TY << 8) | TZ;
  if (L.tile_key == key) return 0;
  const int nt[3] = {(L.n[0] + TX - 1) / TX, (L.n[1] + TY - 1) / TY, (L.n[2] + TZ - 1) / TZ};
  std::vector<int> li, lb;
  for (int tz = 0; tz < nt[2]; tz++) for (int ty = 0; ty < nt[1]; ty++) for (int tx = 0; tx < nt[0]; tx++) {
    const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
    const int lx = std::min(TX, L.n[0] - x0), ly = std::min(TY, L.n[1] - y0), lz = std::min(TZ, L.n[2] - z0);
    const bool touch = (x0 == 0 && bmode[0] == 3) || (x0 + lx == L.n[0] && bmode[1] == 3) ||
                       (y0 == 0 && bmode[2] == 3) || (y0 + ly == L.n[1] && bmode[3] == 3) ||
                       (z0 == 0 && bmode[4] == 3) || (z0 + lz == L.n[2] && bmode[5] == 3);
    (touch ? lb : li).push_back(tx + nt[0] * (ty + nt[1] * tz));
  }
  cudaFree(L.d_tiles_int); cudaFree(L.d_tiles_bnd); cudaFree(L.d_tiles_all);
  L.d_tiles_int = L.d_tiles_bnd = L.d_tiles_all = nullptr;
  {
    std::vector<int> all(li);
    all.insert(all.end(), lb.begin(), lb.end());
    HPDG_CUDA(cudaMalloc(&L.d_tiles_all, sizeof(int) * std::max<size_t>(all.size(), 1)));
    HPDG_CUDA(cudaMemcpy(L.d_tiles_all, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice));
  }
  HPDG_CUDA(cudaMalloc(&L.d_tiles_int, sizeof(int) * std::max<size_t>(li.size(), 1)));
  HPDG_CUDA(cudaMalloc(&L.d_tiles_bnd, sizeof(int) * std::max<size_t>(lb.size(), 1)));
  HPDG_CUDA(cudaMemcpy(L.d_tiles_int, li.data(), sizeof(int) * li.size(), cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMemcpy(L.d_tiles_bnd, lb.data(), sizeof(int) * lb.size(), cudaMemcpyHostToDevice));
  L.n_tiles_int = (long)li.size(); L.n_tiles_bnd = (long)lb.size(); L.tile_key = key;
  return 0;
}

template <int N, int TX, int TY, int TZ, int MINB>
static int launch_uni(Ctx* ctx, Level& L, const double* x, double* y, double factor, int part, cudaStream_t stream) {
  static thread_local UniParams<N> P;  // rebuilt per call (cheap); static to keep it off the stack, per thread: contexts on different host threads launch concurrently
  const DegTable& T = host_tables().deg[N - 1];
  const double c = ctx->sigma * (double)L.pen_uni * L.pen_uni;
  for (int d = 0; d < 3; d++) {
    double kap = 1.0 / L.h[d];
    for (int dd = 0; dd < 3; dd++) if (dd != d) kap *= L.h[dd];
    const double hk = 0.5 * kap;
    P.cohk[d] = c / hk;
    for (int i = 0; i < N; i++) {
      for (int j = 0; j < N; j++)
        P.Dp[d][i * N + j] = kap * T.MinvS[i * kMaxN + j]
                             + T.mt[0][i] * (hk * T.g[0][j] + c * T.t[0][j]) + T.mg[0][i] * (hk * T.t[0][j])
                             + T.mt[1][i] * (-hk * T.g[1][j] + c * T.t[1][j]) + T.mg[1][i] * (-hk * T.t[1][j]);
      P.A0[d][i] = hk * T.mt[0][i];
      P.B0[d][i] = -c * T.mt[0][i] - hk * T.mg[0][i];
      P.A1[d][i] = -hk * T.mt[1][i];
      P.B1[d][i] = -c * T.mt[1][i] + hk * T.mg[1][i];
    }
  }
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { P.M[i * N + j] = T.M[i * kMaxN + j]; P.Mf[i * N + j] = factor * T.M[i * kMaxN + j]; }
  for (int s = 0; s < 2; s++) for (int i = 0; i < N; i++) P.g[s][i] = T.g[s][i];
  P.factor = factor;
  const int tdim[3] = {TX, TY, TZ};
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.ntile[d] = (L.n[d] + tdim[d] - 1) / tdim[d]; }
  const bool finest = (&L == &ctx->levels.back());
  Ghost* Gp = nullptr;
  if (ctx->nranks > 1 && level_ghost(ctx, L, &Gp)) return 1;
  for (int f = 0; f < 6; f++) {
    P.ghost[f] = nullptr;
    if (ctx->bnd_is_rank[f]) {
      P.bmode[f] = 3;
      if (finest && ctx->ghost.p2p) {
        const int par = ctx->ghost.step & 1;
        P.ghost[f] = reinterpret_cast<const double*>(ctx->ghost.arena + ctx->ghost.recv_off[f][par]);
        P.ghost_flag[f] = reinterpret_cast<const int*>(ctx->ghost.arena + ctx->ghost.flag_off) + f * 2 + par;
      } else { P.ghost[f] = Gp->d_recv[f]; P.ghost_flag[f] = nullptr; }
    } else { P.bmode[f] = ctx->dirichlet ? 1 : 2; P.ghost_flag[f] = nullptr; }
  }
  P.ghost_step = (finest && ctx->ghost.p2p && part == 3) ? ctx->ghost.step : 0;
  P.ghost_err = ctx->ghost.p2p ? reinterpret_cast<int*>(ctx->ghost.arena + ctx->ghost.flag_off) + 12 : nullptr;
  P.ghost_err_host = ctx->d_ghost_err; P.ghost_timeout = ctx->halo_timeout_cycles;
  P.x = x; P.y = y; P.part = part; P.accum = ctx->fuse_accum; P.xin_acc = ctx->fuse_xin;
  P.tile_list = nullptr; P.tile_offset = 0; P.tile_rot = 0;
  long nlist = 0;
  if (part == 3) {
    P.tile_rot = (int)(((long)P.ntile[0] * P.ntile[1] * P.ntile[2]) / 2);
  } else if (part != 0) {
    if (uniform_tile_lists(ctx, L, TX, TY, TZ, P.bmode)) return 1;
    P.tile_list = part == 1 ? L.d_tiles_int : part == 2 ? L.d_tiles_bnd : L.d_tiles_all;
    nlist = part == 1 ? L.n_tiles_int : part == 2 ? L.n_tiles_bnd : L.n_tiles_int + L.n_tiles_bnd;
    if (nlist == 0) return 0;
  }
  constexpr int threads = uni_threads<N, TX, TY, TZ>();
  constexpr size_t smem = sizeof(double) * 2 * TX * TY * TZ * Pitch<N>::EP;
  const long ntiles_total = (long)P.ntile[0] * P.ntile[1] * P.ntile[2];
  long ntiles = (part == 1 || part == 2) ? nlist : ntiles_total;
  if (part == 0 && ctx->slab_nz > 0) {  // element layers [slab_z0, slab_z0 + slab_nz): must be multiples of TZ
    if (ctx->slab_z0 % TZ != 0) { ctx->err = "slab not aligned to the tile height"; return 1; }
    P.tile_offset = (ctx->slab_z0 / TZ) * P.ntile[0] * P.ntile[1];
    ntiles = (long)((ctx->slab_nz + TZ - 1) / TZ) * P.ntile[0] * P.ntile[1];
  }
  if constexpr (N == 4 && TX == 4 && TY == 4 && TZ == 4) {
    // default Q3 path: persistent CTAs with bulk-copy prefetch (apply_uniform_q3p.cuh); needs full tiles
    // (not for the two-launch NCCL halo path: its interior launch would starve the halo stream's kernels of SM slots)
    if (uniform_persistent(ctx, L, x) && part != 1 && part != 2) {
      int slots = 0;
      if (kernel_slots(ctx, reinterpret_cast<const void*>(hpdg_k_apply_q3_persist), 256, kQ3pSmemBytes, &slots)) return 1;
      if (q3p_level_setup(ctx, L)) return 1;
      const int grid = (int)std::min<long>(ntiles, ctx->q3p_grid > 0 ? ctx->q3p_grid : slots);
      Q3pPack PK = {};
      if (part == 3) {  // interior tiles first: by the time the CTAs reach the rank-boundary tiles the neighbours' traces have landed
        if (uniform_tile_lists(ctx, L, TX, TY, TZ, P.bmode)) return 1;
        P.tile_list = L.d_tiles_all; P.tile_rot = 0;
      }
      if (part == 3 && P.ghost_step > 0) {  // p2p halo: the tile kernel packs and publishes this rank's face traces itself
        const int par = ctx->ghost.step & 1;
        for (int f = 0; f < 6; f++) {
          if (!ctx->ghost.active[f]) continue;
          PK.out[f] = reinterpret_cast<double*>(ctx->ghost.peer_arena[f] + ctx->ghost.recv_off[f ^ 1][par]);
          PK.flag[f] = reinterpret_cast<int*>(ctx->ghost.peer_arena[f] + ctx->ghost.flag_off) + (f ^ 1) * 2 + par;
        }
        PK.done = ctx->d_sched + 8;
        PK.step = ctx->ghost.step;
        PK.npack = std::max(1, (grid + 2) / 3);
      }
      hpdg_k_apply_q3_persist<<<grid, 256, kQ3pSmemBytes, stream>>>(P, static_cast<const int4*>(L.d_tile_desc), (int)ntiles, (int)ntiles_total, ctx->d_sched + 2 * (part & 3), PK);
      ctx->launches++;
      HPDG_CUDA(cudaGetLastError());
      return 0;
    }
  }
  if (kernel_slots(ctx, reinterpret_cast<const void*>(k_apply_uniform<N, TX, TY, TZ, MINB>), threads, smem, nullptr)) return 1;
  k_apply_uniform<N, TX, TY, TZ, MINB><<<(unsigned)ntiles, threads, smem, stream>>>(P);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

int uniform_tile_height(const Level& L) {
  switch (L.p_uni) { case 4: return 3; case 5: return 2; default: return 4; }
}

// what the persistent Q3 kernels (operator apply, block Jacobi) need per context / level: the tile counters of the dynamic
// scheduler and the tile descriptors {first element, packed tile coordinates, brick-face bits, 0} of the 4x4x4 tiling
int q3p_level_setup(Ctx* ctx, Level& L, int tile_h) {  // tiles of 4 x 4 x tile_h elements
  if (!ctx->d_sched) {  // {next, done} per launch kind: apply parts 0..3 -> ints 0..7, halo pack 8, block Jacobi 10..11
    HPDG_CUDA(cudaMalloc(&ctx->d_sched, 12 * sizeof(int)));
    HPDG_CUDA(cudaMemset(ctx->d_sched, 0, 12 * sizeof(int)));
  }
  if (!L.d_tile_desc) {
    const int nt[3] = {L.n[0] / 4, L.n[1] / 4, L.n[2] / tile_h};
    std::vector<int4> td((size_t)nt[0] * nt[1] * nt[2]);
    for (int tz = 0, i = 0; tz < nt[2]; tz++) for (int ty = 0; ty < nt[1]; ty++) for (int tx = 0; tx < nt[0]; tx++, i++) {
      const int fl = (tx == 0) | (tx == nt[0] - 1) << 1 | (ty == 0) << 2 | (ty == nt[1] - 1) << 3 | (tz == 0) << 4 | (tz == nt[2] - 1) << 5;
      td[i] = make_int4(4 * tx + L.n[0] * (4 * ty + L.n[1] * tile_h * tz), tx | ty << 10 | tz << 20, fl, 0);
    }
    HPDG_CUDA(cudaMalloc(&L.d_tile_desc, sizeof(int4) * td.size()));
    HPDG_CUDA(cudaMemcpy(L.d_tile_desc, td.data(), sizeof(int4) * td.size(), cudaMemcpyHostToDevice));
  }
  return 0;
}

// the level's apply runs the persistent Q3 tile kernel for input vector x (bulk copies need 16-byte aligned rows; x == nullptr:
// alignment not checked).  Option "variant" = 40 switches the persistent kernels off (cross-check against the tile kernel).
int uniform_persistent(const Ctx* ctx, const Level& L, const double* x) {
  return uniform_supported(ctx, L) && L.p_uni == 3 && ctx->variant != 40 &&
         L.n[0] % 4 == 0 && L.n[1] % 4 == 0 && L.n[2] % 4 == 0 && L.n[0] <= 4092 && L.n[1] <= 4092 && L.n[2] <= 4092 &&
         L.ndof < (1L << 31) && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

int uniform_supported(const Ctx* ctx, const Level& L) {
  if (ctx->force_generic) return 0;
  if (!L.uniform || L.dim != 3) return 0;
  return L.p_uni >= 1 && L.p_uni <= 5;
}

int launch_apply_uniform(Ctx* ctx, Level& L, const double* x, double* y, double factor, int part, cudaStream_t stream) {
  if (!stream) stream = ctx->stream;
  switch (L.p_uni) {
    case 1: return launch_uni<2, 4, 4, 4, 4>(ctx, L, x, y, factor, part, stream);
    case 2: return launch_uni<3, 4, 4, 4, 3>(ctx, L, x, y, factor, part, stream);
    case 3: return launch_uni<4, 4, 4, 4, 3>(ctx, L, x, y, factor, part, stream);
    case 4: return launch_uni<5, 3, 3, 3, 3>(ctx, L, x, y, factor, part, stream);
    case 5: return launch_uni<6, 2, 2, 2, 2>(ctx, L, x, y, factor, part, stream);
    default: return -1;
  }
}

// halo buffers of a level: the context's for the finest level, lazily allocated per-level buffers (NCCL transport) otherwise
int level_ghost(Ctx* ctx, Level& L, Ghost** out) {
  if (&L == &ctx->levels.back()) { *out = &ctx->ghost; return 0; }
  Ghost& G = L.cg;
  if (!G.d_recv[0] && !G.d_recv[1] && !G.d_recv[2] && !G.d_recv[3] && !G.d_recv[4] && !G.d_recv[5]) {
    const int N2 = (L.p_uni + 1) * (L.p_uni + 1);
    for (int f = 0; f < 6; f++) {
      G.active[f] = ctx->ghost.active[f]; G.peer[f] = ctx->ghost.peer[f];
      if (!G.active[f]) continue;
      G.count[f] = ((size_t)L.nelem / L.n[f / 2]) * N2 * 2;
      HPDG_CUDA(cudaMalloc(&G.d_send[f], G.count[f] * sizeof(double)));
      HPDG_CUDA(cudaMalloc(&G.d_recv[f], G.count[f] * sizeof(double)));
    }
  }
  *out = &G;
  return 0;
}

template <int N>
static int pack_n(Ctx* ctx, Level& L, const double* x, cudaStream_t stream) {
  const DegTable* dt = ctx->d_tab + (N - 1);
  const bool finest = (&L == &ctx->levels.back());
  Ghost* Gp = nullptr;
  if (level_ghost(ctx, L, &Gp)) return 1;
  PackParams PK;
  long maxtotal = 0;
  for (int f = 0; f < 6; f++) {
    PK.out[f] = !Gp->active[f] ? nullptr
              : (finest && ctx->ghost.p2p) ? reinterpret_cast<double*>(ctx->ghost.peer_arena[f] + ctx->ghost.recv_off[f ^ 1][ctx->ghost.step & 1])
                                           : Gp->d_send[f];
    PK.g[f] = &dt->g[f % 2][0];
    if (Gp->active[f]) maxtotal = std::max<long>(maxtotal, (long)Gp->count[f] / 2);
  }
  if (maxtotal == 0) return 0;
  const int threads = 256;
  dim3 grid((unsigned)std::min<long>((maxtotal + threads - 1) / threads, 2048), 6);
  k_pack_traces<N><<<grid, threads, 0, stream>>>(x, PK, L.n[0], L.n[1], L.n[2]);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

// after the pack kernel (stream order): make its peer stores visible system-wide, then raise the neighbours' flags
struct FlagParams { int* flag[6]; int step; };
__global__ void k_halo_flags(FlagParams F) {
  const int f = threadIdx.x;
  if (f < 6 && F.flag[f]) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(F.flag[f]) = F.step;
    __threadfence_system();
  }
}
int launch_halo_flags(Ctx* ctx, cudaStream_t stream) {
  FlagParams F;
  F.step = ctx->ghost.step;
  for (int f = 0; f < 6; f++)
    F.flag[f] = ctx->ghost.active[f] ? reinterpret_cast<int*>(ctx->ghost.peer_arena[f] + ctx->ghost.flag_off) + (f ^ 1) * 2 + (ctx->ghost.step & 1) : nullptr;
  k_halo_flags<<<1, 32, 0, stream>>>(F);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_traces(Ctx* ctx, Level& L, const double* x, cudaStream_t stream) {
  switch (L.p_uni) {
    case 1: return pack_n<2>(ctx, L, x, stream);
    case 2: return pack_n<3>(ctx, L, x, stream);
    case 3: return pack_n<4>(ctx, L, x, stream);
    case 4: return pack_n<5>(ctx, L, x, stream);
    case 5: return pack_n<6>(ctx, L, x, stream);
    default: ctx->err = "pack_traces: unsupported degree"; return 1;
  }
}

}  // namespace hpdg
