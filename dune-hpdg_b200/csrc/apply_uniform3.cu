// EXPERIMENTAL (variant 21/22, not the default): three-pass variant of the uniform-degree SIPG apply for N = 4 (Q3).
// Measured 134-152 us on cfg2 against 101 us for the five-pass kernel (round 1): 127 registers -> 2 CTAs/SM, local-memory
// traffic and the halo shuffles outweigh the saved shared-memory passes.  Kept for the next round's tuning.
//
// Same operator and same reference interfaces as apply_uniform.cu (Operator::apply over IPDGOperator,
// matrix-free/operator.hh:41-56, matrix-free/localoperators/ipdgoperator.hh:80-390).  Instead of premultiplying by M^-1 and
// applying the three mass sweeps afterwards (5 passes, 11 shared-memory accesses per DoF), the Kronecker sum
//     A = T_x (x) M_y (x) M_z + M_x (x) T_y (x) M_z + M_x (x) M_y (x) T_z
// is evaluated as 7 one-dimensional sweeps in 3 passes, 8 shared-memory accesses per DoF, 2 barriers:
//     X (x-pencils):  a = M_x u,  b = T_x u                       (u straight from global, 32-byte lines)
//     Y (y-pencils):  c = M_y a,  q = M_y b + T_y a
//     Z (z-pencils):  y = factor * (M_z q + T_z c)                 (coalesced store)
// T_d is the 1-D SIPG line operator with the element's own face terms folded into its diagonal block (Dp) and the neighbour
// coupling expressed through the neighbour's (der, val) trace; in this un-premultiplied form the `t` traces are unit vectors, so
// a neighbour costs N + 2 FMAs per line instead of 2N.  The traces that T_y / T_z need from elements outside the tile are
// traces of a = M_x u and c = M_y M_x u: they are formed from the raw u lines of the outside element and the 1-D mass is then
// applied across the 4 lanes that hold a line's nodes with warp shuffles (no shared memory, no extra barrier).
#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "ctx.hpp"

namespace hpdg {

struct Uni3Params {
  double Dp[3][16];      // kappa_d S + own face terms (interior faces on both sides); [2] pre-scaled by factor
  double B0[3][4], B1[3][4];
  double hk[3];          // kappa_d / 2 ([2] pre-scaled by factor)
  double cohk[3];        // c / (kappa_d / 2)  (never scaled)
  double M[16], Mf[16];  // mass, factor * mass
  double g[2][4];
  int n[3];
  int ntile[3];
  int bmode[6];
  const double* ghost[6];
  const double* x;
  double* y;
  const int* tile_list;
  int tile_offset;
};

// out_e = T_dir v_e (+ acc) along one pencil of T elements.  SC: 0 for directions x,y; 1 uses the factor-scaled constants of z.
template <int T, int DIR, bool FULL, class Out>
__device__ __forceinline__ void pencil_T(const Uni3Params& P, const double (&v)[T][4], int len_rt, double pd, double pv, int pmode,
                                         double nd, double nv, int nmode, Out out) {
  constexpr int N = 4;
  const int len = FULL ? T : len_rt;
  double d0[T], d1[T];
#pragma unroll
  for (int e = 0; e < T; e++) {
    double a = 0, b = 0;
#pragma unroll
    for (int m = 0; m < N; m++) { a = fma(P.g[0][m], v[e][m], a); b = fma(P.g[1][m], v[e][m], b); }
    d0[e] = a; d1[e] = b;
  }
  if (pmode == 1) { pd = fma(-P.cohk[DIR], v[0][0], d0[0]); pv = -v[0][0]; }
  else if (pmode == 2) { pd = -d0[0]; pv = v[0][0]; }
  {
    double dl = d1[T - 1], vl = v[T - 1][N - 1];
    if (!FULL) {
#pragma unroll
      for (int e = 0; e < T - 1; e++) if (e == len - 1) { dl = d1[e]; vl = v[e][N - 1]; }
    }
    if (nmode == 1) { nd = fma(P.cohk[DIR], vl, dl); nv = -vl; }
    else if (nmode == 2) { nd = -dl; nv = vl; }
  }
#pragma unroll
  for (int e = 0; e < T; e++) {
    if (FULL || e < len) {
      const double qd = (e == 0) ? pd : d1[e > 0 ? e - 1 : 0];
      const double qv = (e == 0) ? pv : v[e > 0 ? e - 1 : 0][N - 1];
      double rd = (e == T - 1) ? nd : d0[e < T - 1 ? e + 1 : e];
      double rv = (e == T - 1) ? nv : v[e < T - 1 ? e + 1 : e][0];
      if (!FULL && e == len - 1) { rd = nd; rv = nv; }
      double a[N];
#pragma unroll
      for (int i = 0; i < N; i++) {
        double s = 0;
#pragma unroll
        for (int m = 0; m < N; m++) s = fma(P.Dp[DIR][i * N + m], v[e][m], s);
        s = fma(P.B0[DIR][i], qv, s);
        s = fma(P.B1[DIR][i], rv, s);
        a[i] = s;
      }
      a[0] = fma(P.hk[DIR], qd, a[0]);           // A0 = hk t_0 : unit vector
      a[N - 1] = fma(-P.hk[DIR], rd, a[N - 1]);   // A1 = -hk t_1
      out(e, a);
    }
  }
}

__device__ __forceinline__ void mass4(const double (&M)[16], const double (&a)[4], double (&o)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) s = fma(M[i * 4 + m], a[m], s);
    o[i] = s;
  }
}

// 1-D mass applied across the 4 lanes {base, base+stride, base+2 stride, base+3 stride} that hold one line's nodes;
// `me` is this lane's node index on that line.  All 32 lanes must call it.
__device__ __forceinline__ double lane_mass(const double (&M)[16], double val, int base_lane, int stride, int me) {
  double s = 0;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double o = __shfl_sync(0xffffffffu, val, base_lane + m * stride);
    s = fma(M[me * 4 + m], o, s);  // run-time row: a constant-bank load with a register offset
  }
  return s;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_apply_uniform3(const __grid_constant__ Uni3Params P) {
  constexpr int N = 4, T = 4, N2 = 16, N3 = 64;
  constexpr int PP = 17, EP = 68;  // z-plane pitch / element pitch: conflict-free for all three pencil directions (apply_uniform.cu)
  extern __shared__ double sm[];
  double* s0 = sm;                  // a, then c
  double* s1 = sm + T * T * T * EP;  // b, then q
  int tb = P.tile_list ? P.tile_list[blockIdx.x] : blockIdx.x + P.tile_offset;
  const int tx = tb % P.ntile[0]; tb /= P.ntile[0];
  const int ty = tb % P.ntile[1]; const int tz = tb / P.ntile[1];
  const int x0 = tx * T, y0 = ty * T, z0 = tz * T;
  const int lenx = min(T, P.n[0] - x0), leny = min(T, P.n[1] - y0), lenz = min(T, P.n[2] - z0);
  const bool full = lenx == T && leny == T && lenz == T;
  const long sx = N3, sy = (long)P.n[0] * N3, sz = (long)P.n[0] * P.n[1] * N3;
  const double* __restrict__ X = P.x;
  const int tid = threadIdx.x, lane = tid & 31;

  // roles (identical to apply_uniform.cu)
  const int xj = tid % N, xk = (tid / N) % N, xey = (tid / N2) % T, xez = tid / (N2 * T);
  const bool xact = xey < leny && xez < lenz;
  const int yi = tid % N, yex = (tid / N) % T, yk = (tid / (N * T)) % N, yez = tid / (N2 * T);
  const bool yact = yex < lenx && yez < lenz;
  const int zi = tid % N, zj = (tid / N) % N, zex = (tid / N2) % T, zey = tid / (N2 * T);
  const bool zact = zex < lenx && zey < leny;

  // ------------------------------------------------------------------ pass X
  // y-direction outside traces of u (raw), issued first so that they are in flight during the x sweep
  double yraw[2][2] = {{0, 0}, {0, 0}}; int ypm = 0, ynm = 0;
  {
    const long ecol = (long)(x0 + yex) * sx + (long)(z0 + yez) * sz;
    const int node = yi + N * yk;
    if (y0 == 0) {
      ypm = P.bmode[2];
      if (ypm == 3) { if (yact) { const double* gp = P.ghost[2] + (((long)(x0 + yex) + (long)P.n[0] * (z0 + yez)) * N2 + node) * 2; yraw[0][0] = gp[0]; yraw[0][1] = gp[1]; } ypm = 0; }
    } else if (yact) {
      const double* line = X + ecol + (long)(y0 - 1) * sy + yi + N2 * yk;
      double d = 0, u = 0;
#pragma unroll
      for (int m = 0; m < N; m++) { u = __ldg(line + m * N); d = fma(P.g[1][m], u, d); }
      yraw[0][0] = d; yraw[0][1] = u;
    }
    if (y0 + leny == P.n[1]) {
      ynm = P.bmode[3];
      if (ynm == 3) { if (yact) { const double* gp = P.ghost[3] + (((long)(x0 + yex) + (long)P.n[0] * (z0 + yez)) * N2 + node) * 2; yraw[1][0] = gp[0]; yraw[1][1] = gp[1]; } ynm = 0; }
    } else if (yact) {
      const double* line = X + ecol + (long)(y0 + leny) * sy + yi + N2 * yk;
      double d = 0, u0 = 0;
#pragma unroll
      for (int m = 0; m < N; m++) { const double u = __ldg(line + m * N); d = fma(P.g[0][m], u, d); if (m == 0) u0 = u; }
      yraw[1][0] = d; yraw[1][1] = u0;
    }
  }
  if (xact) {
    const long erow = (long)(y0 + xey) * sy + (long)(z0 + xez) * sz;
    const double* base = X + erow + (long)x0 * sx + N * xj + N2 * xk;
    double v[T][N];
#pragma unroll
    for (int e = 0; e < T; e++) {
      if (full || e < lenx) {
        const double2 lo = __ldg(reinterpret_cast<const double2*>(base + e * sx));
        const double2 hi = __ldg(reinterpret_cast<const double2*>(base + e * sx) + 1);
        v[e][0] = lo.x; v[e][1] = lo.y; v[e][2] = hi.x; v[e][3] = hi.y;
      } else { v[e][0] = v[e][1] = v[e][2] = v[e][3] = 0.0; }
    }
    double pd = 0, pv = 0, nd = 0, nv = 0; int pm = 0, nm = 0;
    const int node = xj + N * xk;
    if (x0 == 0) {
      pm = P.bmode[0];
      if (pm == 3) { const double* gp = P.ghost[0] + (((long)(y0 + xey) + (long)P.n[1] * (z0 + xez)) * N2 + node) * 2; pd = gp[0]; pv = gp[1]; pm = 0; }
    } else {
      const double2 lo = __ldg(reinterpret_cast<const double2*>(base - sx));
      const double2 hi = __ldg(reinterpret_cast<const double2*>(base - sx) + 1);
      pd = fma(P.g[1][0], lo.x, fma(P.g[1][1], lo.y, fma(P.g[1][2], hi.x, P.g[1][3] * hi.y))); pv = hi.y;
    }
    if (x0 + lenx == P.n[0]) {
      nm = P.bmode[1];
      if (nm == 3) { const double* gp = P.ghost[1] + (((long)(y0 + xey) + (long)P.n[1] * (z0 + xez)) * N2 + node) * 2; nd = gp[0]; nv = gp[1]; nm = 0; }
    } else {
      const double2 lo = __ldg(reinterpret_cast<const double2*>(base + (long)lenx * sx));
      const double2 hi = __ldg(reinterpret_cast<const double2*>(base + (long)lenx * sx) + 1);
      nd = fma(P.g[0][0], lo.x, fma(P.g[0][1], lo.y, fma(P.g[0][2], hi.x, P.g[0][3] * hi.y))); nv = lo.x;
    }
    const int xbase = T * (xey + T * xez) * EP + N * xj + PP * xk;
    auto outx = [&](int e, const double (&b)[N]) {
      double a[N];
      mass4(P.M, v[e], a);
#pragma unroll
      for (int i = 0; i < N; i++) { s0[xbase + e * EP + i] = a[i]; s1[xbase + e * EP + i] = b[i]; }
    };
    if (full) pencil_T<T, 0, true>(P, v, T, pd, pv, pm, nd, nv, nm, outx);
    else pencil_T<T, 0, false>(P, v, lenx, pd, pv, pm, nd, nv, nm, outx);
  }
  // traces of a = M_x u outside the tile in y: mass across the 4 lanes i = 0..3 of the y-role (all lanes participate)
  double ypd, ypv, ynd, ynv;
  {
    const int b4 = lane & ~3;
    ypd = lane_mass(P.M, yraw[0][0], b4, 1, yi); ypv = lane_mass(P.M, yraw[0][1], b4, 1, yi);
    ynd = lane_mass(P.M, yraw[1][0], b4, 1, yi); ynv = lane_mass(P.M, yraw[1][1], b4, 1, yi);
  }
  __syncthreads();

  // ------------------------------------------------------------------ pass Y
  // z-direction outside traces of u (raw) in flight during the y sweep
  double zraw[2][2] = {{0, 0}, {0, 0}}; int zpm = 0, znm = 0;
  {
    const long ecol = (long)(x0 + zex) * sx + (long)(y0 + zey) * sy;
    const int node = zi + N * zj;
    if (z0 == 0) {
      zpm = P.bmode[4];
      if (zpm == 3) { if (zact) { const double* gp = P.ghost[4] + (((long)(x0 + zex) + (long)P.n[0] * (y0 + zey)) * N2 + node) * 2; zraw[0][0] = gp[0]; zraw[0][1] = gp[1]; } zpm = 0; }
    } else if (zact) {
      const double* line = X + ecol + (long)(z0 - 1) * sz + node;
      double d = 0, u = 0;
#pragma unroll
      for (int m = 0; m < N; m++) { u = __ldg(line + m * N2); d = fma(P.g[1][m], u, d); }
      zraw[0][0] = d; zraw[0][1] = u;
    }
    if (z0 + lenz == P.n[2]) {
      znm = P.bmode[5];
      if (znm == 3) { if (zact) { const double* gp = P.ghost[5] + (((long)(x0 + zex) + (long)P.n[0] * (y0 + zey)) * N2 + node) * 2; zraw[1][0] = gp[0]; zraw[1][1] = gp[1]; } znm = 0; }
    } else if (zact) {
      const double* line = X + ecol + (long)(z0 + lenz) * sz + node;
      double d = 0, u0 = 0;
#pragma unroll
      for (int m = 0; m < N; m++) { const double u = __ldg(line + m * N2); d = fma(P.g[0][m], u, d); if (m == 0) u0 = u; }
      zraw[1][0] = d; zraw[1][1] = u0;
    }
  }
  if (yact) {
    const int ybase = (yex + T * T * yez) * EP + yi + PP * yk;
    double va[T][N];
#pragma unroll
    for (int e = 0; e < T; e++)
#pragma unroll
      for (int j = 0; j < N; j++) va[e][j] = (full || e < leny) ? s0[ybase + T * EP * e + N * j] : 0.0;
    auto outy = [&](int e, const double (&ta)[N]) {
      double vb[N], c[N], mb[N];
#pragma unroll
      for (int j = 0; j < N; j++) vb[j] = s1[ybase + T * EP * e + N * j];
      mass4(P.M, va[e], c);
      mass4(P.M, vb, mb);
#pragma unroll
      for (int j = 0; j < N; j++) { s0[ybase + T * EP * e + N * j] = c[j]; s1[ybase + T * EP * e + N * j] = mb[j] + ta[j]; }
    };
    if (full) pencil_T<T, 1, true>(P, va, T, ypd, ypv, ypm, ynd, ynv, ynm, outy);
    else pencil_T<T, 1, false>(P, va, leny, ypd, ypv, ypm, ynd, ynv, ynm, outy);
  }
  // traces of c = M_y M_x u outside the tile in z: masses across lanes i (stride 1) and j (stride 4) of the z-role
  double zpd, zpv, znd, znv;
  {
    const int b4 = lane & ~3, b16 = (lane & ~15) | (lane & 3);
    double t;
    t = lane_mass(P.M, zraw[0][0], b4, 1, zi); zpd = lane_mass(P.M, t, b16, 4, zj);
    t = lane_mass(P.M, zraw[0][1], b4, 1, zi); zpv = lane_mass(P.M, t, b16, 4, zj);
    t = lane_mass(P.M, zraw[1][0], b4, 1, zi); znd = lane_mass(P.M, t, b16, 4, zj);
    t = lane_mass(P.M, zraw[1][1], b4, 1, zi); znv = lane_mass(P.M, t, b16, 4, zj);
  }
  __syncthreads();

  // ------------------------------------------------------------------ pass Z
  if (zact) {
    const int node = zi + N * zj;
    const int zbase = (zex + T * zey) * EP + node;
    double* yo = P.y + (long)(x0 + zex) * sx + (long)(y0 + zey) * sy + (long)z0 * sz + node;
    double vc[T][N];
#pragma unroll
    for (int e = 0; e < T; e++)
#pragma unroll
      for (int k = 0; k < N; k++) vc[e][k] = (full || e < lenz) ? s0[zbase + T * T * EP * e + PP * k] : 0.0;
    auto outz = [&](int e, const double (&tc)[N]) {
      double vq[N], mq[N];
#pragma unroll
      for (int k = 0; k < N; k++) vq[k] = s1[zbase + T * T * EP * e + PP * k];
      mass4(P.Mf, vq, mq);
#pragma unroll
      for (int k = 0; k < N; k++) yo[(long)e * sz + N2 * k] = mq[k] + tc[k];
    };
    if (full) pencil_T<T, 2, true>(P, vc, T, zpd, zpv, zpm, znd, znv, znm, outz);
    else pencil_T<T, 2, false>(P, vc, lenz, zpd, zpv, zpm, znd, znv, znm, outz);
  }
}

int launch_apply_uniform3(Ctx* ctx, Level& L, const double* x, double* y, double factor, int part, cudaStream_t stream) {
  static thread_local Uni3Params P;
  constexpr int N = 4, T = 4;
  const DegTable& Tb = host_tables().deg[N - 1];
  const double c = ctx->sigma * (double)L.pen_uni * L.pen_uni;
  for (int d = 0; d < 3; d++) {
    double kap = 1.0 / L.h[d];
    for (int dd = 0; dd < 3; dd++) if (dd != d) kap *= L.h[dd];
    const double hk = 0.5 * kap;
    const double sc = d == 2 ? factor : 1.0;  // the z sweep is the last one: fold the operator's factor into its constants
    P.hk[d] = sc * hk;
    P.cohk[d] = c / hk;
    for (int i = 0; i < N; i++) {
      for (int j = 0; j < N; j++)
        P.Dp[d][i * N + j] = sc * (kap * Tb.S[i * kMaxN + j]
                                   + Tb.t[0][i] * (hk * Tb.g[0][j] + c * Tb.t[0][j]) + Tb.g[0][i] * (hk * Tb.t[0][j])
                                   + Tb.t[1][i] * (-hk * Tb.g[1][j] + c * Tb.t[1][j]) + Tb.g[1][i] * (-hk * Tb.t[1][j]));
      P.B0[d][i] = sc * (-c * Tb.t[0][i] - hk * Tb.g[0][i]);
      P.B1[d][i] = sc * (-c * Tb.t[1][i] + hk * Tb.g[1][i]);
    }
  }
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { P.M[i * N + j] = Tb.M[i * kMaxN + j]; P.Mf[i * N + j] = factor * Tb.M[i * kMaxN + j]; }
  for (int s = 0; s < 2; s++) for (int i = 0; i < N; i++) P.g[s][i] = Tb.g[s][i];
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.ntile[d] = (L.n[d] + T - 1) / T; }
  const bool finest = (&L == &ctx->levels.back());
  for (int f = 0; f < 6; f++) {
    P.ghost[f] = nullptr;
    if (ctx->bnd_is_rank[f]) {
      if (!finest || ctx->ghost.p2p) { ctx->err = "the experimental kernel supports the NCCL halo on the finest level only"; return 1; }
      P.bmode[f] = 3; P.ghost[f] = ctx->ghost.d_recv[f];
    } else P.bmode[f] = ctx->dirichlet ? 1 : 2;
  }
  P.x = x; P.y = y; P.tile_list = nullptr; P.tile_offset = 0;
  long ntiles = (long)P.ntile[0] * P.ntile[1] * P.ntile[2];
  if (part != 0) {
    if (uniform_tile_lists(ctx, L, T, T, T, P.bmode)) return 1;
    P.tile_list = part == 1 ? L.d_tiles_int : L.d_tiles_bnd;
    ntiles = part == 1 ? L.n_tiles_int : L.n_tiles_bnd;
    if (ntiles == 0) return 0;
  } else if (ctx->slab_nz > 0) {
    if (ctx->slab_z0 % T != 0) { ctx->err = "slab not aligned to the tile height"; return 1; }
    P.tile_offset = (ctx->slab_z0 / T) * P.ntile[0] * P.ntile[1];
    ntiles = (long)((ctx->slab_nz + T - 1) / T) * P.ntile[0] * P.ntile[1];
  }
  constexpr size_t smem = sizeof(double) * 2 * T * T * T * 68;
  static bool attr_set = false;
  if (!attr_set) {
    HPDG_CUDA(cudaFuncSetAttribute(k_apply_uniform3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HPDG_CUDA(cudaFuncSetAttribute(k_apply_uniform3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (!stream) stream = ctx->stream;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) { ctx->err = "device vectors must be 16-byte aligned"; return 1; }
  if (ctx->variant == 22) k_apply_uniform3<2><<<(unsigned)ntiles, 256, smem, stream>>>(P);
  else k_apply_uniform3<3><<<(unsigned)ntiles, 256, smem, stream>>>(P);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hpdg
