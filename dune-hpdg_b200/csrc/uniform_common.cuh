// Shared pieces of the uniform-degree tile kernels (apply_uniform.cu, apply_uniform_q3p.cu): the parameter block, the
// shared-memory pitches of the padded layout, the pencil sweep and the 1-D mass sweep.  See apply_uniform.cu for the formulation.
#pragma once
#include "ctx.hpp"

namespace hpdg {

template <int N>
struct UniParams {
  // Dp = kappa_d M^-1 S + the element's own face terms (interior faces on both sides) folded in;
  // (A0,B0)/(A1,B1): response to the (der,val) trace of the previous/next element
  double Dp[3][N * N];
  double A0[3][N], B0[3][N], A1[3][N], B1[3][N];
  double M[N * N];
  double Mf[N * N];  // factor * M: the last (z) mass sweep carries the operator's factor
  double g[2][N];
  double cohk[3];  // cpen / (kappa_d / 2)
  double factor;
  int n[3];
  int ntile[3];
  int bmode[6];            // brick face: 1 Dirichlet, 2 natural, 3 ghost traces
  const double* ghost[6];  // [face elem][node][2] = (der, val) of the remote element at its near side
  const int* ghost_flag[6];  // p2p halo: flag the neighbour raises to ghost_step once its traces for this step have landed
  int ghost_step;
  int* ghost_err;
  const double* x;
  double* y;
  int accum; // y = y_old + factor * A x
  int part;  // 0 all tiles, 1 only tiles not touching a ghost face, 2 only tiles touching one
  const int* tile_list;  // part != 0: compact list of the tile ids of that part (grid = list length)
  int tile_offset;       // first tile of this launch (z-slab launches of the chunked host-pointer apply)
  int tile_rot;          // p2p halo: natural tile order rotated by this much, so the first wave is mid-domain tiles
};

// The GL nodes are symmetric about the element centre, so with R the node reflection: M and Dp commute with R, M is symmetric,
// g_1 = -R g_0, A1 = -R A0, B1 = R B0.  The kernels read one representative of each group of equal table entries, which roughly
// halves the number of distinct constants a pass needs (uniform-register pressure, LDCU count).
template <int N> __host__ __device__ constexpr int sym_dp_idx(int i, int m) {
  return (i * N + m <= (N - 1 - i) * N + (N - 1 - m)) ? i * N + m : (N - 1 - i) * N + (N - 1 - m);
}
template <int N> __host__ __device__ constexpr int sym_m_idx(int i, int m) {
  int best = i * N + m;
  const int c1 = m * N + i, c2 = (N - 1 - i) * N + (N - 1 - m), c3 = (N - 1 - m) * N + (N - 1 - i);
  if (c1 < best) best = c1;
  if (c2 < best) best = c2;
  if (c3 < best) best = c3;
  return best;
}

template <int N> struct Pitch {
  static constexpr int PP = (N % 2 == 0) ? N * N + 1 : N * N;  // z-plane pitch
  static constexpr int EP0 = N * PP;
  static constexpr int EP = EP0 + ((N - EP0 % 16) % 16 + 16) % 16;  // element pitch == N (mod 16)
};

// acc[e][:] += (Tt_dir v)_e for the elements e < len of one pencil.
// (pd,pv)/(nd,nv): (der,val) of the element before / after the pencil at its near side;
// pmode/nmode: 0 use them, 1 Dirichlet boundary, 2 natural boundary.  Boundaries are folded in as
// synthetic neighbour traces so the inner loop is branch free:
//   Dirichlet: (der -/+ (c/hk) val, -val)   natural: (-der, val)   of the element's own trace.
// accin(e,i) supplies the accumulator's initial value, out(e,a) consumes element e's N results
// right after they are formed (keeps the live register set to one element's worth).
template <int N, int T, int DIR, bool FULL, class AccIn, class Out>
__device__ __forceinline__ void pencil_apply(const UniParams<N>& P, const double (&v)[T][N], int len_rt, double pd,
                                             double pv, int pmode, double nd, double nv, int nmode, AccIn accin, Out out) {
  const int len = FULL ? T : len_rt;
  double d0[T], d1[T];
#pragma unroll
  for (int e = 0; e < T; e++) {
    double a = 0, b = 0;
#pragma unroll
    for (int m = 0; m < N; m++) { a = fma(P.g[0][m], v[e][m], a); b = fma(-P.g[0][N - 1 - m], v[e][m], b); }
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
        double s = accin(e, i);
#pragma unroll
        for (int m = 0; m < N; m++) s = fma(P.Dp[DIR][sym_dp_idx<N>(i, m)], v[e][m], s);
        s = fma(P.A0[DIR][i], qd, s); s = fma(P.B0[DIR][i], qv, s);
        s = fma(-P.A0[DIR][N - 1 - i], rd, s); s = fma(P.B0[DIR][N - 1 - i], rv, s);
        a[i] = s;
      }
      out(e, a);
    }
  }
}

template <int N, bool SCALED = false>
__device__ __forceinline__ void mass_line(const UniParams<N>& P, double (&a)[N]) {
  double o[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < N; m++) s = fma(SCALED ? P.Mf[sym_m_idx<N>(i, m)] : P.M[sym_m_idx<N>(i, m)], a[m], s);
    o[i] = s;
  }
#pragma unroll
  for (int i = 0; i < N; i++) a[i] = o[i];
}

// Trace (der, val) of the element outside the tile across brick-interior faces: read its DoF line.
template <int N>
__device__ __forceinline__ void outside_trace(const UniParams<N>& P, const double* __restrict__ line, long stride,
                                              int side /* near side of that element */, double& der, double& val) {
  if (N == 4 && stride == 1) {  // an x line is 32 contiguous, 32-byte aligned bytes: two 128-bit loads
    const double2 lo = __ldg(reinterpret_cast<const double2*>(line));
    const double2 hi = __ldg(reinterpret_cast<const double2*>(line) + 1);
    der = fma(P.g[side][0], lo.x, fma(P.g[side][1], lo.y, fma(P.g[side][2], hi.x, P.g[side][3] * hi.y)));
    val = side ? hi.y : lo.x;
    return;
  }
  double d = 0, last = 0, first = 0;
#pragma unroll
  for (int m = 0; m < N; m++) {
    double u = __ldg(line + m * stride);
    d = fma(P.g[side][m], u, d);
    if (m == 0) first = u;
    if (m == N - 1) last = u;
  }
  der = d; val = side ? last : first;
}

}  // namespace hpdg
