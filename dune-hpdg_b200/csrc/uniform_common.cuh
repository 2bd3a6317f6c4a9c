// Shared pieces of the uniform-degree tile kernels (apply_uniform.cu, apply_uniform_q3p.cu): the parameter block, the
// shared-memory pitches of the padded layout, the pencil sweep and the 1-D mass sweep.  See apply_uniform.cu for the formulation.
#pragma once
#include <cstdint>

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
  int* ghost_err;            // p2p halo time-out flag (device, in the halo arena) ...
  int* ghost_err_host;       // ... and its copy in mapped pinned host memory, which the synchronising entry points read
  long long ghost_timeout;   // SM clock cycles a rank-boundary tile waits for the neighbour's flag before giving up
  const double* x;
  double* y;
  int accum; // y = y_old + factor * A x
  double* xin_acc;  // optional: xin_acc += x (the V-cycle's x += c rides on the apply of c: its first pass holds c anyway)
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

// Traces (der, val) of the two elements just outside a pencil, in two steps so that the global-memory latency is not paid
// where the loads are issued: halo_issue() only issues the loads (the raw DoF lines of the outside elements across
// brick-interior faces, or the ghost trace pair on a rank boundary) -- both sides back to back, nothing consumed --
// and halo_reduce() turns them into traces.  The kernels issue before a barrier and reduce after it.
template <int N>
struct HaloRaw { double p[N], n[N]; int pm, nm; };  // modes: 0 interior, 1 Dirichlet, 2 natural, 3 ghost (rank boundary)

template <int N>
__device__ __forceinline__ void halo_load_line(double (&l)[N], const double* __restrict__ line, long stride) {
  if (N == 4 && stride == 1) {  // an x line is 32 contiguous bytes, 32-byte aligned if the vector is: one 256-bit load
    if ((reinterpret_cast<uintptr_t>(line) & 31) == 0) {
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];\n" : "=d"(l[0]), "=d"(l[1]), "=d"(l[2]), "=d"(l[3]) : "l"(line));
      return;
    }
    if ((reinterpret_cast<uintptr_t>(line) & 15) == 0) {  // 16-byte aligned vector: two 128-bit loads
      const double2 lo = __ldg(reinterpret_cast<const double2*>(line));
      const double2 hi = __ldg(reinterpret_cast<const double2*>(line) + 1);
      l[0] = lo.x; l[1] = lo.y; l[2] = hi.x; l[3] = hi.y;
      return;
    }
  }
#pragma unroll
  for (int m = 0; m < N; m++) l[m] = __ldg(line + m * stride);
}

// pm / nm: boundary mode of the low / high end (0 if the pencil's end is not on a brick face); prev / next: this thread's
// DoF line in the element before / after the pencil; glo / ghi: this thread's (der, val) pair in the ghost buffers
template <int N>
__device__ __forceinline__ HaloRaw<N> halo_issue(int pm, int nm, const double* __restrict__ prev, const double* __restrict__ next,
                                                 long stride, const double* __restrict__ glo, const double* __restrict__ ghi) {
  HaloRaw<N> r;
  // (the lines are only read in the modes that load them, but leaving them undefined costs the persistent Q3 kernel 200 bytes of
  // spills: ptxas then keeps the undefined live ranges apart)
#pragma unroll
  for (int m = 0; m < N; m++) r.p[m] = r.n[m] = 0.0;
  r.pm = pm; r.nm = nm;
  if (pm == 0) halo_load_line<N>(r.p, prev, stride);
  else if (pm == 3) { r.p[0] = __ldcg(glo); r.p[1] = __ldcg(glo + 1); }
  if (nm == 0) halo_load_line<N>(r.n, next, stride);
  else if (nm == 3) { r.n[0] = __ldcg(ghi); r.n[1] = __ldcg(ghi + 1); }
  return r;
}

struct HaloTrace { double pd, pv, nd, nv; int pm, nm; };  // pm / nm in {0 use the trace, 1 Dirichlet, 2 natural}
template <int N>
__device__ __forceinline__ HaloTrace halo_reduce(const UniParams<N>& P, const HaloRaw<N>& r) {
  HaloTrace t; t.pd = t.pv = t.nd = t.nv = 0; t.pm = r.pm; t.nm = r.nm;
  if (r.pm == 0) {  // the element before the pencil, at its far (s = 1) side
    double d = P.g[1][0] * r.p[0];
#pragma unroll
    for (int m = 1; m < N; m++) d = fma(P.g[1][m], r.p[m], d);
    t.pd = d; t.pv = r.p[N - 1];
  } else if (r.pm == 3) { t.pd = r.p[0]; t.pv = r.p[1]; t.pm = 0; }
  if (r.nm == 0) {  // the element after the pencil, at its near (s = 0) side
    double d = P.g[0][0] * r.n[0];
#pragma unroll
    for (int m = 1; m < N; m++) d = fma(P.g[0][m], r.n[m], d);
    t.nd = d; t.nv = r.n[0];
  } else if (r.nm == 3) { t.nd = r.n[0]; t.nv = r.n[1]; t.nm = 0; }
  return t;
}

}  // namespace hpdg
