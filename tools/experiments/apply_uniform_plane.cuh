// NOT BUILT INTO THE LIBRARY -- kept as the record of a round-2 experiment.
// Parity-green on a B200 against the oracle (41 tests incl. accumulate mode and multi-tile CTAs), but slower than the pencil kernel
// it was meant to replace: 64^3 Q4 338.6 us against 246.4 us (k_apply_uniform<5,3,3,3>), Q2 73.4 against 66.1 us.  ncu: 112.8 M warp
// instructions against 148.0 M, but 168 registers -> 2 CTAs = 10 warps per SM, issue-active 30 %; the boundary lanes' halo loads
// (25 scattered loads per outer side, 25-50 % lane utilisation, L2 latency exposed three times per tile while the other lanes
// wait at the shuffles' WARPSYNC) cost 128 us: the same kernel with the halo loads removed (wrong results) runs 210.9 us.
// A first version with the tables read per line (3 CTAs/SM, 128 registers) spent 35 M instructions on LDC + R2UR: 411.8 us.
//
// Persistent "plane" tile kernel of the uniform-degree 3-D SIPG operator apply for odd N = p + 1 (Q4: N = 5, Q2: N = 3): the kernel
// of the V-cycle / weak-scaling configurations (cfg4, cfg5).
//
// Same operator and formulation as k_apply_uniform (apply_uniform.cu; reference: Operator::apply over IPDGOperator,
// matrix-free/operator.hh:41-56, matrix-free/localoperators/ipdgoperator.hh:80-390):
//        y = factor * (M_x M_y M_z) (Tt_x + Tt_y + Tt_z) u ,
// but organised around DoF PLANES instead of pencils:
//   tile = 4 x 4 x 2 elements, N warps; lane = element (ex = lane & 3, ey = (lane >> 2) & 3, ez = lane >> 4), warp = plane index.
//   A  (x,y)-plane k = warp of the lane's element: w = Tt_x u + Tt_y u           (u from shared memory line by line, w in registers)
//   B  (x,z)-plane j = warp:                       w = factor M_x M_z (w + Tt_z u)
//   C  (x,y)-plane k = warp:                       w = M_y w, back in place; the tile leaves by one bulk store per element row
// * In-tile face fluxes travel by warp shuffles: the x / y / z neighbour of an element is lane +-1 / +-4 / ^16 of the same warp,
//   so the (der, val) trace of a DoF line is handed to the neighbouring element's lane without touching shared memory.  Traces of
//   the elements outside the tile are read by the boundary lanes from global / L2 (one DoF line per face node) or, on a rank
//   boundary, from the ghost trace buffers; domain boundaries are synthetic traces (uniform_common.cuh).
// * Layout: the 8 rows of four x-contiguous elements of a tile are contiguous in a DynamicBlockVector (4 N^3 doubles: 4 000 B at Q4,
//   864 B at Q2 -- multiples of 16, which a single element is not) and arrive by one bulk async copy each, unpadded, a tile ahead.
//   With lane = element every shared-memory access of a warp has ONE intra-element offset, and the element stride N^3 = 125 (27) is
//   13 (11) mod 16 with the row stride 4 mod 16 (12 mod 16): the 16 lanes of a half warp always hit 16 distinct banks.  All
//   intra-element offsets are compile-time immediates: no index arithmetic in the stages.
// * Shared-memory accesses per DoF: 3 + 3 + 2 = 8 (+ bulk in / out), against 11 of the pencil kernel; 4 block barriers per tile.
// Requires brick extents that are multiples of (4, 4, 2) and 16-byte aligned vectors (otherwise k_apply_uniform is used).
#pragma once
#include <cstddef>
#include <cstdint>

#include "q3p_common.cuh"
#include "uniform_common.cuh"

#ifndef PLANE_MINB5
#define PLANE_MINB5 2
#endif

namespace hpdg {

template <int N> struct PlaneCfg {
  static constexpr int N2 = N * N, N3 = N * N * N, TE = 32, THREADS = 32 * N, ROW = 4 * N3, TILE = TE * N3;
  static constexpr int SMEM = 2 * TILE * 8 + 32;   // u tile, w tile, mbarrier, next-tile slot
  static constexpr int MINB = N == 5 ? PLANE_MINB5 : 5;
};

// ---- table access ------------------------------------------------------------------------------------------------------------
// As in the persistent Q3 kernel (apply_uniform_q3p.cuh): read through the parameter struct, the ~150 table entries a stage uses are
// hoisted, overflow the uniform register file and come back as LDC + R2UR + local-memory traffic (measured: 35 M of 158 M warp
// instructions at 64^3 Q4).  A volatile immediate-offset ld.param at the point of use keeps them rematerialisable constant-bank
// operands.  TAB names the kernel whose parameter block is read.
struct PlaneTabQ4 {
  template <int OFF> static __device__ __forceinline__ double c() {
    double v;
    asm volatile("ld.param.f64 %0, [hpdg_k_apply_plane_q4_param_0+%1];\n" : "=d"(v) : "n"(OFF));
    return v;
  }
};
struct PlaneTabQ2 {
  template <int OFF> static __device__ __forceinline__ double c() {
    double v;
    asm volatile("ld.param.f64 %0, [hpdg_k_apply_plane_q2_param_0+%1];\n" : "=d"(v) : "n"(OFF));
    return v;
  }
};
#define PLANE_C(field, idx) TAB::template c<(int)offsetof(UniParams<N>, field) + 8 * (idx)>()

// The stages work on a whole DoF plane u[a][b] held in registers, with every sweep in COEFFICIENT-STATIONARY order: a table entry
// is read once per plane and feeds N FMAs (one per line), instead of once per line.  ROWS: the sweep runs along the second index
// (lines u[l][.]), else along the first (lines u[.][l]).
template <int N, bool ROWS> __device__ __forceinline__ double& plane_at(double (&u)[N][N], int l, int m) { return ROWS ? u[l][m] : u[m][l]; }
template <int N, bool ROWS> __device__ __forceinline__ double plane_at(const double (&u)[N][N], int l, int m) { return ROWS ? u[l][m] : u[m][l]; }

// (der at side 0, der at side 1) of the N lines; g_1 = -R g_0 (symmetric GL nodes)
template <int N, class TAB, bool ROWS>
__device__ __forceinline__ void plane_traces(const double (&u)[N][N], double (&d0)[N], double (&d1)[N]) {
#pragma unroll
  for (int l = 0; l < N; l++) d0[l] = d1[l] = 0.0;
  q3p_for<N>([&](auto mc) {
    constexpr int m = decltype(mc)::value;
    const double g = PLANE_C(g, m);
#pragma unroll
    for (int l = 0; l < N; l++) { d0[l] = fma(g, plane_at<N, ROWS>(u, l, m), d0[l]); d1[l] = fma(-g, plane_at<N, ROWS>(u, l, N - 1 - m), d1[l]); }
  });
}
// acc += Dp_dir u along the lines (the element-local part of Tt_dir)
template <int N, class TAB, int DIR, bool ROWS>
__device__ __forceinline__ void plane_volume(const double (&u)[N][N], double (&acc)[N][N]) {
  q3p_for<N * N>([&](auto qc) {
    constexpr int i = decltype(qc)::value / N, m = decltype(qc)::value % N;
    const double c = PLANE_C(Dp, DIR * N * N + i * N + m);
#pragma unroll
    for (int l = 0; l < N; l++) plane_at<N, ROWS>(acc, l, i) = fma(c, plane_at<N, ROWS>(u, l, m), plane_at<N, ROWS>(acc, l, i));
  });
}
// acc += response to the traces (qd, qv) of the element before and (rd, rv) of the element after every line
template <int N, class TAB, int DIR, bool ROWS>
__device__ __forceinline__ void plane_response(const double (&qd)[N], const double (&qv)[N], const double (&rd)[N], const double (&rv)[N],
                                               double (&acc)[N][N]) {
  q3p_for<N>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    const double a0 = PLANE_C(A0, DIR * N + i), b0 = PLANE_C(B0, DIR * N + i), a1 = PLANE_C(A1, DIR * N + i), b1 = PLANE_C(B1, DIR * N + i);
#pragma unroll
    for (int l = 0; l < N; l++) {
      double s = plane_at<N, ROWS>(acc, l, i);
      s = fma(a0, qd[l], s); s = fma(b0, qv[l], s); s = fma(a1, rd[l], s); s = fma(b1, rv[l], s);
      plane_at<N, ROWS>(acc, l, i) = s;
    }
  });
}
// o = M a (or factor * M a) along the lines
template <int N, class TAB, bool SCALED, bool ROWS>
__device__ __forceinline__ void plane_mass(const double (&a)[N][N], double (&o)[N][N]) {
#pragma unroll
  for (int l = 0; l < N; l++)
#pragma unroll
    for (int i = 0; i < N; i++) o[l][i] = 0.0;
  q3p_for<N * N>([&](auto qc) {
    constexpr int i = decltype(qc)::value / N, m = decltype(qc)::value % N;
    const double c = SCALED ? PLANE_C(Mf, i * N + m) : PLANE_C(M, i * N + m);
#pragma unroll
    for (int l = 0; l < N; l++) plane_at<N, ROWS>(o, l, i) = fma(c, plane_at<N, ROWS>(a, l, m), plane_at<N, ROWS>(o, l, i));
  });
}

// Traces (der, val) of the element OUTSIDE the tile across the lane's outer side in one direction, for the N DoF lines of the lane's
// plane.  side: 0 the outside element lies before the lane's element, 1 after it.  mode: 0 it is an element of this brick (its DoF
// lines are read from global memory: `line0` = first node of line 0 in the neighbour, `lstep` between lines, `nstep` between the
// nodes of a line), 3 it lives on another rank (`gh` = the lane's first (der, val) pair in the ghost buffer, `gstep` pairs
// between lines).  Modes 1 / 2 (domain boundary) need the lane's own traces and are formed where those are known.
// Reading the line from the shared face outwards makes both sides the same code: der = +- sum g_0[m] h[m], val = h[0].
template <int N, class TAB>
__device__ __forceinline__ void plane_halo(int side, int mode, const double* __restrict__ line0, long lstep,
                                           long nstep, const double* __restrict__ gh, long gstep, double (&hd)[N], double (&hv)[N]) {
  if (mode == 0) {
    const double* base = side == 0 ? line0 + (N - 1) * nstep : line0;   // node of the line on the shared face
    const long st = side == 0 ? -nstep : nstep;
    const double sg = side == 0 ? -1.0 : 1.0;
    double h[N][N];
#pragma unroll
    for (int l = 0; l < N; l++)
#pragma unroll
      for (int m = 0; m < N; m++) h[l][m] = __ldg(base + l * lstep + m * st);
#pragma unroll
    for (int l = 0; l < N; l++) {
      double d = 0;
      q3p_for<N>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        d = fma(PLANE_C(g, m), h[l][m], d);
      });
      hd[l] = sg * d; hv[l] = h[l][0];
    }
  } else if (mode == 3) {
#pragma unroll
    for (int l = 0; l < N; l++) { hd[l] = __ldcg(gh + 2 * l * gstep); hv[l] = __ldcg(gh + 2 * l * gstep + 1); }
  }
}
// the outer trace of a line whose element sits on the tile's outer side: from the halo (modes 0, 3) or synthetic from the line's own
// trace (de, ve) at that end (1 Dirichlet, 2 natural; uniform_common.cuh: pencil_apply)
template <int N, class TAB, int DIR>
__device__ __forceinline__ void plane_outer(int side, int mode, double hd, double hv, double de, double ve, double& od, double& ov) {
  if (mode == 1) { od = fma(side == 0 ? -PLANE_C(cohk, DIR) : PLANE_C(cohk, DIR), ve, de); ov = -ve; }
  else if (mode == 2) { od = -de; ov = ve; }
  else { od = hd; ov = hv; }
}

template <int N, class TAB>
__device__ __forceinline__ void plane_kernel_body(const UniParams<N>& P, const int4* __restrict__ tile_desc, const int ntiles,
                                                  const int ntiles_total, int* __restrict__ sched) {
  using C = PlaneCfg<N>;
  constexpr int N2 = C::N2, N3 = C::N3;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) double pl_sm[];
  double* __restrict__ U = pl_sm;
  double* __restrict__ W = pl_sm + C::TILE;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(pl_sm + 2 * C::TILE);
  volatile int* s_next = reinterpret_cast<volatile int*>(pl_sm + 2 * C::TILE + 1);
  const double* __restrict__ X = P.x;
  const int n0 = P.n[0], n01 = P.n[0] * P.n[1];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int ex = lane & 3, ey = (lane >> 2) & 3, ez = lane >> 4;

  auto descriptor = [&](int t) {
    int tb = P.tile_list ? P.tile_list[t] : t + P.tile_offset;
    if (P.tile_rot) { tb += P.tile_rot; if (tb >= ntiles_total) tb -= ntiles_total; }
    return __ldg(tile_desc + tb);
  };
  auto prefetch = [&](int e0) {   // one thread: the tile's 8 element rows
    q3p_mbar_expect_tx(mbar, (uint32_t)(C::TILE * 8));
#pragma unroll
    for (int row = 0; row < 8; row++)
      q3p_bulk_g2s(U + C::ROW * row, X + (long)(e0 + n0 * (row & 3) + n01 * (row >> 2)) * N3, (uint32_t)(C::ROW * 8), mbar);
  };

  if (tid == 0) {
    q3p_mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  int t = blockIdx.x;
  if (t >= ntiles) return;
  int4 td = descriptor(t);
  if (tid == 0) prefetch(td.x);
  uint32_t phase = 0;
  bool stored = false;   // a bulk store of the w tile is in flight (thread 0 waits for its reads before w is written again)

  for (;;) {
    if (tid == 0) *s_next = (int)gridDim.x + atomicAdd(sched, 1);
    const int e0 = td.x, fl = td.z;
    const int tx4 = (td.y & 1023) * 4, ty4 = ((td.y >> 10) & 1023) * 4, tz2 = (td.y >> 20) * 2;
    if (P.ghost_step > 0) {  // p2p halo: tiles on a rank boundary wait until the neighbour's traces for this step have arrived
      bool any = false;
#pragma unroll
      for (int f = 0; f < 6; f++) any = any || (((fl >> f) & 1) && P.bmode[f] == 3);
      if (any) {
        if (tid == 0) {
          const long long tstart = clock64();
          for (int f = 0; f < 6; f++) {
            if (!(((fl >> f) & 1) && P.bmode[f] == 3)) continue;
            const volatile int* fg = P.ghost_flag[f];
            while (*fg < P.ghost_step) {
              __nanosleep(200);
              if (*reinterpret_cast<volatile int*>(P.ghost_err)) break;
              if (clock64() - tstart > P.ghost_timeout) {
                atomicExch(P.ghost_err, 1); *reinterpret_cast<volatile int*>(P.ghost_err_host) = 1; __threadfence_system(); break;
              }
            }
          }
          __threadfence();
        }
        __syncthreads();
      }
    }
    const long eg = (long)e0 + ex + n0 * ey + n01 * ez;      // the lane's element
    const double* __restrict__ Xe = X + eg * N3;
    // outer sides of the lane's element in the tile: -1 none, else the boundary mode of that side (0: an element of this brick)
    const int xs = ex == 0 ? 0 : 1, ys = ey == 0 ? 0 : 1, zs = ez;
#ifdef PLANE_NOHALO
#define xm xm_
#define ym ym_
#define zm zm_
#endif
    const int xm = (ex == 0 || ex == 3) ? (((fl >> xs) & 1) ? P.bmode[xs] : 0) : -1;
    const int ym = (ey == 0 || ey == 3) ? (((fl >> (2 + ys)) & 1) ? P.bmode[2 + ys] : 0) : -1;
    const int zm = ((fl >> (4 + zs)) & 1) ? P.bmode[4 + zs] : 0;
#ifdef PLANE_NOHALO   // timing experiment only (wrong results): no halo loads
#undef xm
#undef ym
#undef zm
    const int xm = xm_ == 0 ? 2 : xm_, ym = ym_ == 0 ? 2 : ym_, zm = zm_ == 0 ? 2 : zm_;
#endif

    // ---------------- A: (x,y)-plane k = wp: w = Tt_x u + Tt_y u ----------------
    {
      double hxd[N], hxv[N], hyd[N], hyv[N];
#pragma unroll
      for (int l = 0; l < N; l++) hxd[l] = hxv[l] = hyd[l] = hyv[l] = 0.0;
      if (xm == 0 || xm == 3)   // lines j = 0..N-1 of plane k in the x-neighbour; ghost face nodes (j, k) of face element (y, z)
        plane_halo<N, TAB>(xs, xm, Xe + (xs ? N3 : -N3) + N2 * wp, N, 1,
                      P.ghost[xs] + (((long)(ty4 + ey) + (long)P.n[1] * (tz2 + ez)) * N2 + N * wp) * 2, 1, hxd, hxv);
      if (ym == 0 || ym == 3)   // lines i = 0..N-1 of plane k in the y-neighbour; ghost face nodes (i, k) of face element (x, z)
        plane_halo<N, TAB>(ys, ym, Xe + (ys ? (long)n0 * N3 : -(long)n0 * N3) + N2 * wp, 1, N,
                      P.ghost[2 + ys] + (((long)(tx4 + ex) + (long)n0 * (tz2 + ez)) * N2 + N * wp) * 2, 1, hyd, hyv);
      while (!q3p_mbar_try_wait(mbar, phase)) {}
      phase ^= 1;
      const double* __restrict__ up = U + N3 * lane + N2 * wp;
      double u[N][N], acc[N][N];   // [j][i]
#pragma unroll
      for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = 0; i < N; i++) { u[j][i] = up[N * j + i]; acc[j][i] = 0.0; }
      {   // x lines u[j][.]
        double d0[N], d1[N], qd[N], qv[N], rd[N], rv[N];
        plane_traces<N, TAB, true>(u, d0, d1);
        plane_volume<N, TAB, 0, true>(u, acc);
#pragma unroll
        for (int j = 0; j < N; j++) {
          qd[j] = __shfl_up_sync(FULL, d1[j], 1); qv[j] = __shfl_up_sync(FULL, u[j][N - 1], 1);
          rd[j] = __shfl_down_sync(FULL, d0[j], 1); rv[j] = __shfl_down_sync(FULL, u[j][0], 1);
          if (ex == 0) plane_outer<N, TAB, 0>(0, xm, hxd[j], hxv[j], d0[j], u[j][0], qd[j], qv[j]);
          if (ex == 3) plane_outer<N, TAB, 0>(1, xm, hxd[j], hxv[j], d1[j], u[j][N - 1], rd[j], rv[j]);
        }
        plane_response<N, TAB, 0, true>(qd, qv, rd, rv, acc);
      }
      {   // y lines u[.][i]
        double d0[N], d1[N], qd[N], qv[N], rd[N], rv[N];
        plane_traces<N, TAB, false>(u, d0, d1);
        plane_volume<N, TAB, 1, false>(u, acc);
#pragma unroll
        for (int i = 0; i < N; i++) {
          qd[i] = __shfl_up_sync(FULL, d1[i], 4); qv[i] = __shfl_up_sync(FULL, u[N - 1][i], 4);
          rd[i] = __shfl_down_sync(FULL, d0[i], 4); rv[i] = __shfl_down_sync(FULL, u[0][i], 4);
          if (ey == 0) plane_outer<N, TAB, 1>(0, ym, hyd[i], hyv[i], d0[i], u[0][i], qd[i], qv[i]);
          if (ey == 3) plane_outer<N, TAB, 1>(1, ym, hyd[i], hyv[i], d1[i], u[N - 1][i], rd[i], rv[i]);
        }
        plane_response<N, TAB, 1, false>(qd, qv, rd, rv, acc);
      }
      if (stored) {   // the previous tile's bulk store must have read w before it is overwritten
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        __syncthreads();
      }
      double* __restrict__ wq = W + N3 * lane + N2 * wp;
#pragma unroll
      for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = 0; i < N; i++) wq[N * j + i] = acc[j][i];
    }
    __syncthreads();

    // ---------------- B: (x,z)-plane j = wp: w = factor * M_x M_z (w + Tt_z u) ----------------
    {
      double hzd[N], hzv[N];
#pragma unroll
      for (int l = 0; l < N; l++) hzd[l] = hzv[l] = 0.0;
      if (zm == 0 || zm == 3)   // lines i = 0..N-1 (node (i, j = wp)) of the z-neighbour; ghost face nodes (i, j) of face element (x, y)
        plane_halo<N, TAB>(zs, zm, Xe + (zs ? (long)n01 * N3 : -(long)n01 * N3) + N * wp, 1, N2,
                      P.ghost[4 + zs] + (((long)(tx4 + ex) + (long)n0 * (ty4 + ey)) * N2 + N * wp) * 2, 1, hzd, hzv);
      const double* __restrict__ up = U + N3 * lane + N * wp;
      double* __restrict__ wq = W + N3 * lane + N * wp;
      double u[N][N], acc[N][N], o[N][N];   // [k][i]
#pragma unroll
      for (int k = 0; k < N; k++)
#pragma unroll
        for (int i = 0; i < N; i++) { u[k][i] = up[N2 * k + i]; acc[k][i] = wq[N2 * k + i]; }
      {   // z lines u[.][i]
        double d0[N], d1[N], qd[N], qv[N], rd[N], rv[N];
        plane_traces<N, TAB, false>(u, d0, d1);
        plane_volume<N, TAB, 2, false>(u, acc);
#pragma unroll
        for (int i = 0; i < N; i++) {
          // the in-tile z-neighbour is lane ^ 16: it needs this element's trace at the shared face
          const double sd = __shfl_xor_sync(FULL, ez == 0 ? d1[i] : d0[i], 16);
          const double sv = __shfl_xor_sync(FULL, ez == 0 ? u[N - 1][i] : u[0][i], 16);
          if (ez == 0) { rd[i] = sd; rv[i] = sv; plane_outer<N, TAB, 2>(0, zm, hzd[i], hzv[i], d0[i], u[0][i], qd[i], qv[i]); }
          else { qd[i] = sd; qv[i] = sv; plane_outer<N, TAB, 2>(1, zm, hzd[i], hzv[i], d1[i], u[N - 1][i], rd[i], rv[i]); }
        }
        plane_response<N, TAB, 2, false>(qd, qv, rd, rv, acc);
      }
      plane_mass<N, TAB, false, false>(acc, o);   // M_z
      plane_mass<N, TAB, true, true>(o, acc);     // factor * M_x
#pragma unroll
      for (int k = 0; k < N; k++)
#pragma unroll
        for (int i = 0; i < N; i++) wq[N2 * k + i] = acc[k][i];
    }
    __syncthreads();

    // the u tile is free: start the next tile's copies; they land during stage C and the next tile's halo loads
    const int tn = *s_next;
    const bool has_next = tn < ntiles;
    if (has_next) {
      td = descriptor(tn);
      if (tid == 0) { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); prefetch(td.x); }
    }

    // ---------------- C: (x,y)-plane k = wp: w = M_y w; bulk store ----------------
    {
      double* __restrict__ wq = W + N3 * lane + N2 * wp;
      double a[N][N], o[N][N];   // [j][i]
#pragma unroll
      for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = 0; i < N; i++) a[j][i] = wq[N * j + i];
      plane_mass<N, TAB, false, false>(a, o);
#pragma unroll
      for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = 0; i < N; i++) wq[N * j + i] = o[j][i];
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the stage's writes before the bulk store's reads of w
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int row = 0; row < 8; row++) {
        double* dst = P.y + (long)(e0 + n0 * (row & 3) + n01 * (row >> 2)) * N3;
        const uint32_t src = q3p_smem_u32(W + C::ROW * row);
        if (P.accum)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"((uint32_t)(C::ROW * 8)) : "memory");
        else
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"((uint32_t)(C::ROW * 8)) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    }
    stored = true;
    if (!has_next) break;
    t = tn;
  }
  if (tid == 0) {
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // shared memory stays valid until the last store has left
    if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; __threadfence(); }
  }
}

#undef PLANE_C

}  // namespace hpdg

// extern "C": the table reads name the kernels' parameter symbols (<kernel>_param_0)
extern "C" __global__ void __launch_bounds__(hpdg::PlaneCfg<5>::THREADS, hpdg::PlaneCfg<5>::MINB)
hpdg_k_apply_plane_q4(const __grid_constant__ hpdg::UniParams<5> P, const int4* __restrict__ tile_desc, const int ntiles,
                      const int ntiles_total, int* __restrict__ sched) {
  hpdg::plane_kernel_body<5, hpdg::PlaneTabQ4>(P, tile_desc, ntiles, ntiles_total, sched);
}
extern "C" __global__ void __launch_bounds__(hpdg::PlaneCfg<3>::THREADS, hpdg::PlaneCfg<3>::MINB)
hpdg_k_apply_plane_q2(const __grid_constant__ hpdg::UniParams<3> P, const int4* __restrict__ tile_desc, const int ntiles,
                      const int ntiles_total, int* __restrict__ sched) {
  hpdg::plane_kernel_body<3, hpdg::PlaneTabQ2>(P, tile_desc, ntiles, ntiles_total, sched);
}
