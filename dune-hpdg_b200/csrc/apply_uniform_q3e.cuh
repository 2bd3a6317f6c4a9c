// Even/odd variant of the persistent Q3 tile kernel (apply_uniform_q3p.cuh): same operator (Operator::apply over
// IPDGOperator, matrix-free/operator.hh:41-56, matrix-free/localoperators/ipdgoperator.hh:80-390), same tiles, same shared
// memory layout, same bulk-copy prefetch, but the arithmetic between the first and the last pass runs in the even/odd
// basis of the GL nodes.
//
// The GL nodes are symmetric about the element centre.  With R the node reflection every 1-D table of the operator is
// (anti)symmetric under R:  M R = R M,  Dp R = R Dp,  g_1 = -R g_0,  A1 = -R A0,  B1 = R B0.  In the unnormalised even/odd
// coordinates of a DoF line  E v = (v0+v3, v1+v2, v0-v3, v1-v2)  a 4x4 table that commutes with R splits into two 2x2
// blocks, the own-trace pair costs 4 multiplies + 2 adds instead of 8 FMAs, and the response to the neighbour traces needs
// 2 FMAs per node (on the sums / differences of the two sides' traces) instead of 4:  a T-sweep is 28 FP64 operations per
// line instead of 40 and a mass sweep 8 instead of 16.  Each direction's transform is applied once (4 adds per line) when
// the pass of that direction first reads u / w and undone (4 adds) after that direction's mass sweep; the three factors
// of 2 of the back transform are folded into the last mass table.  B200 issues a DFMA every other cycle per scheduler, so
// FP64 instructions are the dominant issue cost of this kernel: 35 instead of 43.75 per DoF.
//
// The traces of the elements outside the tile are read from global memory in nodal form and have to be brought to the
// representation of the pass that consumes them: the x-pass needs them even/odd in z, the y-pass even/odd in z and x.
// The thread that owns slot e_a (o_a) of a direction loads the nodal line a (3-a); one shuffle with the partner lane and
// one add per transformed direction gives both their values.
#pragma once
#include "apply_uniform_q3p.cuh"

namespace hpdg {

struct Q3eTab {
  double De[3][4], Do[3][4];  // even / odd 2x2 blocks of Dp, row-major
  double Ae[3][2], Ao[3][2];  // A0[a] +- A0[3-a]
  double Be[3][2], Bo[3][2];  // (B0[a] +- B0[3-a]) / 2   (value traces are carried doubled)
  double ge[2], go[2];        // (g0[b] +- g0[3-b]) / 2
  double hc[3];               // cohk / 2
  double Me[4], Mo[4];        // even / odd blocks of the mass
  double Mfe[4], Mfo[4];      // ... times factor / 8
  double g0[4];               // nodal derivative trace at side 0 (outside elements)
};

template <int OFF>
__device__ __forceinline__ double q3e_c() {
  double v;
  asm volatile("ld.param.f64 %0, [hpdg_k_apply_q3_eo_param_1+%1];\n" : "=d"(v) : "n"(OFF));
  return v;
}
#define Q3E_C(field, idx) q3e_c<(int)offsetof(Q3eTab, field) + 8 * (idx)>()

__device__ __forceinline__ void q3e_fwd(double (&l)[4]) {  // nodal -> (e0, e1, o0, o1)
  const double a = l[0], b = l[1], c = l[2], d = l[3];
  l[0] = a + d; l[1] = b + c; l[2] = a - d; l[3] = b - c;
}
__device__ __forceinline__ void q3e_back(double (&l)[4]) {  // (e0, e1, o0, o1) -> 2 * nodal
  const double e0 = l[0], e1 = l[1], o0 = l[2], o1 = l[3];
  l[0] = e0 + o0; l[1] = e1 + o1; l[2] = e1 - o1; l[3] = e0 - o0;
}
template <bool SCALED>
__device__ __forceinline__ void q3e_mass(double (&l)[4]) {
  const double e0 = l[0], e1 = l[1], o0 = l[2], o1 = l[3];
  if (SCALED) {
    l[0] = fma(Q3E_C(Mfe, 1), e1, Q3E_C(Mfe, 0) * e0); l[1] = fma(Q3E_C(Mfe, 3), e1, Q3E_C(Mfe, 2) * e0);
    l[2] = fma(Q3E_C(Mfo, 1), o1, Q3E_C(Mfo, 0) * o0); l[3] = fma(Q3E_C(Mfo, 3), o1, Q3E_C(Mfo, 2) * o0);
  } else {
    l[0] = fma(Q3E_C(Me, 1), e1, Q3E_C(Me, 0) * e0); l[1] = fma(Q3E_C(Me, 3), e1, Q3E_C(Me, 2) * e0);
    l[2] = fma(Q3E_C(Mo, 1), o1, Q3E_C(Mo, 0) * o0); l[3] = fma(Q3E_C(Mo, 3), o1, Q3E_C(Mo, 2) * o0);
  }
}

// Traces of the neighbours of a pencil: derivative trace and DOUBLED value trace of the element before (p) / after (n).
struct Q3eTrace { double pd, pV, nd, nV; int pm, nm; };

__device__ __forceinline__ void q3e_outside(const double* __restrict__ line, int stride, int side, double& der, double& V) {
  double u[4];
  if (stride == 1) {
    const double2 lo = __ldg(reinterpret_cast<const double2*>(line));
    const double2 hi = __ldg(reinterpret_cast<const double2*>(line) + 1);
    u[0] = lo.x; u[1] = lo.y; u[2] = hi.x; u[3] = hi.y;
  } else {
#pragma unroll
    for (int m = 0; m < 4; m++) u[m] = __ldg(line + m * stride);
  }
  if (side == 0) {
    der = fma(Q3E_C(g0, 0), u[0], fma(Q3E_C(g0, 1), u[1], fma(Q3E_C(g0, 2), u[2], Q3E_C(g0, 3) * u[3])));
    V = u[0] + u[0];
  } else {  // g_1 = -R g_0
    der = -fma(Q3E_C(g0, 3), u[0], fma(Q3E_C(g0, 2), u[1], fma(Q3E_C(g0, 1), u[2], Q3E_C(g0, 0) * u[3])));
    V = u[3] + u[3];
  }
}

template <int DIR, class GIdx>
__device__ __forceinline__ Q3eTrace q3e_halo(const UniParams<4>& P, int fl, const double* __restrict__ prev,
                                             const double* __restrict__ next, int stride, GIdx gidx) {
  Q3eTrace r; r.pd = r.pV = r.nd = r.nV = 0;
  r.pm = (fl >> (2 * DIR)) & 1 ? P.bmode[2 * DIR] : 0;
  r.nm = (fl >> (2 * DIR + 1)) & 1 ? P.bmode[2 * DIR + 1] : 0;
  if (r.pm == 0) q3e_outside(prev, stride, 1, r.pd, r.pV);
  else if (r.pm == 3) { const double* gp = P.ghost[2 * DIR] + gidx() * 2; r.pd = __ldcg(gp); const double v = __ldcg(gp + 1); r.pV = v + v; r.pm = 0; }
  if (r.nm == 0) q3e_outside(next, stride, 0, r.nd, r.nV);
  else if (r.nm == 3) { const double* gp = P.ghost[2 * DIR + 1] + gidx() * 2; r.nd = __ldcg(gp); const double v = __ldcg(gp + 1); r.nV = v + v; r.nm = 0; }
  return r;
}

// even/odd transform of the four trace values across the two lanes that own slots a and a+2 of a direction:
// slot < 2 (even): mine + partner's, slot >= 2 (odd): partner's - mine.  sg = +1 / -1 accordingly.
__device__ __forceinline__ void q3e_halo_eo(Q3eTrace& h, int lane_mask, double sg) {
  h.pd = fma(h.pd, sg, __shfl_xor_sync(0xffffffffu, h.pd, lane_mask));
  h.pV = fma(h.pV, sg, __shfl_xor_sync(0xffffffffu, h.pV, lane_mask));
  h.nd = fma(h.nd, sg, __shfl_xor_sync(0xffffffffu, h.nd, lane_mask));
  h.nV = fma(h.nV, sg, __shfl_xor_sync(0xffffffffu, h.nV, lane_mask));
}

// acc_e = accin_e + (Tt_dir v)_e along a full pencil, everything in even/odd coordinates of the pencil's direction.
// load(e, l): DoF line of element e as (e0, e1, o0, o1); accin(e, a): initial accumulator; out(e, l, a).  Elements are
// processed in the order 1, 2, 3, 0 so that the outside traces (global loads issued just before) are consumed last.
template <int DIR, class Load, class AccIn, class Out>
__device__ __forceinline__ void q3e_pencil(Q3eTrace h, Load load, AccIn accin, Out out) {
  constexpr int T = 4;
  double v[T][4], d0[T], d1[T], V0[T], V3[T];
#pragma unroll
  for (int e = 0; e < T; e++) {
    load(e, v[e]);
    const double se = fma(Q3E_C(ge, 1), v[e][1], Q3E_C(ge, 0) * v[e][0]);
    const double so = fma(Q3E_C(go, 1), v[e][3], Q3E_C(go, 0) * v[e][2]);
    d0[e] = so + se; d1[e] = so - se;
    V0[e] = v[e][0] + v[e][2]; V3[e] = v[e][0] - v[e][2];
  }
#pragma unroll
  for (int ee = 0; ee < T; ee++) {
    const int e = (ee + 1) % T;
    double qd, qV, rd, rV;
    if (e == 0) {
      if (h.pm == 1) { h.pd = fma(-Q3E_C(hc, DIR), V0[0], d0[0]); h.pV = -V0[0]; }
      else if (h.pm == 2) { h.pd = -d0[0]; h.pV = V0[0]; }
      qd = h.pd; qV = h.pV;
    } else { qd = d1[e > 0 ? e - 1 : 0]; qV = V3[e > 0 ? e - 1 : 0]; }
    if (e == T - 1) {
      if (h.nm == 1) { h.nd = fma(Q3E_C(hc, DIR), V3[T - 1], d1[T - 1]); h.nV = -V3[T - 1]; }
      else if (h.nm == 2) { h.nd = -d1[T - 1]; h.nV = V3[T - 1]; }
      rd = h.nd; rV = h.nV;
    } else { rd = d0[e < T - 1 ? e + 1 : e]; rV = V0[e < T - 1 ? e + 1 : e]; }
    const double cdm = qd - rd, cdp = qd + rd, cvp = qV + rV, cvm = qV - rV;
    double a[4];
    accin(e, a);
    a[0] = fma(Q3E_C(Be, DIR * 2 + 0), cvp, fma(Q3E_C(Ae, DIR * 2 + 0), cdm, fma(Q3E_C(De, DIR * 4 + 1), v[e][1], fma(Q3E_C(De, DIR * 4 + 0), v[e][0], a[0]))));
    a[1] = fma(Q3E_C(Be, DIR * 2 + 1), cvp, fma(Q3E_C(Ae, DIR * 2 + 1), cdm, fma(Q3E_C(De, DIR * 4 + 3), v[e][1], fma(Q3E_C(De, DIR * 4 + 2), v[e][0], a[1]))));
    a[2] = fma(Q3E_C(Bo, DIR * 2 + 0), cvm, fma(Q3E_C(Ao, DIR * 2 + 0), cdp, fma(Q3E_C(Do, DIR * 4 + 1), v[e][3], fma(Q3E_C(Do, DIR * 4 + 0), v[e][2], a[2]))));
    a[3] = fma(Q3E_C(Bo, DIR * 2 + 1), cvm, fma(Q3E_C(Ao, DIR * 2 + 1), cdp, fma(Q3E_C(Do, DIR * 4 + 3), v[e][3], fma(Q3E_C(Do, DIR * 4 + 2), v[e][2], a[3]))));
    out(e, v[e], a);
  }
}

}  // namespace hpdg

extern "C" __global__ void __launch_bounds__(256, 3)
hpdg_k_apply_q3_eo(const __grid_constant__ hpdg::UniParams<4> P, const __grid_constant__ hpdg::Q3eTab ET,
                   const int4* __restrict__ tile_desc, const int ntiles, const int ntiles_total) {
  using namespace hpdg;
  constexpr int N = 4, N2 = 16, N3 = 64;
  extern __shared__ __align__(128) double q3p_sm[];
  double* __restrict__ su = q3p_sm;
  double* __restrict__ sw = q3p_sm + 4096;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(q3p_sm + 8192);
  const double* __restrict__ X = P.x;
  const int n0 = P.n[0], n01 = P.n[0] * P.n[1];  // element strides (in elements) in y and z

  auto descriptor = [&](int t) {
    int tb = P.tile_list ? P.tile_list[t] : t + P.tile_offset;
    if (P.tile_rot) { tb += P.tile_rot; if (tb >= ntiles_total) tb -= ntiles_total; }
    return __ldg(tile_desc + tb);
  };
  auto prefetch = [&](int tid, int e0) {
    const int l = tid & 127, half = tid >> 7;
    if (l < 8) {
      if (l == 0) q3p_mbar_expect_tx(mbar, 16384u);
      const int ey = l & 3, ez = 2 * half + (l >> 2);
      q3p_bulk_g2s(su + (4 * ey + 16 * ez) * N3, X + (long)(e0 + n0 * ey + n01 * ez) * N3, 2048u, mbar);
    }
  };

  if (threadIdx.x == 0) {
    q3p_mbar_init(mbar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  int t = blockIdx.x;
  if (t >= ntiles) return;
  int4 td = descriptor(t);
  prefetch(threadIdx.x, td.x);
  uint32_t phase = 0;

  for (;;) {
    const int e0 = td.x, fl = td.z;
    const int ty4 = ((td.y >> 10) & 1023) * 4, tz4 = (td.y >> 20) * 4, tx4 = (td.y & 1023) * 4;  // only used on ghost faces
    if (P.ghost_step > 0) {  // p2p halo: tiles on a rank boundary wait until the neighbour's traces for this step have arrived
      bool touch[6]; bool any = false;
#pragma unroll
      for (int f = 0; f < 6; f++) { touch[f] = ((fl >> f) & 1) && P.bmode[f] == 3; any = any || touch[f]; }
      if (any) {
        if (threadIdx.x == 0) {
          const long long tstart = clock64();
          for (int f = 0; f < 6; f++) {
            if (!touch[f]) continue;
            const volatile int* fg = P.ghost_flag[f];
            while (*fg < P.ghost_step) {
              __nanosleep(200);
              if (clock64() - tstart > 4000000000LL) { atomicExch(P.ghost_err, 1); break; }  // ~2 s: give up, never hang the GPU
            }
          }
          __threadfence();
        }
        __syncthreads();
      }
    }

    // ---------------- P1: z-pencils; nodal u from the prefetched tile -> even/odd in z, w = Tt_z u ----------------
    // z-role: node (i, j) of element column (ex, ey); a half warp = one column
    {
      const int tid = q3p_tid();
      const int zq = tid & 15, zex = (tid >> 4) & 3, zey = tid >> 6;
      const int zcol = (zex + 4 * zey) * N3;
      const double* colp = X + (long)(e0 + zex + n0 * zey) * N3 + zq;  // element (x, y, z0), this node
      const Q3eTrace h = q3e_halo<2>(P, fl, colp - (long)n01 * N3, colp + (long)n01 * (4 * N3), N2,
                                     [&]() { return ((long)(tx4 + zex) + (long)n0 * (ty4 + zey)) * N2 + zq; });
      while (!q3p_mbar_try_wait(mbar, phase)) {}
      phase ^= 1;
      // in-place rewrite: every lane of the half warp reads its raw lines before any lane overwrites the column
      double v[4][4];
#pragma unroll
      for (int e = 0; e < 4; e++)
#pragma unroll
        for (int k = 0; k < 4; k++) v[e][k] = su[zcol + 1024 * e + 16 * k + zq];
      __syncwarp();
      q3e_pencil<2>(h,
        [&](int e, double (&l)[4]) {
#pragma unroll
          for (int k = 0; k < 4; k++) l[k] = v[e][k];
          q3e_fwd(l);
        },
        [](int, double (&a)[4]) { a[0] = a[1] = a[2] = a[3] = 0.0; },
        [&](int e, const double (&l)[4], const double (&a)[4]) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int o = zcol + 1024 * e + 16 * k + (((zq ^ (4 * k)) + 2 * (e & 1)) & 15);
            su[o] = l[k]; sw[o] = a[k];
          }
        });
    }

    // ---------------- P2: x-pencils (128-bit shared-memory accesses); u, w -> even/odd in x ----------------
    // x-role: line (j, z-slot k) of element row (ey, ez); a quarter warp = 4 j x 2 element layers
    {
      const int tid = q3p_tid();
      const int xj = tid & 3, xez = ((tid >> 2) & 1) | ((tid >> 6) & 2), xk = (tid >> 3) & 3, xey = (tid >> 5) & 3;
      const int xq = 2 * (xj ^ xk) + (xez & 1);
      const int xo0 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * (xq & 7);        // line entries 0,1 (+ 64 e)
      const int xo1 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * ((xq + 1) & 7);  // line entries 2,3
      const int kn = xk < 2 ? xk : 5 - xk;  // the nodal z-plane this thread reads outside the tile
      const double* rowp = X + (long)(e0 + n0 * xey + n01 * xez) * N3 + N * xj + N2 * kn;  // element (x0, y, z), nodal line (j, kn)
      Q3eTrace h = q3e_halo<0>(P, fl, rowp - N3, rowp + 4 * N3, 1,
                               [&]() { return ((long)(ty4 + xey) + (long)P.n[1] * (tz4 + xez)) * N2 + xj + N * kn; });
      q3e_halo_eo(h, 16, xk < 2 ? 1.0 : -1.0);  // z: partner lane holds slot xk ^ 2
      __syncthreads();
      q3e_pencil<0>(h,
        [&](int e, double (&l)[4]) {
          const double2 lo = *reinterpret_cast<const double2*>(su + xo0 + 64 * e);
          const double2 hi = *reinterpret_cast<const double2*>(su + xo1 + 64 * e);
          l[0] = lo.x; l[1] = lo.y; l[2] = hi.x; l[3] = hi.y;
          q3e_fwd(l);
          *reinterpret_cast<double2*>(su + xo0 + 64 * e) = make_double2(l[0], l[1]);  // the y-pass reads u even/odd in z and x
          *reinterpret_cast<double2*>(su + xo1 + 64 * e) = make_double2(l[2], l[3]);
        },
        [&](int e, double (&a)[4]) {
          const double2 lo = *reinterpret_cast<const double2*>(sw + xo0 + 64 * e);
          const double2 hi = *reinterpret_cast<const double2*>(sw + xo1 + 64 * e);
          a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
          q3e_fwd(a);
        },
        [&](int e, const double (&)[4], const double (&a)[4]) {
          *reinterpret_cast<double2*>(sw + xo0 + 64 * e) = make_double2(a[0], a[1]);
          *reinterpret_cast<double2*>(sw + xo1 + 64 * e) = make_double2(a[2], a[3]);
        });
    }

    // ---------------- P3: y-pencils, M_y, back to nodal in y ----------------
    // y-role: line (x-slot i, z-slot k) of element row (ex, ez); a half warp = 4 i x 4 k
    {
      const int tid = q3p_tid();
      const int yi = tid & 3, yk = (tid >> 2) & 3, yex = (tid >> 4) & 3, yez = tid >> 6;
      const int ybase = (yex + 16 * yez) * N3 + 16 * yk;
      const int yr = yi + 2 * (yez & 1);
      auto yo = [&](int j) { return ybase + (((4 * j) ^ (4 * yk)) + yr & 15); };  // entry j of the line (+ 256 e)
      const int in = yi < 2 ? yi : 5 - yi, kn = yk < 2 ? yk : 5 - yk;  // the nodal line this thread reads outside the tile
      const double* colp = X + (long)(e0 + yex + n01 * yez) * N3 + in + N2 * kn;  // element (x, y0, z), nodal line (in, kn)
      Q3eTrace h = q3e_halo<1>(P, fl, colp - (long)n0 * N3, colp + (long)n0 * (4 * N3), N,
                               [&]() { return ((long)(tx4 + yex) + (long)n0 * (tz4 + yez)) * N2 + in + N * kn; });
      q3e_halo_eo(h, 2, yi < 2 ? 1.0 : -1.0);  // x: partner lane holds slot yi ^ 2
      q3e_halo_eo(h, 8, yk < 2 ? 1.0 : -1.0);  // z: partner lane holds slot yk ^ 2
      q3p_bar_half(tid >> 7);
      q3e_pencil<1>(h,
        [&](int e, double (&l)[4]) {
#pragma unroll
          for (int j = 0; j < 4; j++) l[j] = su[yo(j) + 256 * e];
          q3e_fwd(l);
        },
        [&](int e, double (&a)[4]) {
#pragma unroll
          for (int j = 0; j < 4; j++) a[j] = sw[yo(j) + 256 * e];
          q3e_fwd(a);
        },
        [&](int e, const double (&)[4], const double (&a)[4]) {
          double b[4];
#pragma unroll
          for (int j = 0; j < 4; j++) b[j] = a[j];
          q3e_mass<false>(b);
          q3e_back(b);
#pragma unroll
          for (int j = 0; j < 4; j++) sw[yo(j) + 256 * e] = b[j];
        });
    }

    // this half's u rows are free: start the next tile's copies; they land during P4 and P5
    const int tn = t + (int)gridDim.x;
    const bool has_next = tn < ntiles;
    {
      const int tid = q3p_tid();
      if (has_next) td = descriptor(tn);
      q3p_bar_half(tid >> 7);
      if (has_next) {
        if ((tid & 127) < 8) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        prefetch(tid, td.x);
      }
    }

    // ---------------- P4: M_x, back to nodal in x ----------------
    {
      const int tid = q3p_tid();
      const int xj = tid & 3, xez = ((tid >> 2) & 1) | ((tid >> 6) & 2), xk = (tid >> 3) & 3, xey = (tid >> 5) & 3;
      const int xq = 2 * (xj ^ xk) + (xez & 1);
      const int xo0 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * (xq & 7);
      const int xo1 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * ((xq + 1) & 7);
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const double2 lo = *reinterpret_cast<const double2*>(sw + xo0 + 64 * e);
        const double2 hi = *reinterpret_cast<const double2*>(sw + xo1 + 64 * e);
        double a[4] = {lo.x, lo.y, hi.x, hi.y};
        q3e_mass<false>(a);
        q3e_back(a);
        *reinterpret_cast<double2*>(sw + xo0 + 64 * e) = make_double2(a[0], a[1]);
        *reinterpret_cast<double2*>(sw + xo1 + 64 * e) = make_double2(a[2], a[3]);
      }
    }
    __syncthreads();

    // ---------------- P5: (factor / 8) * M_z, back to nodal in z, coalesced store ----------------
    {
      const int tid = q3p_tid();
      const int zq = tid & 15, zex = (tid >> 4) & 3, zey = tid >> 6;
      const int zcol = (zex + 4 * zey) * N3;
      double* __restrict__ yo_g = P.y + (long)(e0 + zex + n0 * zey) * N3 + zq;
      auto tile_out = [&](auto accum) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          double a[4];
#pragma unroll
          for (int k = 0; k < 4; k++) a[k] = sw[zcol + 1024 * e + 16 * k + (((zq ^ (4 * k)) + 2 * (e & 1)) & 15)];
          q3e_mass<true>(a);
          q3e_back(a);
          double* yo_e = yo_g + (long)(n01 * e) * N3;
#pragma unroll
          for (int k = 0; k < 4; k++) yo_e[N2 * k] = decltype(accum)::value ? yo_e[N2 * k] + a[k] : a[k];
        }
      };
      if (P.accum) tile_out(std::true_type{}); else tile_out(std::false_type{});
    }
    if (!has_next) break;
    t = tn;
  }
}
