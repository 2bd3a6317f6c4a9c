// Persistent Q3 (N = 4) tile kernel of the fast-diagonalisation block Jacobi on a uniform-degree 3-D brick.
//
//   c_e = damping * (Vx x Vy x Vz) diag(1/(lx_i + ly_j + lz_k)) (Vx x Vy x Vz)^T r_e          (see jacobi.cu, jacobi_uniform.cu)
//
// Replaces IPDGBlockJacobi driven by Operator::apply (matrix-free/localoperators/ipdgblockjacobi.hh:58-178) with an exact local
// solver.  Same skeleton as the persistent operator kernel (apply_uniform_q3p.cuh): persistent CTAs, tiles of 4x4x4 elements
// from a global counter, unpadded swizzled shared memory, five pencil passes
//   P1 z-pencils: Vz^T      P2 x-pencils: Vx^T      P3 y-pencils: Vy^T, scale by 1/(sum of eigenvalues), Vy
//   P4 x-pencils: Vx        P5 z-pencils: Vz -> global c (+ optional x += c); the damping is folded into the reciprocals
// but there is no neighbour coupling, so one array per tile is enough and the tile buffer is DOUBLE-buffered: the bulk copies of
// the next tile are issued at the start of the current one and have a whole tile time to land.
// The 1-D factor of a direction depends only on whether the element is the first / last of its grid line at a domain boundary
// (variants 1 / 2; 0 = interior faces on both sides).  Inside a pencil only element 0 / 3 of a boundary tile can differ, so
// the variant is a warp-uniform branch on two of the four unrolled elements.  The reciprocals 1/(lx + ly + lz) come from a
// small host-built table (27 variant combinations x 64 doubles, L1/L2 resident): no FP64 divisions in the kernel.
// The interior factor is mirror symmetric (even / odd eigenvectors): 8 instead of 16 distinct table entries per direction, so
// the tables of all three directions fit in the uniform register file.
#pragma once
#include "q3p_common.cuh"

namespace hpdg {

struct Q3jParams {
  // [direction][variant][row-major 4x4: node x eigenvector].  Variant 0 (interior faces on both sides) is mirror symmetric: the
  // host orders its eigenvectors even, odd, even, odd under the node reflection (jacobi_uniform.cu), so V[3 - i][k] = (-1)^k V[i][k]
  // and the kernel reads only rows 0 and 1 (8 doubles per direction).  The damping is folded into `inv`.
  double V[3][3][16];
  const double* r;
  double* c;
  double* xacc;          // optional: x += c
  const double* inv;     // [vx][vz][vy][i][k][j] = damping / (lx_i + ly_j + lz_k)
  const int4* tile_desc;
  int* sched;
  int n[3];
  int bnd[6];            // brick face f is a domain boundary (as opposed to a rank boundary)
  int ntiles;
};

template <int OFF>
__device__ __forceinline__ double q3j_c() {
  double v;
  asm volatile("ld.param.f64 %0, [hpdg_k_jacobi_fd_q3_persist_param_0+%1];\n" : "=d"(v) : "n"(OFF));
  return v;
}

// a <- V^T a (TRANS) or V a with V = V[D][0], read through its mirror symmetry
template <int D, bool TRANS>
__device__ __forceinline__ void q3j_line(double (&a)[4]) {
  double o[4];
  q3p_for<4>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    double s = 0;
    q3p_for<4>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      constexpr int node = TRANS ? m : i, mode = TRANS ? i : m;
      constexpr int off = (int)offsetof(Q3jParams, V) + 8 * (D * 3 * 16 + (node < 2 ? node : 3 - node) * 4 + mode);
      if constexpr (node >= 2 && (mode & 1)) s = fma(-q3j_c<off>(), a[m], s);
      else s = fma(q3j_c<off>(), a[m], s);
    });
    o[i] = s;
  });
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = o[i];
}
// Boundary variants (the first / last element of a grid line at a domain boundary; only in tiles on the brick surface): the table
// is read from a small shared-memory copy with a run-time index.  Selecting between three inlined constant-bank versions makes
// ptxas preload the tables of all variants and spill the 63 uniform registers (R2UR / local memory) in every pass; an
// out-of-line function costs uniform-register saves around every call site.
template <bool TRANS>
__device__ __forceinline__ void q3j_line_rt(const double* __restrict__ V, double (&a)[4]) {
  double o[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) s = fma(V[TRANS ? m * 4 + i : i * 4 + m], a[m], s);
    o[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = o[i];
}
// vb: shared-memory copy of the boundary variants, [D][var - 1][16].
// E = position of the element in its pencil (compile time); var: 0, or 1 / 2 if the element is the first / last of its grid
// line at a domain boundary (only possible for E = 0 / 3)
template <int D, bool TRANS, int E>
__device__ __forceinline__ void q3j_sweep(const double* __restrict__ vb, int var, double (&a)[4]) {
  if ((E == 0 || E == 3) && var != 0) q3j_line_rt<TRANS>(vb + ((D * 2 + var - 1) << 4), a);
  else q3j_line<D, TRANS>(a);
}

constexpr int kQ3jSmemBytes = 2 * 4096 * 8 + 32 + 6 * 16 * 8;

}  // namespace hpdg

extern "C" __global__ void __launch_bounds__(256, 3)
hpdg_k_jacobi_fd_q3_persist(const __grid_constant__ hpdg::Q3jParams P) {
  using namespace hpdg;
  constexpr int N3 = 64;
  extern __shared__ __align__(128) double q3j_sm[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(q3j_sm + 8192);  // one per buffer
  volatile int* s_next = reinterpret_cast<volatile int*>(q3j_sm + 8194);
  double* __restrict__ vb = q3j_sm + 8196;
  const double* __restrict__ R = P.r;
  const int n0 = P.n[0], n01 = P.n[0] * P.n[1];
  const int ntiles = P.ntiles;

  // threads 0..15 fetch the tile's 16 rows of four x-contiguous elements (2 KB each) into buffer b
  auto prefetch = [&](int tid, int e0, int b) {
    if (tid < 16) {
      if (tid == 0) q3p_mbar_expect_tx(mbar + b, 32768u);
      const int ey = tid & 3, ez = tid >> 2;
      q3p_bulk_g2s(q3j_sm + 4096 * b + (4 * ey + 16 * ez) * N3, R + (long)(e0 + n0 * ey + n01 * ez) * N3, 2048u, mbar + b);
    }
  };

  if (threadIdx.x < 96) {
    const int d = threadIdx.x >> 5, v = (threadIdx.x >> 4) & 1, i = threadIdx.x & 15;
    vb[threadIdx.x] = P.V[d][v + 1][i];
  }
  if (threadIdx.x == 0) {
    q3p_mbar_init(mbar, 1);
    q3p_mbar_init(mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  int t = blockIdx.x;
  if (t >= ntiles) return;
  int4 td = __ldg(P.tile_desc + t);
  prefetch(threadIdx.x, td.x, 0);
  uint32_t phase = 0;  // bit b: phase of buffer b
  int buf = 0;

  for (;;) {
    // dynamic tile scheduling as in the operator kernel: thread 0 draws the next tile, the others read it after the first barrier
    if (threadIdx.x == 0) *s_next = (int)gridDim.x + atomicAdd(P.sched, 1);
    const int e0 = td.x, fl = td.z;
    double* __restrict__ sw = q3j_sm + 4096 * buf;
    const bool lox = (fl & 1) && P.bnd[0], hix = (fl & 2) && P.bnd[1];
    const bool loy = (fl & 4) && P.bnd[2], hiy = (fl & 8) && P.bnd[3];
    const bool loz = (fl & 16) && P.bnd[4], hiz = (fl & 32) && P.bnd[5];
    auto variant = [](int e, bool lo, bool hi) { return (e == 0 && lo) ? 1 : (e == 3 && hi) ? 2 : 0; };

    // ---------------- P1: z-pencils, Vz^T; raw tile rewritten in place (swizzled) ----------------
    {
      const int tid = q3p_tid();
      const int zq = tid & 15, zex = (tid >> 4) & 3, zey = tid >> 6;
      const int zcol = (zex + 4 * zey) * N3;
      while (!q3p_mbar_try_wait(mbar + buf, (phase >> buf) & 1)) {}
      phase ^= 1u << buf;
      double v[4][4];
#pragma unroll
      for (int e = 0; e < 4; e++)
#pragma unroll
        for (int k = 0; k < 4; k++) v[e][k] = sw[zcol + 1024 * e + 16 * k + zq];
      __syncwarp();
      q3p_for<4>([&](auto ec) {
        constexpr int e = decltype(ec)::value;
        q3j_sweep<2, true, e>(vb, variant(e, loz, hiz), v[e]);
#pragma unroll
        for (int k = 0; k < 4; k++) sw[zcol + 1024 * e + 16 * k + (((zq ^ (4 * k)) + 2 * (e & 1)) & 15)] = v[e][k];
      });
    }
    __syncthreads();

    // the other buffer is free (every warp is past pass 5 of the previous tile): fetch the next tile, a whole tile ahead
    const int tn = *s_next;
    const bool has_next = tn < ntiles;
    if (has_next) {
      td = __ldg(P.tile_desc + tn);
      const int tid = q3p_tid();
      if (tid < 16) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      prefetch(tid, td.x, buf ^ 1);
    }

    // ---------------- P2: x-pencils, Vx^T (128-bit shared-memory accesses) ----------------
    {
      const int tid = q3p_tid();
      const int xj = tid & 3, xez = ((tid >> 2) & 1) | ((tid >> 6) & 2), xk = (tid >> 3) & 3, xey = (tid >> 5) & 3;
      const int xq = 2 * (xj ^ xk) + (xez & 1);
      const int xo0 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * (xq & 7);
      const int xo1 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * ((xq + 1) & 7);
      q3p_for<4>([&](auto ec) {
        constexpr int e = decltype(ec)::value;
        const double2 lo = *reinterpret_cast<const double2*>(sw + xo0 + 64 * e);
        const double2 hi = *reinterpret_cast<const double2*>(sw + xo1 + 64 * e);
        double a[4] = {lo.x, lo.y, hi.x, hi.y};
        q3j_sweep<0, true, e>(vb, variant(e, lox, hix), a);
        *reinterpret_cast<double2*>(sw + xo0 + 64 * e) = make_double2(a[0], a[1]);
        *reinterpret_cast<double2*>(sw + xo1 + 64 * e) = make_double2(a[2], a[3]);
      });
    }

    // ---------------- P3: y-pencils: Vy^T, scale by the reciprocal eigenvalue sums, Vy ----------------
    {
      const int tid = q3p_tid();
      const int yi = tid & 3, yk = (tid >> 2) & 3, yex = (tid >> 4) & 3, yez = tid >> 6;
      const int ybase = (yex + 16 * yez) * N3 + 16 * yk;
      const int yr = yi + 2 * (yez & 1);
      auto yo = [&](int j) { return ybase + (((4 * j) ^ (4 * yk)) + yr & 15); };
      const int vx = (yex == 0 && lox) ? 1 : (yex == 3 && hix) ? 2 : 0;
      const int vz = (yez == 0 && loz) ? 1 : (yez == 3 && hiz) ? 2 : 0;
      const double* __restrict__ ip = P.inv + ((vx * 3 + vz) * 3) * 64 + (yi * 4 + yk) * 4;  // + vy * 64 + j
      const double2 i0a = __ldg(reinterpret_cast<const double2*>(ip)), i0b = __ldg(reinterpret_cast<const double2*>(ip) + 1);
      q3p_bar_half(tid >> 7);
      q3p_for<4>([&](auto ec) {
        constexpr int e = decltype(ec)::value;
        double a[4];
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = sw[yo(j) + 256 * e];
        q3j_sweep<1, true, e>(vb, variant(e, loy, hiy), a);
        if ((e == 0 && loy) || (e == 3 && hiy)) {
          const double* q = ip + (e == 0 ? 64 : 128);
          const double2 ia = __ldg(reinterpret_cast<const double2*>(q)), ib = __ldg(reinterpret_cast<const double2*>(q) + 1);
          a[0] *= ia.x; a[1] *= ia.y; a[2] *= ib.x; a[3] *= ib.y;
        } else { a[0] *= i0a.x; a[1] *= i0a.y; a[2] *= i0b.x; a[3] *= i0b.y; }
        q3j_sweep<1, false, e>(vb, variant(e, loy, hiy), a);
#pragma unroll
        for (int j = 0; j < 4; j++) sw[yo(j) + 256 * e] = a[j];
      });
      q3p_bar_half(tid >> 7);
    }

    // ---------------- P4: x-pencils, Vx ----------------
    {
      const int tid = q3p_tid();
      const int xj = tid & 3, xez = ((tid >> 2) & 1) | ((tid >> 6) & 2), xk = (tid >> 3) & 3, xey = (tid >> 5) & 3;
      const int xq = 2 * (xj ^ xk) + (xez & 1);
      const int xo0 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * (xq & 7);
      const int xo1 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * ((xq + 1) & 7);
      q3p_for<4>([&](auto ec) {
        constexpr int e = decltype(ec)::value;
        const double2 lo = *reinterpret_cast<const double2*>(sw + xo0 + 64 * e);
        const double2 hi = *reinterpret_cast<const double2*>(sw + xo1 + 64 * e);
        double a[4] = {lo.x, lo.y, hi.x, hi.y};
        q3j_sweep<0, false, e>(vb, variant(e, lox, hix), a);
        *reinterpret_cast<double2*>(sw + xo0 + 64 * e) = make_double2(a[0], a[1]);
        *reinterpret_cast<double2*>(sw + xo1 + 64 * e) = make_double2(a[2], a[3]);
      });
    }
    __syncthreads();

    // ---------------- P5: damping * Vz, coalesced store (and x += c) ----------------
    {
      const int tid = q3p_tid();
      const int zq = tid & 15, zex = (tid >> 4) & 3, zey = tid >> 6;
      const int zcol = (zex + 4 * zey) * N3;
      const long gofs = (long)(e0 + zex + n0 * zey) * N3 + zq;
      double* __restrict__ co = P.c + gofs;
      auto tile_out = [&](auto acc) {
        q3p_for<4>([&](auto ec) {
          constexpr int e = decltype(ec)::value;
          double a[4];
#pragma unroll
          for (int k = 0; k < 4; k++) a[k] = sw[zcol + 1024 * e + 16 * k + (((zq ^ (4 * k)) + 2 * (e & 1)) & 15)];
          q3j_sweep<2, false, e>(vb, variant(e, loz, hiz), a);
          const long eo = (long)(n01 * e) * N3;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            co[eo + 16 * k] = a[k];
            if (decltype(acc)::value) P.xacc[gofs + eo + 16 * k] += a[k];
          }
        });
      };
      if (P.xacc) tile_out(std::true_type{}); else tile_out(std::false_type{});
    }
    if (!has_next) break;
    t = tn; buf ^= 1;
  }
  if (threadIdx.x == 0 && atomicAdd(P.sched + 1, 1) == (int)gridDim.x - 1) { P.sched[0] = 0; P.sched[1] = 0; __threadfence(); }
}
