// Persistent Q4 (N = 5) tile kernel of the fast-diagonalisation block Jacobi on a uniform-degree 3-D brick: the three-stage
// structure of jacobi_uniform_q3p.cuh carried to the degree of the V-cycle / weak-scaling configurations (cfg4, cfg5).  Default for Q4
// bricks with extents multiple of (4, 4, 2) (option "variant" = 40 switches back to k_jacobi_fd_uniform<5,..>); measured 101.6 us per
// sweep at 64^3 (78.9 % of the measured HBM peak, one-tile-per-CTA kernel: 175.5 us) and 689.5 us at 128^3 (93.1 %).
//
//   c_e = damping * (Vx x Vy x Vz) diag(1/(lx_i + ly_j + lz_k)) (Vx x Vy x Vz)^T r_e     (ipdgblockjacobi.hh:58-178, exact local solve)
//
// Tile = 4 x 4 x 2 elements (32 000 B), 160 threads = 5 warps, up to 136 registers per thread at 3 CTAs/SM.
//   warp = plane index (x-node i in stages A and C, z-node k in stage B), lane = element (ex = lane & 3, row = ey + 4 ez = lane >> 2)
//   A  the (y,z)-plane i of the lane's element (25 values in registers): Vy^T along j, Vz^T along k; in place
//   B  the x-lines (j = 0..4, k) of the lane's element: Vx^T, scale by damping/(lx_i+ly_j+lz_k), Vx; in place
//   C  planes again: Vz, Vy; back in place; block barrier; one bulk store per row (V-cycle mode: cooperative coalesced c and x += c)
// Layout: the tile's 8 rows of four x-contiguous elements are contiguous in a DynamicBlockVector (4 000 B: a multiple of 16, which a
// single 1 000-byte element is not) and arrive by one bulk copy each, unpadded: the row stride is 500 doubles = 4 mod 16 and the
// element stride 125 = 13 mod 16, so the 16 lanes of a half warp (4 ex x 4 rows) hit banks 13 ex + 4 row = all 16 distinct at any
// fixed node, in every stage.
// Tables: the interior factor is mirror symmetric (eigenvectors ordered even, odd, even, odd, even by the host): rows 0..2 are read
// (15 doubles per direction, the centre node of an odd eigenvector is zero and skipped) as immediate-offset constant-bank operands.
// Tiles on a domain boundary take one of 7 further instantiations whose sweeps normal to a touched boundary read the element's
// factor from a shared-memory copy into registers once per stage.
#pragma once
#include <type_traits>

#include "q3p_common.cuh"

namespace hpdg {

struct Q4jParams {
  double V[3][3][25];    // [direction][variant][node * 5 + eigenvector]; variant 0 in the mirror-canonical order
  const double* r;
  double* c;
  double* xacc;          // optional: x += c
  const double* inv;     // [vx][vy][vz][k][j][i] = damping / (lx_i + ly_j + lz_k)
  const int4* tile_desc; // .x first element, .z brick-face bits of the 4x4x2 tile
  int* sched;
  int n[3];
  int bnd[6];
  int ntiles;
};

template <int OFF>
__device__ __forceinline__ double q4j_c() {
  double v;
  asm volatile("ld.param.f64 %0, [hpdg_k_jacobi_fd_q4_persist_param_0+%1];\n" : "=d"(v) : "n"(OFF));
  return v;
}

// a <- V^T a (TRANS) or V a with V = V[D][0] read through its mirror symmetry: V[4 - i][k] = (-1)^k V[i][k], V[2][odd k] = 0
template <int D, bool TRANS>
__device__ __forceinline__ void q4j_line(double (&a)[5]) {
  double o[5];
  q3p_for<5>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    double s = 0;
    q3p_for<5>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      constexpr int node = TRANS ? m : i, mode = TRANS ? i : m;
      constexpr int off = (int)offsetof(Q4jParams, V) + 8 * (D * 3 * 25 + (node < 3 ? node : 4 - node) * 5 + mode);
      if constexpr (node == 2 && (mode & 1)) { /* zero entry */ }
      else if constexpr (node >= 3 && (mode & 1)) s = fma(-q4j_c<off>(), a[m], s);
      else s = fma(q4j_c<off>(), a[m], s);
    });
    o[i] = s;
  });
#pragma unroll
  for (int i = 0; i < 5; i++) a[i] = o[i];
}
template <bool TRANS>
__device__ __forceinline__ void q4j_line_rt(const double (&V)[25], double (&a)[5]) {
  double o[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    double s = 0;
#pragma unroll
    for (int m = 0; m < 5; m++) s = fma(V[TRANS ? m * 5 + i : i * 5 + m], a[m], s);
    o[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 5; i++) a[i] = o[i];
}
template <int D, bool TRANS, bool GENERAL>
__device__ __forceinline__ void q4j_sweep(const double (&V)[25], double (&a)[5]) {
  if constexpr (GENERAL) q4j_line_rt<TRANS>(V, a);
  else q4j_line<D, TRANS>(a);
}
template <bool GENERAL>
__device__ __forceinline__ void q4j_load_factor(const double* __restrict__ src, double (&V)[25]) {
  if constexpr (GENERAL) {
#pragma unroll
    for (int q = 0; q < 25; q++) V[q] = src[q];
  }
}

constexpr int kQ4jBuf = 32 * 125;   // one tile buffer (doubles)
constexpr int kQ4jSmemBytes = (2 * kQ4jBuf + 4 + 9 * 25 + 1) * 8;

}  // namespace hpdg

extern "C" __global__ void __launch_bounds__(160, 3)
hpdg_k_jacobi_fd_q4_persist(const __grid_constant__ hpdg::Q4jParams P) {
  using namespace hpdg;
  constexpr int N3 = 125, RS = 500;
  extern __shared__ __align__(128) double q4j_sm[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(q4j_sm + 2 * kQ4jBuf);  // one per buffer
  volatile int* s_next = reinterpret_cast<volatile int*>(q4j_sm + 2 * kQ4jBuf + 2);
  double* __restrict__ vb = q4j_sm + 2 * kQ4jBuf + 4;                  // [direction][variant][25]
  const double* __restrict__ R = P.r;
  const int n0 = P.n[0], n01 = P.n[0] * P.n[1];
  const int ntiles = P.ntiles;

  // lane 0 of warp w moves rows w and w + 5 (rows = ey + 4 ez, 8 per tile)
  auto row_src = [&](int e0, int row) { return (long)(e0 + n0 * (row & 3) + n01 * (row >> 2)) * N3; };
  auto prefetch = [&](int tid, int e0, int b) {
    if ((tid & 31) == 0) {
      if (tid == 0) q3p_mbar_expect_tx(mbar + b, 32000u);
      for (int row = tid >> 5; row < 8; row += 5) q3p_bulk_g2s(q4j_sm + kQ4jBuf * b + RS * row, R + row_src(e0, row), 4000u, mbar + b);
    }
  };

  for (int q = threadIdx.x; q < 225; q += 160) vb[q] = (&P.V[0][0][0])[q];
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
  const int bmask = (P.bnd[0] ? 1 : 0) | (P.bnd[1] ? 2 : 0) | (P.bnd[2] ? 4 : 0) | (P.bnd[3] ? 8 : 0) | (P.bnd[4] ? 16 : 0) | (P.bnd[5] ? 32 : 0);

  for (;;) {
    if (threadIdx.x == 0) *s_next = (int)gridDim.x + atomicAdd(P.sched, 1);
    const int e0 = td.x, fl = td.z & bmask;  // faces of the tile on a domain boundary
    double* __restrict__ sw = q4j_sm + kQ4jBuf * buf;
    bool has_next = false;
    int tn = 0;

    auto tile = [&](auto gc) {
      constexpr int G = decltype(gc)::value;
      constexpr bool GX = G & 1, GY = (G >> 1) & 1, GZ = (G >> 2) & 1;
      // boundary variant of an element along one direction (tile extents 4, 4, 2): 1 / 2 = first / last of its grid line at a boundary
      auto var = [&](int ec, int dir) {
        const int last = dir == 2 ? 1 : 3;
        return (ec == 0 && ((fl >> (2 * dir)) & 1)) ? 1 : (ec == last && ((fl >> (2 * dir + 1)) & 1)) ? 2 : 0;
      };
      const int tid = q3p_tid();
      const int w = tid >> 5, lane = tid & 31;
      const int ex = lane & 3, row = lane >> 2;  // ey = row & 3, ez = row >> 2
      const int ebase = RS * row + N3 * ex;

      // ---------------- A: (y,z)-plane i = w of the lane's element: Vy^T along j, Vz^T along k; in place ----------------
      {
        const int base = ebase + w;
        while (!q3p_mbar_try_wait(mbar + buf, (phase >> buf) & 1)) {}
        phase ^= 1u << buf;
        double a[5][5];  // [k][j]
#pragma unroll
        for (int k = 0; k < 5; k++)
#pragma unroll
          for (int j = 0; j < 5; j++) a[k][j] = sw[base + 5 * j + 25 * k];
        {
          double V[25];
          q4j_load_factor<GY>(vb + (3 + var(row & 3, 1)) * 25, V);
#pragma unroll
          for (int k = 0; k < 5; k++) q4j_sweep<1, true, GY>(V, a[k]);
        }
        {
          double V[25];
          q4j_load_factor<GZ>(vb + (6 + var(row >> 2, 2)) * 25, V);
#pragma unroll
          for (int j = 0; j < 5; j++) {
            double l[5] = {a[0][j], a[1][j], a[2][j], a[3][j], a[4][j]};
            q4j_sweep<2, true, GZ>(V, l);
#pragma unroll
            for (int k = 0; k < 5; k++) a[k][j] = l[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 5; k++)
#pragma unroll
          for (int j = 0; j < 5; j++) sw[base + 5 * j + 25 * k] = a[k][j];
      }
      // the bulk stores of the tile before (other buffer) have had this stage to read their source
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      __syncthreads();

      // the other buffer is free: fetch the next tile, a whole tile ahead
      tn = *s_next;
      has_next = tn < ntiles;
      if (has_next) {
        td = __ldg(P.tile_desc + tn);
        if (lane == 0) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        prefetch(tid, td.x, buf ^ 1);
      }

      // ---------------- B: the x-lines (j, k = w) of the lane's element: Vx^T, scale, Vx; in place ----------------
      {
        const int vx = GX ? var(ex, 0) : 0, vy = GY ? var(row & 3, 1) : 0, vz = GZ ? var(row >> 2, 2) : 0;
        const double* __restrict__ ip = P.inv + ((vx * 3 + vy) * 3 + vz) * N3 + 25 * w;  // [j][i] of this k
        double V[25];
        q4j_load_factor<GX>(vb + vx * 25, V);
#pragma unroll
        for (int j = 0; j < 5; j++) {
          double* __restrict__ lp = sw + ebase + 25 * w + 5 * j;
          double a[5];
#pragma unroll
          for (int i = 0; i < 5; i++) a[i] = lp[i];
          q4j_sweep<0, true, GX>(V, a);
#pragma unroll
          for (int i = 0; i < 5; i++) a[i] *= __ldg(ip + 5 * j + i);
          q4j_sweep<0, false, GX>(V, a);
#pragma unroll
          for (int i = 0; i < 5; i++) lp[i] = a[i];
        }
      }
      __syncthreads();

      // ---------------- C: planes again: Vz along k, Vy along j; back in place ----------------
      {
        const int base = ebase + w;
        double a[5][5];  // [k][j]
        {
          double V[25];
          q4j_load_factor<GZ>(vb + (6 + var(row >> 2, 2)) * 25, V);
#pragma unroll
          for (int j = 0; j < 5; j++) {
            double l[5];
#pragma unroll
            for (int k = 0; k < 5; k++) l[k] = sw[base + 5 * j + 25 * k];
            q4j_sweep<2, false, GZ>(V, l);
#pragma unroll
            for (int k = 0; k < 5; k++) a[k][j] = l[k];
          }
        }
        {
          double V[25];
          q4j_load_factor<GY>(vb + (3 + var(row & 3, 1)) * 25, V);
#pragma unroll
          for (int k = 0; k < 5; k++) q4j_sweep<1, false, GY>(V, a[k]);
        }
#pragma unroll
        for (int k = 0; k < 5; k++)
#pragma unroll
          for (int j = 0; j < 5; j++) sw[base + 5 * j + 25 * k] = a[k][j];
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncthreads();  // a row holds the planes of all five warps
      if (lane == 0) {
        // one bulk store per row; V-cycle: additionally x += c as a bulk reduce-add of the same row (the add runs at the L2:
        // no thread waits on a load of x)
        for (int row2 = w; row2 < 8; row2 += 5) {
          q3p_bulk_s2g(P.c + row_src(e0, row2), sw + RS * row2, 4000u);
          if (P.xacc) q3p_bulk_s2g_add(P.xacc + row_src(e0, row2), sw + RS * row2, 4000u);
        }
      }
    };
    switch ((fl & 3 ? 1 : 0) | (fl & 12 ? 2 : 0) | (fl & 48 ? 4 : 0)) {
      case 0: tile(std::integral_constant<int, 0>{}); break;
      case 1: tile(std::integral_constant<int, 1>{}); break;
      case 2: tile(std::integral_constant<int, 2>{}); break;
      case 3: tile(std::integral_constant<int, 3>{}); break;
      case 4: tile(std::integral_constant<int, 4>{}); break;
      case 5: tile(std::integral_constant<int, 5>{}); break;
      case 6: tile(std::integral_constant<int, 6>{}); break;
      default: tile(std::integral_constant<int, 7>{}); break;
    }
    if (!has_next) break;
    t = tn; buf ^= 1;
  }
  // shared memory must outlive the bulk stores that read it
  if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
  if (threadIdx.x == 0 && atomicAdd(P.sched + 1, 1) == (int)gridDim.x - 1) { P.sched[0] = 0; P.sched[1] = 0; __threadfence(); }
}
