// Persistent Q3 (N = 4) tile kernel of the fast-diagonalisation block Jacobi on a uniform-degree 3-D brick.
//
//   c_e = damping * (Vx x Vy x Vz) diag(1/(lx_i + ly_j + lz_k)) (Vx x Vy x Vz)^T r_e          (see jacobi.cu, jacobi_uniform.cu)
//
// Replaces IPDGBlockJacobi driven by Operator::apply (matrix-free/localoperators/ipdgblockjacobi.hh:58-178) with an exact local
// solver.  Skeleton of the persistent operator kernel (apply_uniform_q3p.cuh): persistent CTAs, tiles of 4x4x4 elements from a
// global counter, the tile of r double-buffered in shared memory by bulk async copies (cp.async.bulk + mbarrier) issued a whole
// tile ahead, the result written back in place and sent off by bulk stores.  There is no neighbour coupling, so the six 1-D
// sweeps are grouped into THREE stages that each keep their data in registers and rewrite the tile in place (the shared-memory
// pipe was the limiter of the five-pass version, ncu: 63 % busy):
//   A  one thread per (y,z)-plane of an element (fixed x-node i): Vy^T along j, Vz^T along k                 [16 values / thread]
//   B  one thread per x-line, line (j,k) of the 4 elements of a row: Vx^T, scale by damping/(lx_i+ly_j+lz_k), Vx [128-bit accesses]
//   C  planes again: Vz, Vy, back in place, bulk store to c  (V-cycle mode: c and x += c straight from the registers)
// i.e. 5 shared-memory accesses per DoF instead of 9 and 2 block barriers per tile instead of 4.
// Layout: the tile's 16 rows (ey + 4 ez) of four x-contiguous elements (2 KB, contiguous in a DynamicBlockVector) arrive by one bulk
// copy each and sit 258 doubles apart (16 bytes of padding per row): with the lane assignments below every access of stages A-C is
// bank-conflict free (A/C: a half warp = 4 i x 4 rows of stride 2 at one x-position, bank (2 row + i + 4 j) mod 16; B: a quarter
// warp = 4 j x 2 neighbouring rows, 16-byte unit (row + 2 j) mod 8).  A bulk copy costs ~10 issue slots in a per-lane loop, hence
// rows in (16 per tile); the result leaves element by element (64 per tile) because in stage C a warp owns whole elements but
// not whole rows, and a warp-local hand-over to the async proxy needs no block barrier.
// The 1-D factor of a direction depends only on whether the element is the first / last of its grid line at a domain boundary
// (variants 1 / 2; 0 = interior faces on both sides).  Tiles that touch no domain boundary (2/3 of cfg2) run with the interior
// tables as immediate-offset constant-bank operands: the interior factor is mirror symmetric (even / odd eigenvectors), 8 instead
// of 16 distinct entries per direction, so the tables of all three directions fit in the uniform register file.  Tiles on the
// domain boundary take one of seven further instantiations (one per set of touched directions) in which only the sweeps normal
// to a touched boundary read their element's factor from a shared-memory copy into registers, once per stage, with run-time
// variant indices (no divergence inside a tile).  The reciprocals come from a host-built table (27 variant combinations x 64
// doubles, L1 resident): no FP64 divisions in the kernel.
#pragma once
#include <type_traits>

#include "q3p_common.cuh"

namespace hpdg {

struct Q3jParams {
  // [direction][variant][row-major 4x4: node x eigenvector].  Variant 0 (interior faces on both sides) is mirror symmetric: the
  // host orders its eigenvectors even, odd, even, odd under the node reflection (jacobi_uniform.cu), so V[3 - i][k] = (-1)^k V[i][k]
  // and the fast path reads only rows 0 and 1 (8 doubles per direction).  The damping is folded into `inv`.
  double V[3][3][16];
  const double* r;
  double* c;
  double* xacc;          // optional: x += c
  const double* inv;     // [vx][vy][vz][j][k][i] = damping / (lx_i + ly_j + lz_k)
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
// the same with a run-time table held in registers (boundary variants)
template <bool TRANS>
__device__ __forceinline__ void q3j_line_rt(const double (&V)[16], double (&a)[4]) {
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
// GENERAL: V is the element's factor of direction D (loaded from shared memory); otherwise the interior factor from the constant bank
template <int D, bool TRANS, bool GENERAL>
__device__ __forceinline__ void q3j_sweep(const double (&V)[16], double (&a)[4]) {
  if constexpr (GENERAL) q3j_line_rt<TRANS>(V, a);
  else q3j_line<D, TRANS>(a);
}
template <bool GENERAL>
__device__ __forceinline__ void q3j_load_factor(const double* __restrict__ src, double (&V)[16]) {
  if constexpr (GENERAL) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const double2 v = reinterpret_cast<const double2*>(src)[q];
      V[2 * q] = v.x; V[2 * q + 1] = v.y;
    }
  }
}

constexpr int kQ3jRStride = 258;                   // doubles between the rows (4 x-contiguous elements, 256 doubles) of a tile in shared memory
constexpr int kQ3jBuf = 16 * kQ3jRStride;          // one tile buffer
constexpr int kQ3jSmemBytes = (2 * kQ3jBuf + 4 + 9 * 16) * 8;

}  // namespace hpdg

extern "C" __global__ void __launch_bounds__(256, 3)
hpdg_k_jacobi_fd_q3_persist(const __grid_constant__ hpdg::Q3jParams P) {
  using namespace hpdg;
  constexpr int N3 = 64, RS = kQ3jRStride;
  extern __shared__ __align__(128) double q3j_sm[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(q3j_sm + 2 * kQ3jBuf);  // one per buffer
  volatile int* s_next = reinterpret_cast<volatile int*>(q3j_sm + 2 * kQ3jBuf + 2);
  double* __restrict__ vb = q3j_sm + 2 * kQ3jBuf + 4;                  // [direction][variant][16]
  const double* __restrict__ R = P.r;
  const int n0 = P.n[0], n01 = P.n[0] * P.n[1];
  const int ntiles = P.ntiles;

  // the first 2 lanes of every warp fetch one row of 4 x-contiguous elements each (2 KB) into buffer b: a bulk copy is issued lane
  // by lane (~10 instructions per copy), so the 16 copies of a tile are spread over the 8 warps
  auto prefetch = [&](int tid, int e0, int b) {
    if ((tid & 31) < 2) {
      if (tid == 0) q3p_mbar_expect_tx(mbar + b, 32768u);
      const int row = (tid & 1) + 2 * (tid >> 5);  // ey + 4 ez
      q3p_bulk_g2s(q3j_sm + kQ3jBuf * b + RS * row, R + (long)(e0 + n0 * (row & 3) + n01 * (row >> 2)) * N3, 2048u, mbar + b);
    }
  };

  if (threadIdx.x < 144) vb[threadIdx.x] = (&P.V[0][0][0])[threadIdx.x];
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
    // dynamic tile scheduling as in the operator kernel: thread 0 draws the next tile, the others read it after the first barrier
    if (threadIdx.x == 0) *s_next = (int)gridDim.x + atomicAdd(P.sched, 1);
    const int e0 = td.x, fl = td.z & bmask;  // faces of the tile on a domain boundary
    double* __restrict__ sw = q3j_sm + kQ3jBuf * buf;
    bool has_next = false;
    int tn = 0;

    // G: bit d set = the tile touches a domain boundary in direction d, the sweeps of that direction take the elements' factors
    // from shared memory (run-time variant, uniform over the tile's code path: no divergence)
    auto tile = [&](auto gc) {
      constexpr int G = decltype(gc)::value;
      constexpr bool GX = G & 1, GY = (G >> 1) & 1, GZ = (G >> 2) & 1;
      // boundary variant of an element of the tile along one direction: 1 / 2 = first / last of its grid line at a domain boundary
      auto var = [&](int ec, int dir) { return (ec == 0 && ((fl >> (2 * dir)) & 1)) ? 1 : (ec == 3 && ((fl >> (2 * dir + 1)) & 1)) ? 2 : 0; };

      // ---------------- A: (y,z)-planes: Vy^T along j, Vz^T along k; in place ----------------
      {
        const int tid = q3p_tid();
        const int row = 2 * ((tid >> 2) & 3) + ((tid >> 4) & 1) + 8 * ((tid >> 5) & 1);  // ey + 4 ez
        const int base = RS * row + 64 * (tid >> 6) + (tid & 3);
        while (!q3p_mbar_try_wait(mbar + buf, (phase >> buf) & 1)) {}
        phase ^= 1u << buf;
        double a[4][4];  // [k][j]
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int j = 0; j < 4; j++) a[k][j] = sw[base + 4 * j + 16 * k];
        {
          double V[16];
          q3j_load_factor<GY>(vb + (3 + var(row & 3, 1)) * 16, V);
#pragma unroll
          for (int k = 0; k < 4; k++) q3j_sweep<1, true, GY>(V, a[k]);
        }
        {
          double V[16];
          q3j_load_factor<GZ>(vb + (6 + var(row >> 2, 2)) * 16, V);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            double l[4] = {a[0][j], a[1][j], a[2][j], a[3][j]};
            q3j_sweep<2, true, GZ>(V, l);
#pragma unroll
            for (int k = 0; k < 4; k++) a[k][j] = l[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int j = 0; j < 4; j++) sw[base + 4 * j + 16 * k] = a[k][j];
      }
      // the bulk stores of the tile before (other buffer) have had this stage to read their source
      if ((q3p_tid() & 31) < 8) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      __syncthreads();

      // the other buffer is free (every warp is past stage C of the previous tile and its stores have been read): fetch the next
      // tile, a whole tile ahead
      tn = *s_next;
      has_next = tn < ntiles;
      if (has_next) {
        td = __ldg(P.tile_desc + tn);
        const int tid = q3p_tid();
        if ((tid & 31) < 2) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        prefetch(tid, td.x, buf ^ 1);
      }

      // ---------------- B: x-lines, line (j,k) of the 4 elements of a row: Vx^T, scale, Vx; in place, 128-bit accesses ----------------
      {
        const int tid = q3p_tid();
        const int j = tid & 3, k = (tid >> 3) & 3;
        const int row = ((tid >> 2) & 1) + 2 * (tid >> 5);
        double* __restrict__ lp = sw + RS * row + 4 * j + 16 * k;
        const double* __restrict__ ip = P.inv + 16 * j + 4 * k;  // [variant combination][j][k][i]
        double2 s0, s1;
        if constexpr (G == 0) { s0 = __ldg(reinterpret_cast<const double2*>(ip)); s1 = __ldg(reinterpret_cast<const double2*>(ip) + 1); }
        const int vy = GY ? var(row & 3, 1) : 0, vz = GZ ? var(row >> 2, 2) : 0;
        q3p_for<4>([&](auto exc) {
          constexpr int ex = decltype(exc)::value;
          constexpr bool RT = GX && (ex == 0 || ex == 3);  // only the ends of the row can sit on an x-boundary
          const int vx = RT ? var(ex, 0) : 0;
          double V[16];
          q3j_load_factor<RT>(vb + vx * 16, V);
          if constexpr (G != 0) {
            const double2* q = reinterpret_cast<const double2*>(ip + ((vx * 3 + vy) * 3 + vz) * 64);
            s0 = __ldg(q); s1 = __ldg(q + 1);
          }
          const double2 lo = *reinterpret_cast<const double2*>(lp + 64 * ex), hi = *reinterpret_cast<const double2*>(lp + 64 * ex + 2);
          double a[4] = {lo.x, lo.y, hi.x, hi.y};
          q3j_sweep<0, true, RT>(V, a);
          a[0] *= s0.x; a[1] *= s0.y; a[2] *= s1.x; a[3] *= s1.y;
          q3j_sweep<0, false, RT>(V, a);
          *reinterpret_cast<double2*>(lp + 64 * ex) = make_double2(a[0], a[1]);
          *reinterpret_cast<double2*>(lp + 64 * ex + 2) = make_double2(a[2], a[3]);
        });
      }
      __syncthreads();

      // ---------------- C: planes again: Vz along k, Vy along j; out ----------------
      {
        const int tid = q3p_tid();
        const int row = 2 * ((tid >> 2) & 3) + ((tid >> 4) & 1) + 8 * ((tid >> 5) & 1);
        const int base = RS * row + 64 * (tid >> 6) + (tid & 3);
        double a[4][4];  // [k][j]
        {
          double V[16];
          q3j_load_factor<GZ>(vb + (6 + var(row >> 2, 2)) * 16, V);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            double l[4];
#pragma unroll
            for (int k = 0; k < 4; k++) l[k] = sw[base + 4 * j + 16 * k];
            q3j_sweep<2, false, GZ>(V, l);
#pragma unroll
            for (int k = 0; k < 4; k++) a[k][j] = l[k];
          }
        }
        {
          double V[16];
          q3j_load_factor<GY>(vb + (3 + var(row & 3, 1)) * 16, V);
#pragma unroll
          for (int k = 0; k < 4; k++) q3j_sweep<1, false, GY>(V, a[k]);
        }
        {
          // back in place, then one bulk store per element (512 B) issued by the warp that owns the element in this stage;
          // V-cycle: additionally x += c as a bulk reduce-add of the same element (the add runs at the L2)
#pragma unroll
          for (int k = 0; k < 4; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) sw[base + 4 * j + 16 * k] = a[k][j];
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          __syncwarp();
          if ((tid & 31) < 8) {  // this warp's 8 elements: x-position tid >> 6, rows 8 ((tid >> 5) & 1) + 0..7
            const int rs = (tid & 7) + 8 * ((tid >> 5) & 1), exs = tid >> 6;
            const long go = (long)(e0 + exs + n0 * (rs & 3) + n01 * (rs >> 2)) * N3;
            q3p_bulk_s2g(P.c + go, sw + RS * rs + 64 * exs, 512u);
            if (P.xacc) q3p_bulk_s2g_add(P.xacc + go, sw + RS * rs + 64 * exs, 512u);
          }
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
  if ((threadIdx.x & 31) < 8) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
  if (threadIdx.x == 0 && atomicAdd(P.sched + 1, 1) == (int)gridDim.x - 1) { P.sched[0] = 0; P.sched[1] = 0; __threadfence(); }
}
