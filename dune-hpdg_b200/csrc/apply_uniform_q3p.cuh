// Persistent, prefetching Q3 (N = 4) tile kernel of the uniform-degree 3-D SIPG operator apply: the headline kernel.
//
// Same operator, formulation and five pencil passes as k_apply_uniform (apply_uniform.cu; reference: Operator::apply over
// IPDGOperator, matrix-free/operator.hh:41-56, matrix-free/localoperators/ipdgoperator.hh:80-390).  What changes is how the
// DoF blocks move:
//  * CTAs are persistent (one per SM slot; tiles are handed out x-fastest by a global counter, so concurrent CTAs work on
//    neighbouring tiles and their halo reads hit L2, and the last wave stays short).  The 4x4x4-element tile of u arrives
//    in shared memory by 16 bulk async copies (cp.async.bulk, 2 KB each: the four x-contiguous elements of a row are
//    contiguous in a DynamicBlockVector) that complete on an mbarrier.  The copies for the next tile are issued, per CTA
//    half, right after pass 3 of the current one (the last reader of the u buffer), so they are in flight during passes
//    4 and 5 and no thread waits on a global load of u.
//  * Shared memory is unpadded (2 x 32 KB per CTA -> 3 CTAs/SM).  Pass 1 rewrites u in place into a swizzled layout:
//    within the 128-byte row of an (element, z-plane k) the 32-byte x-line j sits at line slot j^k, rotated by 16 bytes
//    for odd element layers.  With that, all three pencil directions are bank-conflict free: z-pencils (64-bit, a half
//    warp = the 16 nodes of a plane), y-pencils (64-bit, half warp = 4 i x 4 k), x-pencils (128-bit, quarter warp = 4 j x
//    2 element layers).  The in-place rewrite only permutes data among the lanes of one half warp (they own the column of
//    elements they read), so a __syncwarp between the reads and the writes is all it needs.
//  * Passes 2-4 couple the element layers {0,1} and {2,3} of a tile separately: 128-thread named barriers.  Pass 5 and the
//    next tile's pass 1 use the same thread -> column mapping, so no block barrier separates two tiles.
//  * Multi-GPU (NVLink peer-memory halo): the CTAs pack and publish the brick's boundary traces before their first tile.
// Requires brick extents that are multiples of 4 (otherwise the masked k_apply_uniform is used).  Measurements, what bounds
// the kernel and what was tried: DESIGN.md section 6.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>

#include "q3p_common.cuh"
#include "uniform_common.cuh"

namespace hpdg {

// ---- table access ------------------------------------------------------------------------------------------------------
// The 1-D tables live in the kernel's parameter block (constant bank).  Read through the struct, NVVM hoists all ~180
// table loads out of the persistent tile loop; they then overflow the uniform register file and ptxas spills them to local
// memory and moves them back with R2UR inside the passes.  Reading them with a (volatile, immediate-offset) ld.param keeps
// the loads where they are used; ptxas still merges and schedules them like ordinary constant-bank operands.
template <int OFF>
__device__ __forceinline__ double q3p_c() {
  double v;
  asm volatile("ld.param.f64 %0, [hpdg_k_apply_q3_persist_param_0+%1];\n" : "=d"(v) : "n"(OFF));
  return v;
}
#define Q3P_C(field, idx) q3p_c<(int)offsetof(UniParams<4>, field) + 8 * (idx)>()

// The GL nodes are symmetric about the element centre, so with R the node reflection: M, Dp commute with R, M is symmetric,
// g_1 = -R g_0, A1 = -R A0, B1 = R B0.  Only one representative of each group of equal entries is read (27 distinct doubles
// for a T-sweep + mass pass instead of 57), which lets a pass's tables fit in the uniform register file.
__host__ __device__ constexpr int q3p_dp_idx(int dir, int i, int m) { return dir * 16 + (i < 2 ? i * 4 + m : (3 - i) * 4 + (3 - m)); }
__host__ __device__ constexpr int q3p_m_idx(int i, int m) {
  int best = i * 4 + m;
  const int c1 = m * 4 + i, c2 = (3 - i) * 4 + (3 - m), c3 = (3 - m) * 4 + (3 - i);
  if (c1 < best) best = c1;
  if (c2 < best) best = c2;
  if (c3 < best) best = c3;
  return best;
}


// acc_e = accin_e + (Tt_dir v)_e for the 4 elements of a full pencil (N = 4); same arithmetic as pencil_apply
// (uniform_common.cuh).  load(e, v) fetches the DoF line of element e, accin(e, a) fills the accumulator's initial values
// of element e, out(e, v, a) consumes the element's line and results.  The elements are processed in the order 1, 2, 3, 0
// so that the traces of the elements outside the pencil (global loads issued just before the call) are consumed last.
template <int DIR, class Load, class AccIn, class Out>
__device__ __forceinline__ void q3p_pencil(double pd, double pv, int pmode, double nd, double nv, int nmode, Load load,
                                           AccIn accin, Out out) {
  constexpr int N = 4, T = 4;
  double v[T][N], d0[T], d1[T];
#pragma unroll
  for (int e = 0; e < T; e++) {
    load(e, v[e]);
    double a = 0, b = 0;
    q3p_for<N>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      a = fma(Q3P_C(g, m), v[e][m], a); b = fma(-Q3P_C(g, N - 1 - m), v[e][m], b);
    });
    d0[e] = a; d1[e] = b;
  }
#pragma unroll
  for (int ee = 0; ee < T; ee++) {
    const int e = (ee + 1) % T;
    double qd, qv, rd, rv;
    if (e == 0) {
      if (pmode == 1) { pd = fma(-Q3P_C(cohk, DIR), v[0][0], d0[0]); pv = -v[0][0]; }
      else if (pmode == 2) { pd = -d0[0]; pv = v[0][0]; }
      qd = pd; qv = pv;
    } else { qd = d1[e > 0 ? e - 1 : 0]; qv = v[e > 0 ? e - 1 : 0][N - 1]; }
    if (e == T - 1) {
      if (nmode == 1) { nd = fma(Q3P_C(cohk, DIR), v[T - 1][N - 1], d1[T - 1]); nv = -v[T - 1][N - 1]; }
      else if (nmode == 2) { nd = -d1[T - 1]; nv = v[T - 1][N - 1]; }
      rd = nd; rv = nv;
    } else { rd = d0[e < T - 1 ? e + 1 : e]; rv = v[e < T - 1 ? e + 1 : e][0]; }
    double a[N];
    accin(e, a);
    q3p_for<N>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      double s = a[i];
      q3p_for<N>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        s = fma(Q3P_C(Dp, q3p_dp_idx(DIR, i, m)), v[e][m], s);
      });
      s = fma(Q3P_C(A0, DIR * N + i), qd, s); s = fma(Q3P_C(B0, DIR * N + i), qv, s);
      s = fma(-Q3P_C(A0, DIR * N + N - 1 - i), rd, s); s = fma(Q3P_C(B0, DIR * N + N - 1 - i), rv, s);
      a[i] = s;
    });
    out(e, v[e], a);
  }
}

template <bool SCALED>
__device__ __forceinline__ void q3p_mass(double (&a)[4]) {
  double o[4];
  q3p_for<4>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    double s = 0;
    q3p_for<4>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      s = fma(SCALED ? Q3P_C(Mf, q3p_m_idx(i, m)) : Q3P_C(M, q3p_m_idx(i, m)), a[m], s);
    });
    o[i] = s;
  });
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = o[i];
}

constexpr int kQ3pSmemBytes = 2 * 4096 * 8 + 16;

// Loads of the (der, val) traces of the elements before / after a pencil in direction DIR; issued before the barrier that opens
// the pass and reduced after it (uniform_common.cuh: halo_issue / halo_reduce).  fl: which brick faces the tile touches;
// prev / next: this thread's DoF line in the element before / after the pencil (stride in doubles between its nodes);
// gidx: index of this thread's face node in the ghost trace buffers of the direction.
template <int DIR>
__device__ __forceinline__ HaloRaw<4> q3p_halo_issue(const UniParams<4>& P, int fl, const double* __restrict__ prev,
                                                     const double* __restrict__ next, int stride, long gidx) {
  const int pm = (fl >> (2 * DIR)) & 1 ? P.bmode[2 * DIR] : 0;
  const int nm = (fl >> (2 * DIR + 1)) & 1 ? P.bmode[2 * DIR + 1] : 0;
  return halo_issue<4>(pm, nm, prev, next, stride, P.ghost[2 * DIR] + gidx * 2, P.ghost[2 * DIR + 1] + gidx * 2);
}

// NVLink peer-memory halo, sender side, fused into the tile kernel (multi-GPU): before their first tile the packing
// CTAs (a third of the grid) together compute the brick's boundary traces -- for brick face f = (d, s) and every boundary element and face node the
// pair (der, val) of the element's DoF line normal to the face at side s -- and store them straight into the neighbours'
// arenas; the last CTA to finish raises the neighbours' step flags.  No separate pack / flag kernels, no stream joins.
struct Q3pPack {
  double* out[6];  // receive buffer of the neighbour across face f (mapped peer memory), or null
  int* flag[6];    // that neighbour's step flag for the face
  int* done;       // local counter of CTAs that have finished packing
  int step;        // 0: no packing in this launch
  int npack;       // CTAs 0 .. npack-1 pack (a third of the grid: in launch order one per SM, so the pack of an SM's first CTA
                   // overlaps the first tiles of its other two)
};
__device__ __forceinline__ void q3p_pack(const UniParams<4>& P, const Q3pPack& K) {
  const double* __restrict__ x = P.x;
  const int n0 = P.n[0], n1 = P.n[1], n2 = P.n[2];
#pragma unroll 1
  for (int f = 0; f < 6; f++) {
    double* __restrict__ out = K.out[f];
    if (!out) continue;
    const int d = f >> 1, sd = f & 1;
    const int na = d == 0 ? n1 : n0, nb = d == 2 ? n1 : n2;  // face element extents (low dim fastest)
    const int total = na * nb * 16;
    for (int t = blockIdx.x * 256 + threadIdx.x; t < total; t += K.npack * 256) {
      const int node = t & 15, fe = t >> 4;
      const int a = fe % na, b = fe / na;
      const int pp = node & 3, q = node >> 2;
      long e; int off, stride;
      if (d == 0) { e = (sd ? n0 - 1 : 0) + (long)n0 * (a + (long)n1 * b); off = 4 * pp + 16 * q; stride = 1; }
      else if (d == 1) { e = a + (long)n0 * ((sd ? n1 - 1 : 0) + (long)n1 * b); off = pp + 16 * q; stride = 4; }
      else { e = a + (long)n0 * (b + (long)n1 * (sd ? n2 - 1 : 0)); off = pp + 4 * q; stride = 16; }
      const double* line = x + e * 64 + off;
      const double u0 = line[0], u1 = line[stride], u2 = line[2 * stride], u3 = line[3 * stride];
      double der;
      if (sd == 0) der = fma(P.g[0][0], u0, fma(P.g[0][1], u1, fma(P.g[0][2], u2, P.g[0][3] * u3)));
      else der = fma(P.g[1][0], u0, fma(P.g[1][1], u1, fma(P.g[1][2], u2, P.g[1][3] * u3)));
      reinterpret_cast<double2*>(out)[t] = make_double2(der, sd ? u3 : u0);
    }
  }
  // release: the CTA's peer stores are ordered before thread 0's system-scope fence by the barrier (cumulativity), that fence
  // before its count; the last packing CTA acquires every count before it raises the flags
  __syncthreads();
  if (threadIdx.x == 0 && (__threadfence_system(), atomicAdd(K.done, 1)) == K.npack - 1) {
    __threadfence_system();
    for (int f = 0; f < 6; f++)
      if (K.flag[f]) *reinterpret_cast<volatile int*>(K.flag[f]) = K.step;
    __threadfence_system();
    *K.done = 0;
  }
}


}  // namespace hpdg

// Tile descriptors (built once per level on the host): .x = index of the tile's first element, .y = tile coordinates
// tx | ty << 10 | tz << 20, .z = bit f set if the tile touches brick face f.
// extern "C": the table reads name the kernel's parameter symbol (<kernel>_param_0)
extern "C" __global__ void __launch_bounds__(256, 3)
hpdg_k_apply_q3_persist(const __grid_constant__ hpdg::UniParams<4> P, const int4* __restrict__ tile_desc, const int ntiles,
                        const int ntiles_total, int* __restrict__ sched, const __grid_constant__ hpdg::Q3pPack PK) {
  using namespace hpdg;
  constexpr int N = 4, N2 = 16, N3 = 64;
  extern __shared__ __align__(128) double q3p_sm[];
  double* __restrict__ su = q3p_sm;
  double* __restrict__ sw = q3p_sm + 4096;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(q3p_sm + 8192);
  volatile int* s_next = reinterpret_cast<volatile int*>(q3p_sm + 8193);
  const double* __restrict__ X = P.x;
  const int n0 = P.n[0], n01 = P.n[0] * P.n[1];  // element strides (in elements) in y and z

  auto descriptor = [&](int t) {
    int tb = P.tile_list ? P.tile_list[t] : t + P.tile_offset;
    if (P.tile_rot) { tb += P.tile_rot; if (tb >= ntiles_total) tb -= ntiles_total; }
    return __ldg(tile_desc + tb);
  };
  // The two halves of the CTA (element layers ez = 0,1 / 2,3) are independent between the first and the last pass: each
  // fetches its own half of the u tile, 8 rows of four x-contiguous elements (2 KB each)
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
  prefetch(threadIdx.x, td.x);   // the first tile's copies are in flight while the CTA packs
  if (PK.step > 0 && (int)blockIdx.x < PK.npack) q3p_pack(P, PK);
  uint32_t phase = 0;

  for (;;) {
    // dynamic tile scheduling: the CTA's first tile is blockIdx.x, the following ones come from a global counter (sched[0],
    // reset by the last CTA to finish), so that CTAs on less loaded SMs take more tiles and the last wave stays short.
    // Thread 0 draws the next tile here; the others read it after pass 3 (three barriers later).
    if (threadIdx.x == 0) *s_next = (int)gridDim.x + atomicAdd(sched, 1);
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

    // ---------------- P1: z-pencils; u from the prefetched tile, rewritten in place (swizzled) ----------------
    // z-role: node (i, j) of element column (ex, ey); a half warp = one column
    {
      const int tid = q3p_tid();
      const int zq = tid & 15, zex = (tid >> 4) & 3, zey = tid >> 6;
      const int zcol = (zex + 4 * zey) * N3;
      const double* colp = X + (long)(e0 + zex + n0 * zey) * N3 + zq;  // element (x, y, z0), this node
      const HaloRaw<4> hr = q3p_halo_issue<2>(P, fl, colp - (long)n01 * N3, colp + (long)n01 * (4 * N3), N2,
                                              ((long)(tx4 + zex) + (long)n0 * (ty4 + zey)) * N2 + zq);
      while (!q3p_mbar_try_wait(mbar, phase)) {}
      phase ^= 1;
      // in-place swizzle: every lane of the half warp reads its raw lines before any lane overwrites the column
      double v[4][4];
#pragma unroll
      for (int e = 0; e < 4; e++)
#pragma unroll
        for (int k = 0; k < 4; k++) v[e][k] = su[zcol + 1024 * e + 16 * k + zq];
      __syncwarp();
      if (P.xin_acc) {   // V-cycle: x += c fused into the apply of c
        double* __restrict__ xa = P.xin_acc + (colp - X);
#pragma unroll
        for (int e = 0; e < 4; e++)
#pragma unroll
          for (int k = 0; k < 4; k++) xa[(long)n01 * (N3 * e) + N2 * k] += v[e][k];
      }
      const HaloTrace h = halo_reduce<4>(P, hr);
      q3p_pencil<2>(h.pd, h.pv, h.pm, h.nd, h.nv, h.nm,
        [&](int e, double (&l)[4]) {
#pragma unroll
          for (int k = 0; k < 4; k++) l[k] = v[e][k];
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

    // ---------------- P2: x-pencils (128-bit shared-memory accesses) ----------------
    // x-role: line (j, k) of element row (ey, ez); a quarter warp = 4 j x 2 element layers
    {
      const int tid = q3p_tid();
      const int xj = tid & 3, xez = ((tid >> 2) & 1) | ((tid >> 6) & 2), xk = (tid >> 3) & 3, xey = (tid >> 5) & 3;
      const int xq = 2 * (xj ^ xk) + (xez & 1);
      const int xo0 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * (xq & 7);        // nodes i = 0,1 of the line (+ 64 e)
      const int xo1 = (4 * xey + 16 * xez) * N3 + 16 * xk + 2 * ((xq + 1) & 7);  // nodes i = 2,3
      const double* rowp = X + (long)(e0 + n0 * xey + n01 * xez) * N3 + N * xj + N2 * xk;  // element (x0, y, z), this line
      const HaloRaw<4> hr = q3p_halo_issue<0>(P, fl, rowp - N3, rowp + 4 * N3, 1,
                                              ((long)(ty4 + xey) + (long)P.n[1] * (tz4 + xez)) * N2 + xj + N * xk);
      __syncthreads();
      const HaloTrace h = halo_reduce<4>(P, hr);
      q3p_pencil<0>(h.pd, h.pv, h.pm, h.nd, h.nv, h.nm,
        [&](int e, double (&l)[4]) {
          const double2 lo = *reinterpret_cast<const double2*>(su + xo0 + 64 * e);
          const double2 hi = *reinterpret_cast<const double2*>(su + xo1 + 64 * e);
          l[0] = lo.x; l[1] = lo.y; l[2] = hi.x; l[3] = hi.y;
        },
        [&](int e, double (&a)[4]) {
          const double2 lo = *reinterpret_cast<const double2*>(sw + xo0 + 64 * e);
          const double2 hi = *reinterpret_cast<const double2*>(sw + xo1 + 64 * e);
          a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
        },
        [&](int e, const double (&)[4], const double (&a)[4]) {
          *reinterpret_cast<double2*>(sw + xo0 + 64 * e) = make_double2(a[0], a[1]);
          *reinterpret_cast<double2*>(sw + xo1 + 64 * e) = make_double2(a[2], a[3]);
        });
    }

    // ---------------- P3: y-pencils, then M_y ----------------
    // y-role: line (i, k) of element row (ex, ez); a half warp = 4 i x 4 k
    {
      const int tid = q3p_tid();
      const int yi = tid & 3, yk = (tid >> 2) & 3, yex = (tid >> 4) & 3, yez = tid >> 6;
      const int ybase = (yex + 16 * yez) * N3 + 16 * yk;
      const int yr = yi + 2 * (yez & 1);
      auto yo = [&](int j) { return ybase + (((4 * j) ^ (4 * yk)) + yr & 15); };  // node j of the line (+ 256 e)
      const double* colp = X + (long)(e0 + yex + n01 * yez) * N3 + yi + N2 * yk;  // element (x, y0, z), this line
      const HaloRaw<4> hr = q3p_halo_issue<1>(P, fl, colp - (long)n0 * N3, colp + (long)n0 * (4 * N3), N,
                                              ((long)(tx4 + yex) + (long)n0 * (tz4 + yez)) * N2 + yi + N * yk);
      q3p_bar_half(tid >> 7);
      const HaloTrace h = halo_reduce<4>(P, hr);
      q3p_pencil<1>(h.pd, h.pv, h.pm, h.nd, h.nv, h.nm,
        [&](int e, double (&l)[4]) {
#pragma unroll
          for (int j = 0; j < 4; j++) l[j] = su[yo(j) + 256 * e];
        },
        [&](int e, double (&a)[4]) {
#pragma unroll
          for (int j = 0; j < 4; j++) a[j] = sw[yo(j) + 256 * e];
        },
        [&](int e, const double (&)[4], const double (&a)[4]) {
          double b[4];
#pragma unroll
          for (int j = 0; j < 4; j++) b[j] = a[j];
          q3p_mass<false>(b);
#pragma unroll
          for (int j = 0; j < 4; j++) sw[yo(j) + 256 * e] = b[j];
        });
    }

    // this half's u rows are free: start the next tile's copies; they land during P4 and P5
    const int tn = *s_next;
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

#ifndef Q3P_NO_FUSE45
    // ---------------- P4+5: factor * M_z M_x on (x,z)-planes, 256-bit stores ----------------
    // plane role: x-z plane j of one element of this half; a quarter warp = 4 j x 2 element layers (the x-role's lanes, with
    // the x-role's k bits selecting the element of the row instead).  The 16 values of a plane cross shared memory once
    // (8 128-bit loads), both mass sweeps run in registers, and every lane stores whole 32-byte sectors: the four j lanes
    // of an element together write one full 128-byte line per z-plane.
    {
      const int tid = q3p_tid();
      const int pj = tid & 3, pez = ((tid >> 2) & 1) | ((tid >> 6) & 2), pex = (tid >> 3) & 3, pey = (tid >> 5) & 3;
      const int pbase = (pex + 4 * pey + 16 * pez) * N3;
      double b[4][4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int q = 2 * (pj ^ k) + (pez & 1);
        const double2 lo = *reinterpret_cast<const double2*>(sw + pbase + 16 * k + 2 * (q & 7));
        const double2 hi = *reinterpret_cast<const double2*>(sw + pbase + 16 * k + 2 * ((q + 1) & 7));
        double a[4] = {lo.x, lo.y, hi.x, hi.y};
        q3p_mass<false>(a);
#pragma unroll
        for (int i = 0; i < 4; i++) b[k][i] = a[i];
      }
      double o[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        double c[4] = {b[0][i], b[1][i], b[2][i], b[3][i]};
        q3p_mass<true>(c);
#pragma unroll
        for (int k = 0; k < 4; k++) o[k][i] = c[k];
      }
      double* __restrict__ yo_p = P.y + (long)(e0 + pex + n0 * pey + n01 * pez) * N3 + N * pj;
      if (P.accum) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          double y0, y1, y2, y3;
          asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];\n" : "=d"(y0), "=d"(y1), "=d"(y2), "=d"(y3) : "l"(yo_p + N2 * k));
          o[k][0] += y0; o[k][1] += y1; o[k][2] += y2; o[k][3] += y3;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};\n" ::"l"(yo_p + N2 * k), "d"(o[k][0]), "d"(o[k][1]), "d"(o[k][2]), "d"(o[k][3]) : "memory");
    }
    if (!has_next) break;
    __syncthreads();  // the plane role's reads of w precede the next tile's first-pass writes (different thread -> DoF mapping)
    t = tn;
    continue;
#else
    // ---------------- P4: M_x ----------------
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
        q3p_mass<false>(a);
        *reinterpret_cast<double2*>(sw + xo0 + 64 * e) = make_double2(a[0], a[1]);
        *reinterpret_cast<double2*>(sw + xo1 + 64 * e) = make_double2(a[2], a[3]);
      }
    }
    __syncthreads();

    // ---------------- P5: factor * M_z, coalesced store ----------------
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
          q3p_mass<true>(a);
          double* yo_e = yo_g + (long)(n01 * e) * N3;
#pragma unroll
          for (int k = 0; k < 4; k++) yo_e[N2 * k] = decltype(accum)::value ? yo_e[N2 * k] + a[k] : a[k];
        }
      };
      if (P.accum) tile_out(std::true_type{}); else tile_out(std::false_type{});
    }
#endif
    if (!has_next) break;
    t = tn;
  }
  if (threadIdx.x == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; __threadfence(); }
}
