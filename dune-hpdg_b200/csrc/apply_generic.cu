// Generic hp SIPG operator apply: elements bucketed by degree, any per-element degree 0..13, 2-D and 3-D.
//
// Replaces, for any per-element degree map, the reference's Operator::apply over an IPDGOperator
// (matrix-free/operator.hh:41-56, matrix-free/localoperators/ipdgoperator.hh:80-390) and equally the
// assembled DynamicBCRSMatrix::mv (common/matrixwindow.hh:196-209).  Formulation:
// element-centric "pull" -- each element computes its own rows from its own block and the face
// traces of its neighbours, so there is no scatter into neighbour rows (the reference's
// "not thread-safe" write, ipdgoperator.hh:233-243) and the summation order is fixed.
//
//   y_e = factor * (M x M x M) [ sum_d kappa_d (M^-1 S)_d u_e
//                                + sum_{faces (d,s)} ( mt_s (x) alpha_{d,s} + mg_s (x) beta_{d,s} ) ]
//
// with, per face node, alpha/beta built from the own trace (der_s, val_s) and -- through the
// L2 projection P = (M^{ee})^-1 M^{eo} in the tangential directions -- the neighbour's trace.
//
// Two kernels per apply:
//  * k_face_traces (ONE launch for the whole level): every element writes its OWN face traces once -- per face node the pair
//    (der, val) = (g_s . line, t_s . line) of the DoF line normal to the face -- into the level's trace array.
//  * k_apply_generic<DIM, N1> (one launch per degree bucket, buckets concurrent on side streams): a CTA takes several
//    elements of the bucket, N1^(DIM-1) threads per element (one per DoF line).  Neighbour traces are PULLED from the trace
//    array (N_o^(DIM-1) coalesced 16-byte loads per face) instead of being re-derived from the neighbour's full block; for
//    mixed-degree faces the tangential projection runs in two separable stages through shared memory.  Then, per direction,
//    a thread holds its DoF line in registers: own traces, alpha/beta of its two face nodes, T~_d sweep; finally the three
//    mass sweeps.  All index arithmetic on the element's own block is compile-time (DIM, N1 template parameters).
// The trace array is also exactly what crosses a rank boundary in the distributed hp path (csrc/api.cu).
#include <algorithm>
#include <cstdio>

#include "ctx.hpp"

namespace hpdg {

struct GenericParams {
  int dim;
  int n[3];
  double h[3];
  double sigma;
  int dirichlet;
  const int* deg;
  const int* pdeg;
  const long* off;
  const int* elist;
  long ebegin;  // bucket range in elist
  const DegTable* tab;
  const double* P;
  const double* x;
  double* y;
  double factor;
  int accum;
  const long* troff;     // per element: offset (in pairs) of its face traces, face f at troff[e] + f * N_e^(dim-1)
  const double2* tr;     // face traces (der, val) of every element (k_face_traces)
  // distributed hp (rank-local brick): faces on a rank boundary take the neighbour's degree and traces from the ghost layer
  int bnd_is_rank[6];
  const int* ghost_deg[6];      // [face element] degree of the remote element across brick face f
  const int* ghost_pdeg[6];     // [face element] its finest-level degree (face penalty)
  const long* ghost_troff[6];   // [face element] offset (pairs) of its traces in ghost_tr[f]
  const double2* ghost_tr[6];
  // face metadata of every element (generic_face_table): [element][2 dim][fslots]; non-conforming 2-D meshes: two intersections per
  // side (fslots = 2) and the sub-face couplings of the tangential bases (tables.hpp: Pnc_eo, Pnc_ee)
  const FaceInfo* finfo;
  int fslots;
  int tpe;   // threads per element of this launch (>= N1^(dim-1))
  const double* Pnc_eo;
  const double* Pnc_ee;
};

__host__ __device__ __forceinline__ int ipow_d(int b, int e) { int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }

template <int B, int E> struct CPow { static constexpr int v = B * CPow<B, E - 1>::v; };
template <int B> struct CPow<B, 0> { static constexpr int v = 1; };

// Base offset and stride (inside a block of N1^DIM doubles, x-fastest) of DoF line `line` along direction d.  The line index
// is also the face-node index of the line's end points on the two faces normal to d (tangential coordinates, lower
// direction fastest): line = j + N1 k (d = 0), i + N1 k (d = 1), i + N1 j (d = 2).
template <int DIM, int N1>
__device__ __forceinline__ int line_base_c(int d, int line) {
  if (d == 0) return N1 * line;
  if (DIM == 2) return line;
  return d == 1 ? (line % N1) + N1 * N1 * (line / N1) : line;
}
template <int N1> __device__ __forceinline__ int line_stride_c(int d) { return d == 0 ? 1 : (d == 1 ? N1 : N1 * N1); }

// ---- pass 1: own face traces of every element, one launch for all degree buckets --------------------------------------------
constexpr int kTraceThreads = 128;
struct TraceParams {
  int nb;                     // degree buckets
  int bucket_n1[kMaxN];
  long bucket_ebegin[kMaxN + 1];   // bucket ranges in elist
  long cta_begin[kMaxN + 1];       // first CTA of every bucket
  const int* elist; const long* off; const long* troff;
  const double* x; double2* tr;
};
// 1-D end-point tables (l_i'(s), l_i(s)) of all degrees, stride kMaxN, in constant memory
__constant__ double c_end_g[kMaxN][2][kMaxN];
__constant__ double c_end_t[kMaxN][2][kMaxN];

template <int DIM, int N1>
__device__ __forceinline__ void face_traces_body(const TraceParams& P, long ebegin, long cnt, long cta) {
  constexpr int nf = CPow<N1, DIM - 1>::v;
  constexpr int epc = nf >= kTraceThreads ? 1 : kTraceThreads / nf;
  for (int t = threadIdx.x; t < epc * nf; t += kTraceThreads) {
    const int lel = t / nf, line = t % nf;
    const long idx = cta * epc + lel;
    if (idx >= cnt) continue;
    const long e = P.elist[ebegin + idx];
    const double* __restrict__ xe = P.x + P.off[e];
    double2* __restrict__ out = P.tr + P.troff[e];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const int base = line_base_c<DIM, N1>(d, line), sd = line_stride_c<N1>(d);
      double v[N1];
#pragma unroll
      for (int k = 0; k < N1; k++) v[k] = __ldg(xe + base + k * sd);
      double d0 = 0, d1 = 0, v0 = 0, v1 = 0;
#pragma unroll
      for (int k = 0; k < N1; k++) {
        d0 = fma(c_end_g[N1 - 1][0][k], v[k], d0); d1 = fma(c_end_g[N1 - 1][1][k], v[k], d1);
        v0 = fma(c_end_t[N1 - 1][0][k], v[k], v0); v1 = fma(c_end_t[N1 - 1][1][k], v[k], v1);
      }
      out[(2 * d) * nf + line] = make_double2(d0, v0);
      out[(2 * d + 1) * nf + line] = make_double2(d1, v1);
    }
  }
}
template <int DIM>
__global__ void __launch_bounds__(kTraceThreads) k_face_traces(const __grid_constant__ TraceParams P) {
  int b = 0;
  while (b + 1 < P.nb && (long)blockIdx.x >= P.cta_begin[b + 1]) b++;
  const long cta = blockIdx.x - P.cta_begin[b], eb = P.bucket_ebegin[b], cnt = P.bucket_ebegin[b + 1] - eb;
  switch (P.bucket_n1[b]) {
#define HPDG_TR_CASE(NN) case NN: face_traces_body<DIM, NN>(P, eb, cnt, cta); break;
    HPDG_TR_CASE(1) HPDG_TR_CASE(2) HPDG_TR_CASE(3) HPDG_TR_CASE(4) HPDG_TR_CASE(5) HPDG_TR_CASE(6) HPDG_TR_CASE(7)
    HPDG_TR_CASE(8) HPDG_TR_CASE(9) HPDG_TR_CASE(10) HPDG_TR_CASE(11) HPDG_TR_CASE(12) HPDG_TR_CASE(13) HPDG_TR_CASE(14)
#undef HPDG_TR_CASE
  }
}

// ---- pass 2: the element kernel ------------------------------------------------------------------------------------------------
// 1-D tables of the bucket's degree, passed by value (constant bank)
template <int N1> struct GenTab { double MinvS[N1 * N1], M[N1 * N1], mt[2][N1], mg[2][N1], g[2][N1]; };

// Matrix-free block Gauss-Seidel (MODE 1): what the element pass needs beyond the operator apply
struct GsParams {
  const double* b;   // right-hand side
  double* x;         // iterate (== GenericParams::x), updated in place
  double2* tr;       // the level's trace array (== GenericParams::tr), the updated elements' traces are rewritten
};

// MODE 0: y_e = factor * (A x)_e.  MODE 1: one DynamicBlockGS row update of every element of the launch (the elements of one
// hyperplane and degree bucket, mutually independent): r_e = b_e - (A x)_e with the current x, x_e += (L_ee + D_ee)^-1 r_e
// (GSCore), then the element's own face traces are recomputed (blockgs_mf.cu).
template <int DIM, int N1, int MODE>
__device__ __forceinline__ void generic_element_pass(const GenericParams& P, const GenTab<N1>& T, const int maxno1,
                                                     const long cnt, const int epc, const int per_elem, const int mixed,
                                                     const GsParams& G) {
  extern __shared__ double sm_all[];
  constexpr int nfaces = 2 * DIM;
  constexpr int pe = N1 - 1;
  constexpr int ne = CPow<N1, DIM>::v;
  constexpr int nf = CPow<N1, DIM - 1>::v;   // face nodes == DoF lines per direction == threads per element
  const int maxtmp = N1 * maxno1;            // stage-1 projection results per face (3-D)
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long first = (long)blockIdx.x * epc;
  const int nel = (int)min((long)epc, cnt - first);   // elements of this CTA
  // per-element shared memory: su (block), sw (accumulator), face info, stage-1 projections (der / val) of the mixed faces
  auto SU = [&](int el) { return sm_all + (size_t)el * per_elem; };
  auto SW = [&](int el) { return SU(el) + ne; };
  const int fslots = P.fslots;
  auto FI = [&](int el) { return reinterpret_cast<FaceInfo*>(SW(el) + ne); };   // [face][slot]
  auto TD = [&](int el, int f) { return SW(el) + ne + 4 * nfaces * fslots + (size_t)(2 * f) * maxtmp; };
  auto TV = [&](int el, int f) { return TD(el, f) + maxtmp; };
  // non-conforming 2-D meshes: the element's own (der, val) traces of its 4 sides, [face][node][2] (the coarse side of a hanging
  // face mixes its face nodes); shares the place of the 3-D projection scratch
  auto OT = [&](int el, int f) { return SW(el) + ne + 4 * nfaces * fslots + (size_t)f * 2 * N1; };

  // ---- phase 0: load the blocks (coalesced), face metadata -------------------------------------------------------------
  for (int t = tid; t < nel * ne; t += nthr) {
    const int el = t / ne, i = t % ne;
    const long e = P.elist[P.ebegin + first + el];
    SU(el)[i] = __ldg(P.x + P.off[e] + i);
  }
  for (int t = tid; t < nel * nfaces * fslots; t += nthr) {
    const int el = t / (nfaces * fslots), q = t % (nfaces * fslots);
    const long e = P.elist[P.ebegin + first + el];
    FI(el)[q] = P.finfo[(size_t)e * nfaces * fslots + q];
  }
  __syncthreads();

  // threads per element: nf (one per DoF line), or a full warp for the low degrees of a 3-D mixed-degree level, whose tangential
  // projections (up to 2 x 6 x N1 x 14 x 14 FMAs per element) would otherwise be serialised over 4 or 9 threads
  const int tpe = P.tpe;
  const int lel = tid / tpe, line = tid % tpe;     // (element in CTA, DoF line / face node)
  const bool pact = lel < nel;                     // the thread belongs to an element of this CTA
  const bool lact = pact && line < nf;             // ... and owns a DoF line

  // ---- phase 1 (3-D, mixed-degree faces): stage 1 of the tangential projection, tmp[i + N1 b] = sum_a P[i][a] raw[a + No b] ----
  if (DIM == 3 && mixed) {
    if (pact) {
#pragma unroll 1
      for (int f = 0; f < nfaces; f++) {
        const FaceInfo F = FI(lel)[f * fslots];
        if (F.mode != 3) continue;
        const int no1 = F.po + 1;
        const double* __restrict__ Pm = P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
        const double2* __restrict__ raw = (F.ghost ? P.ghost_tr[f] : P.tr) + F.tro;
        double* __restrict__ td = TD(lel, f); double* __restrict__ tv = TV(lel, f);
        for (int q = line; q < N1 * no1; q += tpe) {
          const int i = q % N1, b = q / N1;
          double a0 = 0, a1 = 0;
#pragma unroll 4
          for (int a = 0; a < no1; a++) {
            const double pv = __ldg(Pm + i * kMaxN + a);
            const double2 rv = F.ghost ? __ldcg(raw + a + no1 * b) : __ldg(raw + a + no1 * b);
            a0 = fma(pv, rv.x, a0); a1 = fma(pv, rv.y, a1);
          }
          td[q] = a0; tv[q] = a1;
        }
      }
    }
    __syncthreads();
  }

  // ---- non-conforming 2-D meshes: the own traces of the element's four sides, for the coarse side of hanging faces ------------
  if (DIM == 2 && fslots == 2) {
    if (lact) {
      const double* su = SU(lel);
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const int base = line_base_c<DIM, N1>(d, line), sd = line_stride_c<N1>(d);
        double d0 = 0, d1 = 0;
#pragma unroll
        for (int k = 0; k < N1; k++) { const double v = su[base + k * sd]; d0 = fma(T.g[0][k], v, d0); d1 = fma(T.g[1][k], v, d1); }
        OT(lel, 2 * d)[2 * line] = d0; OT(lel, 2 * d)[2 * line + 1] = su[base];
        OT(lel, 2 * d + 1)[2 * line] = d1; OT(lel, 2 * d + 1)[2 * line + 1] = su[base + (N1 - 1) * sd];
      }
    }
    __syncthreads();
  }

  // ---- line passes: X: w = T~_x u   Y: w += T~_y u   Z: w += T~_z u, w = M_z w   then M_y, then M_x -> global ---------------
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    if (lact) {
      const double* su = SU(lel); double* sw = SW(lel);
      const int base = line_base_c<DIM, N1>(d, line), sd = line_stride_c<N1>(d);
      double v[N1], w[N1];
#pragma unroll
      for (int k = 0; k < N1; k++) v[k] = su[base + k * sd];
      double al[2], be[2];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const int f = 2 * d + s;
        const FaceInfo F = FI(lel)[f * fslots];
        double a = 0, b = 0;
        if (DIM == 2 && (F.kind == 1 || F.kind == 2)) {
          // coarse side of a hanging face (sfipdg.hh:472-491): one intersection per half of this side, each with its own fine
          // neighbour and penalty.  Face-node fields are L2 projections on the side: own terms through (M^ee)^-1 M^ee_half, the
          // neighbour's through (M^ee)^-1 M^eo_half; |F| = h_t / 2 doubles the penalty weight, h_n^o = h_n / 2 its derivative.
          for (int slot = 0; slot < 2; slot++) {
            const FaceInfo G = FI(lel)[f * fslots + slot];
            if (G.mode < 2) continue;
            const int no1 = G.po + 1, hk = G.kind - 1;
            const double* __restrict__ Pee = P.Pnc_ee + ((size_t)hk * (kMaxP + 1) + pe) * kMaxN * kMaxN + line * kMaxN;
            const double* __restrict__ Peo = P.Pnc_eo + (((size_t)hk * (kMaxP + 1) + pe) * (kMaxP + 1) + G.po) * kMaxN * kMaxN + line * kMaxN;
            const double* __restrict__ own = OT(lel, f);
            const double2* __restrict__ raw = P.tr + G.tro;
            double sa = 0, sb = 0;
            for (int j = 0; j < N1; j++) {
              const double pv = __ldg(Pee + j), De = own[2 * j], Ve = own[2 * j + 1];
              sa = fma(pv, fma(-0.5 * G.nuk, De, 2.0 * G.cpen * Ve), sa);
              sb = fma(pv, -0.5 * G.nuk * Ve, sb);
            }
            for (int a2 = 0; a2 < no1; a2++) {
              const double pv = __ldg(Peo + a2);
              const double2 rv = __ldg(raw + a2);
              sa = fma(pv, fma(-0.5 * G.nuk, 2.0 * rv.x, -2.0 * G.cpen * rv.y), sa);
              sb = fma(pv, 0.5 * G.nuk * rv.y, sb);
            }
            a += sa; b += sb;
          }
        } else if (F.mode != 0) {
          double der = 0;
#pragma unroll
          for (int k = 0; k < N1; k++) der = fma(T.g[s][k], v[k], der);
          const double val = v[s ? N1 - 1 : 0];   // GL nodes include the end points (p = 0: the constant)
          const double wnk = (F.mode == 1 ? 1.0 : 0.5) * F.nuk;
          a = fma(-wnk, der, F.cpen * val);
          b = -wnk * val;
          if (F.mode >= 2) {
            double nd, nv;
            const double2* __restrict__ raw = (F.ghost ? P.ghost_tr[f] : P.tr) + F.tro;
            if (F.mode == 2) {
              const double2 rv = F.ghost ? __ldcg(raw + line) : __ldg(raw + line);
              nd = rv.x; nv = rv.y;
            } else {
              const int no1 = F.po + 1;
              // fine side of a hanging face (kind 3 / 4): the whole own side against half of the coarse neighbour's side
              const double* __restrict__ Pm = (DIM == 2 && F.kind >= 3)
                  ? P.Pnc_eo + (((size_t)(F.kind - 1) * (kMaxP + 1) + pe) * (kMaxP + 1) + F.po) * kMaxN * kMaxN
                  : P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
              nd = 0; nv = 0;
              if (DIM == 2) {
#pragma unroll 4
                for (int a2 = 0; a2 < no1; a2++) {
                  const double pv = __ldg(Pm + line * kMaxN + a2);
                  const double2 rv = F.ghost ? __ldcg(raw + a2) : __ldg(raw + a2);
                  nd = fma(pv, rv.x, nd); nv = fma(pv, rv.y, nv);
                }
                if (F.kind >= 3) nd *= 0.5;   // h_n^o = 2 h_n: the neighbour's reference derivative in this element's scaling
              } else {
                const int i = line % N1, j = line / N1;
                const double* __restrict__ td = TD(lel, f); const double* __restrict__ tv = TV(lel, f);
#pragma unroll 4
                for (int b2 = 0; b2 < no1; b2++) {
                  const double pv = __ldg(Pm + j * kMaxN + b2);
                  nd = fma(pv, td[i + N1 * b2], nd); nv = fma(pv, tv[i + N1 * b2], nv);
                }
              }
            }
            a = fma(-0.5 * F.nuk, nd, a); a = fma(-F.cpen, nv, a);
            b = fma(0.5 * F.nuk, nv, b);
          }
        }
        al[s] = a; be[s] = b;
      }
      double kap = 1.0 / P.h[d];
#pragma unroll
      for (int dd = 0; dd < DIM; dd++) if (dd != d) kap *= P.h[dd];
#pragma unroll
      for (int i = 0; i < N1; i++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < N1; k++) s = fma(T.MinvS[i * N1 + k], v[k], s);
        s *= kap;
        s = fma(T.mt[0][i], al[0], s); s = fma(T.mg[0][i], be[0], s);
        s = fma(T.mt[1][i], al[1], s); s = fma(T.mg[1][i], be[1], s);
        w[i] = (d == 0) ? s : s + sw[base + i * sd];
      }
      if (d == DIM - 1) {  // last direction: its mass sweep acts on the same line
        double o[N1];
#pragma unroll
        for (int i = 0; i < N1; i++) {
          double s = 0;
#pragma unroll
          for (int k = 0; k < N1; k++) s = fma(T.M[i * N1 + k], w[k], s);
          o[i] = s;
        }
#pragma unroll
        for (int i = 0; i < N1; i++) w[i] = o[i];
      }
#pragma unroll
      for (int i = 0; i < N1; i++) sw[base + i * sd] = w[i];
    }
    __syncthreads();
  }
  // remaining mass sweeps, directions DIM-2 .. 0; the last one (x lines, contiguous) writes to global
#pragma unroll
  for (int d = DIM - 2; d >= 0; d--) {
    if (lact) {
      double* sw = SW(lel);
      const int base = line_base_c<DIM, N1>(d, line), sd = line_stride_c<N1>(d);
      double w[N1];
#pragma unroll
      for (int k = 0; k < N1; k++) w[k] = sw[base + k * sd];
      double* yo = d == 0 ? P.y + P.off[P.elist[P.ebegin + first + lel]] : nullptr;
#pragma unroll
      for (int i = 0; i < N1; i++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < N1; k++) s = fma(T.M[i * N1 + k], w[k], s);
        if (d == 0 && MODE == 0) yo[base + i] = P.accum ? yo[base + i] + P.factor * s : P.factor * s;
        else sw[base + i * sd] = s;
      }
    }
    if (d > 0 || MODE == 1) __syncthreads();
  }
  if (MODE == 0) return;

  // ---- block Gauss-Seidel epilogue (iterationsteps/dynamicblockgs.hh:94-126 with GSCore, :17-40) -------------------------
  // sw holds (A x)_e.  The diagonal block is never stored: A_ee = sum_d M x..x D_d x..x M with the 1-D factors D_d of the
  // element (own-face terms folded in, the same factors the block-Jacobi setup builds: jacobi.cu build_dir_matrix), so
  // entry (j, a) costs a handful of shared-memory reads.  GSCore from zero is the forward substitution (L + D) c = r.
  {
    constexpr int n2 = N1 * N1;
    double* sM = sm_all + (size_t)epc * per_elem;                                // [n2] 1-D mass (one copy per CTA)
    auto SD = [&](int el, int d) { return SW(el) + ne + 4 * nfaces * fslots + (mixed && DIM == 3 ? 2 * nfaces * maxtmp : (fslots == 2 ? 2 * nfaces * N1 : 0)) + d * n2; };
    auto SC = [&](int el) { return SD(el, DIM); };                               // [ne] correction
    for (int t = tid; t < n2; t += nthr) sM[t] = T.M[t];
    for (int t = tid; t < nel * DIM * n2; t += nthr) {
      const int el = t / (DIM * n2), d = (t / n2) % DIM, i = (t % n2) / N1, j = t % N1;
      const DegTable& DT = P.tab[pe];
      double kap = 1.0 / P.h[d];
      for (int dd = 0; dd < DIM; dd++) if (dd != d) kap *= P.h[dd];
      double v = kap * DT.S[i * kMaxN + j];
      for (int s = 0; s < 2; s++) {
        const FaceInfo F = FI(el)[(2 * d + s) * fslots];
        if (F.mode == 0) continue;
        const double wgt = F.mode == 1 ? 1.0 : 0.5;
        v += -wgt * F.nuk * (DT.t[s][i] * DT.g[s][j] + DT.g[s][i] * DT.t[s][j]) + F.cpen * DT.t[s][i] * DT.t[s][j];
      }
      SD(el, d)[i * N1 + j] = v;
    }
    long e_own = 0;
    if (lact) {
      e_own = P.elist[P.ebegin + first + lel];
      double* sw = SW(lel);
      const double* __restrict__ be = G.b + P.off[e_own];
#pragma unroll
      for (int k = 0; k < N1; k++) { const int j = line + nf * k; sw[j] = be[j] - sw[j]; }   // r_e = b_e - (A x)_e
    }
    const int j0 = line % N1, j1 = DIM == 3 ? line / N1 : 0;
    const bool warp_local = 32 % tpe == 0;   // an element's threads are lanes of ONE warp: no block barrier inside the substitution
    if (warp_local) __syncthreads();
    for (int a = 0; a < ne; a++) {
      if (warp_local) __syncwarp(); else __syncthreads();
      if (lact) {
        double* sw = SW(lel);
        const double* Dx = SD(lel, 0); const double* Dy = SD(lel, 1); const double* Dz = SD(lel, DIM - 1);
        const int a0 = a % N1, a1 = (a / N1) % N1, a2 = a / n2;
        double maa, c1, c2;
        if (DIM == 3) {
          maa = Dx[a0 * N1 + a0] * sM[a1 * N1 + a1] * sM[a2 * N1 + a2] + sM[a0 * N1 + a0] * Dy[a1 * N1 + a1] * sM[a2 * N1 + a2] +
                sM[a0 * N1 + a0] * sM[a1 * N1 + a1] * Dz[a2 * N1 + a2];
          c1 = Dx[j0 * N1 + a0] * sM[j1 * N1 + a1] + sM[j0 * N1 + a0] * Dy[j1 * N1 + a1];
          c2 = sM[j0 * N1 + a0] * sM[j1 * N1 + a1];
        } else {
          maa = Dx[a0 * N1 + a0] * sM[a1 * N1 + a1] + sM[a0 * N1 + a0] * Dy[a1 * N1 + a1];
          c1 = Dx[j0 * N1 + a0]; c2 = sM[j0 * N1 + a0];
        }
        const double xa = (fabs(maa) == 0.) ? 0.0 : sw[a] / maa;               // dynamicblockgs.hh:26-27,36
        if (a % nf == line) SC(lel)[a] = xa;
        const int ak = DIM == 3 ? a2 : a1;                                       // the column's index in the owned direction
#pragma unroll
        for (int k = 0; k < N1; k++) {
          const int j = line + nf * k;
          if (j > a) {
            const double mja = DIM == 3 ? sM[k * N1 + ak] * c1 + Dz[k * N1 + ak] * c2 : sM[k * N1 + ak] * c1 + Dy[k * N1 + ak] * c2;
            sw[j] = fma(-mja, xa, sw[j]);
          }
        }
      }
    }
    __syncthreads();
    if (lact) {
      double* su = SU(lel); const double* sc = SC(lel);
      double* __restrict__ xe = G.x + P.off[e_own];
#pragma unroll
      for (int k = 0; k < N1; k++) { const int j = line + nf * k; const double xn = su[j] + sc[j]; su[j] = xn; xe[j] = xn; }
    }
    __syncthreads();
    if (lact) {   // the element's own face traces follow its new values (the later hyperplanes pull them)
      const double* su = SU(lel);
      double2* __restrict__ out = G.tr + P.troff[e_own];
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const int base = line_base_c<DIM, N1>(d, line), sd = line_stride_c<N1>(d);
        double d0 = 0, d1 = 0, v0 = 0, v1 = 0;
#pragma unroll
        for (int k = 0; k < N1; k++) {
          const double v = su[base + k * sd];
          d0 = fma(c_end_g[N1 - 1][0][k], v, d0); d1 = fma(c_end_g[N1 - 1][1][k], v, d1);
          v0 = fma(c_end_t[N1 - 1][0][k], v, v0); v1 = fma(c_end_t[N1 - 1][1][k], v, v1);
        }
        out[(2 * d) * nf + line] = make_double2(d0, v0);
        out[(2 * d + 1) * nf + line] = make_double2(d1, v1);
      }
    }
  }
}

template <int DIM, int N1>
__global__ void k_apply_generic(const __grid_constant__ GenericParams P, const __grid_constant__ GenTab<N1> T, const int maxno1,
                                const long cnt, const int epc, const int per_elem, const int mixed) {
  generic_element_pass<DIM, N1, 0>(P, T, maxno1, cnt, epc, per_elem, mixed, GsParams());
}
template <int DIM, int N1>
__global__ void k_blockgs_generic(const __grid_constant__ GenericParams P, const __grid_constant__ GenTab<N1> T, const int maxno1,
                                  const long cnt, const int epc, const int per_elem, const int mixed, const GsParams G) {
  generic_element_pass<DIM, N1, 1>(P, T, maxno1, cnt, epc, per_elem, mixed, G);
}

// the level's trace array (allocated on first use) and the offsets of every element's traces
int generic_trace_setup(Ctx* ctx, Level& L) {
  if (L.d_troff) return 0;
  std::vector<long> troff(L.nelem + 1, 0);
  for (long e = 0; e < L.nelem; e++) troff[e + 1] = troff[e] + 2 * L.dim * ipow_d(L.deg[e] + 1, L.dim - 1);
  L.tr_pairs = troff[L.nelem];
  L.troff_h = troff;
  HPDG_CUDA(cudaMalloc(&L.d_troff, sizeof(long) * (L.nelem + 1)));
  HPDG_CUDA(cudaMemcpy(L.d_troff, troff.data(), sizeof(long) * (L.nelem + 1), cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMalloc(&L.d_tr, sizeof(double) * 2 * std::max<long>(L.tr_pairs, 1)));
  if (!ctx->end_tables_set) {  // once per context (= device): the end-point tables of all degrees in constant memory
    static thread_local double g[kMaxN][2][kMaxN], t[kMaxN][2][kMaxN];
    for (int p = 0; p <= kMaxP; p++) for (int s = 0; s < 2; s++) for (int k = 0; k < kMaxN; k++) {
      g[p][s][k] = k <= p ? host_tables().deg[p].g[s][k] : 0.0; t[p][s][k] = k <= p ? host_tables().deg[p].t[s][k] : 0.0;
    }
    HPDG_CUDA(cudaMemcpyToSymbol(c_end_g, g, sizeof(g)));
    HPDG_CUDA(cudaMemcpyToSymbol(c_end_t, t, sizeof(t)));
    ctx->end_tables_set = true;
  }
  return 0;
}

// The level's face table: FaceInfo of every (element, side, intersection), built once on the host.  Structured meshes: one
// intersection per side, the neighbour from the lexicographic element index; rank-boundary sides of a distributed level point
// into the ghost layer (degrees exchanged once, hp_ghost_setup).  Non-conforming meshes: the table comes from the mesh builder
// (api.cu: hpdg_create_refined_2d), only the trace offsets are filled in here.
int generic_face_table(Ctx* ctx, Level& L) {
  if (L.d_finfo) return 0;
  if (generic_trace_setup(ctx, L)) return 1;
  const int dim = L.dim, nfaces = 2 * dim;
  L.fslots = L.nc ? 2 : 1;
  std::vector<FaceInfo> tab((size_t)L.nelem * nfaces * L.fslots);
  if (L.nc) {
    tab = L.nc_faces;
    for (long e = 0; e < L.nelem; e++) for (int q = 0; q < nfaces * 2; q++) {
      FaceInfo& F = tab[(size_t)e * nfaces * 2 + q];
      const long o = L.nc_nbr[(size_t)e * nfaces * 2 + q];
      if (F.mode >= 2 && o >= 0) F.tro = L.troff_h[o] + (long)((q / 2) ^ 1) * ipow_d(L.deg[o] + 1, dim - 1);
    }
  } else {
    for (long e = 0; e < L.nelem; e++) {
      long r = e; int ijk[3];
      ijk[0] = (int)(r % L.n[0]); r /= L.n[0]; ijk[1] = (int)(r % L.n[1]); r /= L.n[1]; ijk[2] = (int)r;
      const int pe = L.deg[e];
      for (int f = 0; f < nfaces; f++) {
        const int d = f / 2, s = f % 2;
        const int c = ijk[d] + (s ? 1 : -1);
        FaceInfo F;
        double kappa = 1.0 / L.h[d];
        for (int dd = 0; dd < dim; dd++) if (dd != d) kappa *= L.h[dd];
        F.nuk = s ? kappa : -kappa;
        F.po = (short)pe; F.tro = 0; F.ghost = 0; F.kind = 0;
        if (c >= 0 && c < L.n[d]) {
          const long stride = d == 0 ? 1 : d == 1 ? L.n[0] : (long)L.n[0] * L.n[1];
          const long o = e + (s ? stride : -stride);
          F.po = (short)L.deg[o];
          F.tro = L.troff_h[o] + (long)(2 * d + (1 - s)) * ipow_d(F.po + 1, dim - 1);
          const int pm = std::max(L.pdeg[e], L.pdeg[o]);
          F.cpen = ctx->sigma * (double)pm * pm;
          F.mode = F.po == pe ? 2 : 3;
        } else if (ctx->nranks > 1 && ctx->bnd_is_rank[f]) {
          // face elements of a brick face are numbered with the lower tangential direction fastest
          const int ta = d == 0 ? 1 : 0, tb = d == 2 ? 1 : 2;
          const long fe = ijk[ta] + (long)L.n[ta] * (dim == 3 ? ijk[tb] : 0);
          F.ghost = 1;
          F.po = (short)L.hpg.h_deg[f][fe]; F.tro = L.hpg.h_troff[f][fe];
          const int pm = std::max(L.pdeg[e], L.hpg.h_pdeg[f][fe]);
          F.cpen = ctx->sigma * (double)pm * pm;
          F.mode = F.po == pe ? 2 : 3;
        } else {
          F.cpen = ctx->sigma * (double)L.pdeg[e] * L.pdeg[e];
          F.mode = ctx->dirichlet ? 1 : 0;
        }
        tab[(size_t)e * nfaces + f] = F;
      }
    }
  }
  HPDG_CUDA(cudaMalloc(&L.d_finfo, sizeof(FaceInfo) * tab.size()));
  HPDG_CUDA(cudaMemcpy(L.d_finfo, tab.data(), sizeof(FaceInfo) * tab.size(), cudaMemcpyHostToDevice));
  if (L.nc && !ctx->d_Pnc_eo) {
    const HostTables& H = host_tables();
    HPDG_CUDA(cudaMalloc(&ctx->d_Pnc_eo, sizeof(double) * H.Pnc_eo.size()));
    HPDG_CUDA(cudaMemcpy(ctx->d_Pnc_eo, H.Pnc_eo.data(), sizeof(double) * H.Pnc_eo.size(), cudaMemcpyHostToDevice));
    HPDG_CUDA(cudaMalloc(&ctx->d_Pnc_ee, sizeof(double) * H.Pnc_ee.size()));
    HPDG_CUDA(cudaMemcpy(ctx->d_Pnc_ee, H.Pnc_ee.data(), sizeof(double) * H.Pnc_ee.size(), cudaMemcpyHostToDevice));
  }
  return 0;
}

// pass 1 of the hp apply (also the first step of the distributed hp halo): own face traces of every element of the level
int launch_face_traces(Ctx* ctx, Level& L, const double* x, cudaStream_t stream) {
  if (generic_trace_setup(ctx, L)) return 1;
  static thread_local TraceParams TP;
  TP.nb = (int)L.bucket_p.size();
  long ctas = 0;
  for (int b = 0; b < TP.nb; b++) {
    const int n1 = L.bucket_p[b] + 1, nf = ipow_d(n1, L.dim - 1);
    const int epc = nf >= kTraceThreads ? 1 : kTraceThreads / nf;
    TP.bucket_n1[b] = n1; TP.bucket_ebegin[b] = L.bucket_begin[b]; TP.cta_begin[b] = ctas;
    ctas += (L.bucket_begin[b + 1] - L.bucket_begin[b] + epc - 1) / epc;
  }
  TP.bucket_ebegin[TP.nb] = L.bucket_begin[TP.nb]; TP.cta_begin[TP.nb] = ctas;
  TP.elist = L.d_elist; TP.off = L.d_off; TP.troff = L.d_troff; TP.x = x; TP.tr = reinterpret_cast<double2*>(L.d_tr);
  if (ctas == 0) return 0;
  if (L.dim == 2) k_face_traces<2><<<(unsigned)ctas, kTraceThreads, 0, stream>>>(TP);
  else k_face_traces<3><<<(unsigned)ctas, kTraceThreads, 0, stream>>>(TP);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

// distributed hp: gather the face-f traces of this rank's boundary elements into the contiguous send buffer of brick face f
struct HpPackParams { const double2* tr; double2* send[6]; const long* src[6]; const long* dst[6]; long nface[6]; };
__global__ void k_hp_pack(const __grid_constant__ HpPackParams P) {
  const int f = blockIdx.y;
  if (!P.send[f]) return;
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long)gridDim.x * (blockDim.x >> 5);
  for (long fe = warp; fe < P.nface[f]; fe += nw) {
    const long s0 = P.src[f][fe], d0 = P.dst[f][fe], n = P.dst[f][fe + 1] - d0;
    for (long i = lane; i < n; i += 32) P.send[f][d0 + i] = P.tr[s0 + i];
  }
}
int launch_hp_pack(Ctx* ctx, Level& L, cudaStream_t stream) {
  HpPackParams PK;
  long mx = 0;
  for (int f = 0; f < 6; f++) {
    PK.send[f] = reinterpret_cast<double2*>(L.hpg.d_send[f]); PK.src[f] = L.hpg.d_send_src[f]; PK.dst[f] = L.hpg.d_send_dst[f];
    PK.nface[f] = L.hpg.nface[f];
    if (PK.send[f]) mx = std::max(mx, PK.nface[f]);
  }
  PK.tr = reinterpret_cast<const double2*>(L.d_tr);
  if (mx == 0) return 0;
  k_hp_pack<<<dim3((unsigned)std::min<long>((mx + 7) / 8, 1024), 6), 256, 0, stream>>>(PK);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

int launch_apply_generic(Ctx* ctx, Level& L, const double* x, double* y, double factor) {
  static thread_local GenericParams P;
  P.dim = L.dim;
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.h[d] = L.h[d]; }
  P.sigma = ctx->sigma; P.dirichlet = ctx->dirichlet;
  P.deg = L.d_deg; P.pdeg = L.d_pdeg; P.off = L.d_off; P.elist = L.d_elist;
  P.tab = ctx->d_tab; P.P = ctx->d_P; P.x = x; P.y = y; P.factor = factor; P.accum = ctx->fuse_accum;
  // ---- pass 1: every element's own face traces (one launch); distributed hp: exchange the rank-boundary traces ----
  if (launch_face_traces(ctx, L, x, ctx->stream)) return 1;
  P.troff = L.d_troff; P.tr = reinterpret_cast<const double2*>(L.d_tr);
  int maxno1 = L.maxp + 1;
  for (int f = 0; f < 6; f++) { P.bnd_is_rank[f] = 0; P.ghost_deg[f] = P.ghost_pdeg[f] = nullptr; P.ghost_troff[f] = nullptr; P.ghost_tr[f] = nullptr; }
  if (ctx->nranks > 1) {
    if (hp_halo_exchange(ctx, L)) return 1;
    for (int f = 0; f < 6; f++) {
      if (!ctx->bnd_is_rank[f]) continue;
      P.bnd_is_rank[f] = 1; P.ghost_deg[f] = L.hpg.d_deg[f]; P.ghost_pdeg[f] = L.hpg.d_pdeg[f]; P.ghost_troff[f] = L.hpg.d_troff[f];
      P.ghost_tr[f] = reinterpret_cast<const double2*>(L.hpg.d_recv[f]);
    }
    maxno1 = std::max(maxno1, L.hpg.maxp + 1);
  }
  const int mixed = (!L.uniform || ctx->nranks > 1 || L.nc) ? 1 : 0;
  if (generic_face_table(ctx, L)) return 1;
  P.finfo = L.d_finfo; P.fslots = L.fslots; P.Pnc_eo = ctx->d_Pnc_eo; P.Pnc_ee = ctx->d_Pnc_ee;
  // ---- pass 2: the degree buckets write disjoint rows of y: side streams, so that small buckets overlap the large ones ----
  const size_t nb = L.bucket_p.size();
  const bool fork = nb > 1;
  int nstreams = 1;
  if (fork) {
    if (!ctx->bucket_stream[0]) {
      for (int k = 0; k < kBucketStreams; k++) HPDG_CUDA(cudaStreamCreateWithFlags(&ctx->bucket_stream[k], cudaStreamNonBlocking));
      for (int k = 0; k <= kBucketStreams; k++) HPDG_CUDA(cudaEventCreateWithFlags(&ctx->bucket_ev[k], cudaEventDisableTiming));
    }
    nstreams = (int)std::min<size_t>(nb, kBucketStreams);
    HPDG_CUDA(cudaEventRecord(ctx->bucket_ev[kBucketStreams], ctx->stream));
    for (int k = 0; k < nstreams; k++) HPDG_CUDA(cudaStreamWaitEvent(ctx->bucket_stream[k], ctx->bucket_ev[kBucketStreams], 0));
  }
  // largest buckets first (they determine the critical path)
  std::vector<size_t> order(nb);
  for (size_t b = 0; b < nb; b++) order[b] = b;
  auto work = [&](size_t k) { return (double)(L.bucket_begin[k + 1] - L.bucket_begin[k]) * ipow_d(L.bucket_p[k] + 1, L.dim + 1); };
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return work(a) > work(b); });
  for (size_t bi = 0; bi < nb; bi++) {
    const size_t b = order[bi];
    long cnt = L.bucket_begin[b + 1] - L.bucket_begin[b];
    if (cnt == 0) continue;
    cudaStream_t lstream = fork ? ctx->bucket_stream[bi % nstreams] : ctx->stream;
    const int p = L.bucket_p[b], n1 = p + 1;
    const int ne = ipow_d(n1, L.dim), nf = ipow_d(n1, L.dim - 1);
    const int nfaces = 2 * L.dim;
    const int per_elem = 2 * ne + 4 * nfaces * L.fslots + ((mixed && L.dim == 3) ? 2 * nfaces * n1 * maxno1 : (L.nc ? 2 * nfaces * n1 : 0));
    P.ebegin = L.bucket_begin[b];
    // elements per CTA: ~128 threads, at most 48 KB of shared memory (256-thread CTAs for the high degrees, i.e. fewer idle lanes in
    // the last warp but more elements per block barrier, measured slower: cfg3 192.9 against 187.9 us)
    const int tpe = (mixed && L.dim == 3 && nf < 32) ? 32 : nf;
    P.tpe = tpe;
    int epc = tpe >= 128 ? 1 : 128 / tpe;
    epc = (int)std::max<long>(1, std::min<long>(epc, (48 * 1024) / ((long)per_elem * 8)));
    epc = (int)std::min<long>(epc, cnt);
    const size_t smem_l = (size_t)epc * per_elem * sizeof(double);
    const unsigned grid = (unsigned)((cnt + epc - 1) / epc);
    const int threads = std::max(32, (epc * tpe + 31) / 32 * 32);
    const DegTable& HT = host_tables().deg[p];
#define HPDG_GEN_LAUNCH(D, NN)                                                                                            \
  do {                                                                                                                    \
    static thread_local GenTab<NN> T;                                                                                     \
    for (int i = 0; i < NN; i++) {                                                                                        \
      for (int j = 0; j < NN; j++) { T.MinvS[i * NN + j] = HT.MinvS[i * kMaxN + j]; T.M[i * NN + j] = HT.M[i * kMaxN + j]; } \
      for (int sd = 0; sd < 2; sd++) { T.mt[sd][i] = HT.mt[sd][i]; T.mg[sd][i] = HT.mg[sd][i]; T.g[sd][i] = HT.g[sd][i]; } \
    }                                                                                                                     \
    if (smem_l > 48 * 1024)                                                                                               \
      HPDG_CUDA(cudaFuncSetAttribute(k_apply_generic<D, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));  \
    k_apply_generic<D, NN><<<grid, threads, smem_l, lstream>>>(P, T, maxno1, cnt, epc, per_elem, mixed);                  \
  } while (0)
#define HPDG_GEN_CASE(NN)                                                      \
  case NN:                                                                     \
    if (L.dim == 2) HPDG_GEN_LAUNCH(2, NN); else HPDG_GEN_LAUNCH(3, NN);       \
    break;
    switch (n1) {
      HPDG_GEN_CASE(1) HPDG_GEN_CASE(2) HPDG_GEN_CASE(3) HPDG_GEN_CASE(4) HPDG_GEN_CASE(5) HPDG_GEN_CASE(6) HPDG_GEN_CASE(7)
      HPDG_GEN_CASE(8) HPDG_GEN_CASE(9) HPDG_GEN_CASE(10) HPDG_GEN_CASE(11) HPDG_GEN_CASE(12) HPDG_GEN_CASE(13) HPDG_GEN_CASE(14)
      default: ctx->err = "degree out of range"; return 1;
    }
#undef HPDG_GEN_CASE
#undef HPDG_GEN_LAUNCH
    ctx->launches++;
    HPDG_CUDA(cudaGetLastError());
  }
  if (fork) {
    for (int k = 0; k < nstreams; k++) {
      HPDG_CUDA(cudaEventRecord(ctx->bucket_ev[k], ctx->bucket_stream[k]));
      HPDG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->bucket_ev[k], 0));
    }
  }
  return 0;
}

// ---- matrix-free block Gauss-Seidel: DynamicBlockGS::iterate (iterationsteps/dynamicblockgs.hh:94-126) without a matrix ----------
// The reference sweeps the block rows in ascending element order: r_i = b_i - sum_j A_ij x_j with the current x, then
// x_i += GSCore(A_ii, r_i).  Elements on a hyperplane ix + iy + iz = const are never face neighbours and depend only on lower
// hyperplanes, so sweeping hyperplane by hyperplane (elements of one hyperplane in parallel) reproduces the sequential sweep
// with the same operands for every row -- as the assembled hpdg_blockgs_iterate does (assemble.cu), but (A x)_i comes from the
// element pass of the operator kernel (own block + the neighbours' current face traces) and A_ii from its Kronecker factors:
// no matrix, so the reference's default smoother runs at the named mesh sizes (the assembled one needs 229 KB per Q3 element).
static int blockgs_mf_setup(Ctx* ctx, Level& L) {
  if (L.d_gs_elist) return 0;
  const int nb = (int)L.bucket_p.size();
  const int nw = L.n[0] + L.n[1] + (L.dim == 3 ? L.n[2] : 1) - 2;
  std::vector<int> bucket_of(kMaxP + 1, -1);
  for (int b = 0; b < nb; b++) bucket_of[L.bucket_p[b]] = b;
  std::vector<long> cntv((size_t)nw * nb + 1, 0);
  auto seg = [&](long e) {
    long r = e; const int i0 = (int)(r % L.n[0]); r /= L.n[0]; const int i1 = (int)(r % L.n[1]); r /= L.n[1];
    return (size_t)(i0 + i1 + (int)r) * nb + bucket_of[L.deg[e]];
  };
  for (long e = 0; e < L.nelem; e++) cntv[seg(e) + 1]++;
  for (size_t k = 0; k < (size_t)nw * nb; k++) cntv[k + 1] += cntv[k];
  L.gs_seg = cntv;
  std::vector<int> list(L.nelem);
  std::vector<long> pos(cntv.begin(), cntv.end() - 1);
  for (long e = 0; e < L.nelem; e++) list[pos[seg(e)]++] = (int)e;   // ascending element index inside a segment
  HPDG_CUDA(cudaMalloc(&L.d_gs_elist, sizeof(int) * std::max<long>(L.nelem, 1)));
  HPDG_CUDA(cudaMemcpy(L.d_gs_elist, list.data(), sizeof(int) * L.nelem, cudaMemcpyHostToDevice));
  L.gs_nw = nw;
  return 0;
}

template <int NN>
static void fill_gentab(GenTab<NN>& T, const DegTable& HT) {
  for (int i = 0; i < NN; i++) {
    for (int j = 0; j < NN; j++) { T.MinvS[i * NN + j] = HT.MinvS[i * kMaxN + j]; T.M[i * NN + j] = HT.M[i * kMaxN + j]; }
    for (int sd = 0; sd < 2; sd++) { T.mt[sd][i] = HT.mt[sd][i]; T.mg[sd][i] = HT.mg[sd][i]; T.g[sd][i] = HT.g[sd][i]; }
  }
}

int blockgs_mf_iterate(Ctx* ctx, Level& L, const double* b, double* x) {
  if (ctx->nranks > 1) { ctx->err = "matrix-free block Gauss-Seidel is rank-local (use the L1 smoother or block Jacobi across ranks)"; return 1; }
  if (blockgs_mf_setup(ctx, L)) return 1;
  static thread_local GenericParams P;
  P.dim = L.dim;
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.h[d] = L.h[d]; }
  P.sigma = ctx->sigma; P.dirichlet = ctx->dirichlet;
  P.deg = L.d_deg; P.pdeg = L.d_pdeg; P.off = L.d_off; P.elist = L.d_gs_elist;
  P.tab = ctx->d_tab; P.P = ctx->d_P; P.x = x; P.y = nullptr; P.factor = 1.0; P.accum = 0;
  if (launch_face_traces(ctx, L, x, ctx->stream)) return 1;   // traces of the incoming iterate; the sweep keeps them current
  P.troff = L.d_troff; P.tr = reinterpret_cast<const double2*>(L.d_tr);
  for (int f = 0; f < 6; f++) { P.bnd_is_rank[f] = 0; P.ghost_deg[f] = P.ghost_pdeg[f] = nullptr; P.ghost_troff[f] = nullptr; P.ghost_tr[f] = nullptr; }
  if (L.nc) { ctx->err = "matrix-free block Gauss-Seidel: not available on non-conforming meshes"; return 1; }
  if (generic_face_table(ctx, L)) return 1;
  P.finfo = L.d_finfo; P.fslots = L.fslots; P.Pnc_eo = ctx->d_Pnc_eo; P.Pnc_ee = ctx->d_Pnc_ee;
  GsParams G; G.b = b; G.x = x; G.tr = reinterpret_cast<double2*>(L.d_tr);
  const int maxno1 = L.maxp + 1;
  const int mixed = L.uniform ? 0 : 1;
  const int nb = (int)L.bucket_p.size();
  for (int w = 0; w < L.gs_nw; w++) {
    for (int bk = 0; bk < nb; bk++) {
      const long begin = L.gs_seg[(size_t)w * nb + bk], cnt = L.gs_seg[(size_t)w * nb + bk + 1] - begin;
      if (cnt == 0) continue;
      const int p = L.bucket_p[bk], n1 = p + 1;
      const int ne = ipow_d(n1, L.dim), nf = ipow_d(n1, L.dim - 1), nfaces = 2 * L.dim;
      const int per_elem = 2 * ne + 4 * nfaces + ((mixed && L.dim == 3) ? 2 * nfaces * n1 * maxno1 : 0);
      const int per_elem_gs = per_elem + L.dim * n1 * n1 + ne;
      P.ebegin = begin;
      const int tpe = (mixed && L.dim == 3 && nf < 32) ? 32 : nf;
      P.tpe = tpe;
      int epc = tpe >= 128 ? 1 : 128 / tpe;
      epc = (int)std::max<long>(1, std::min<long>(epc, (48 * 1024) / ((long)per_elem_gs * 8)));
      epc = (int)std::min<long>(epc, cnt);
      const size_t smem_l = ((size_t)epc * per_elem_gs + n1 * n1) * sizeof(double);
      const unsigned grid = (unsigned)((cnt + epc - 1) / epc);
      const int threads = std::max(32, (epc * tpe + 31) / 32 * 32);
      if (threads > 1024) { ctx->err = "matrix-free block Gauss-Seidel: degree too high for one CTA per element"; return 1; }
      const DegTable& HT = host_tables().deg[p];
#define HPDG_GS_LAUNCH(D, NN)                                                                                             \
  do {                                                                                                                    \
    static thread_local GenTab<NN> T;                                                                                     \
    fill_gentab<NN>(T, HT);                                                                                               \
    if (kernel_slots(ctx, reinterpret_cast<const void*>(k_blockgs_generic<D, NN>), threads, 200 * 1024, nullptr)) return 1; \
    k_blockgs_generic<D, NN><<<grid, threads, smem_l, ctx->stream>>>(P, T, maxno1, cnt, epc, per_elem_gs, mixed, G);      \
  } while (0)
#define HPDG_GS_CASE(NN)                                                     \
  case NN:                                                                   \
    if (L.dim == 2) HPDG_GS_LAUNCH(2, NN); else HPDG_GS_LAUNCH(3, NN);       \
    break;
      switch (n1) {
        HPDG_GS_CASE(1) HPDG_GS_CASE(2) HPDG_GS_CASE(3) HPDG_GS_CASE(4) HPDG_GS_CASE(5) HPDG_GS_CASE(6) HPDG_GS_CASE(7)
        HPDG_GS_CASE(8) HPDG_GS_CASE(9) HPDG_GS_CASE(10) HPDG_GS_CASE(11) HPDG_GS_CASE(12) HPDG_GS_CASE(13) HPDG_GS_CASE(14)
        default: ctx->err = "degree out of range"; return 1;
      }
#undef HPDG_GS_CASE
#undef HPDG_GS_LAUNCH
      ctx->launches++;
    }
  }
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hpdg
