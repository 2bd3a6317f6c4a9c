// Generic hp SIPG operator apply: one CTA per element, elements bucketed by degree.
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
};

__device__ __forceinline__ int ipow_d(int b, int e) { int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }

// Per-face metadata of the element a CTA works on (computed once by 2*dim threads, read by all).
struct FaceInfo {
  int has_nb;      // neighbour element exists
  int skip;        // natural boundary: no face term at all (ipdgoperator.hh:97-105)
  int po;          // neighbour degree
  long uo;         // offset of the neighbour's block
  double w_nu_k;   // w * nu * kappa   (w = 1/2 interior, 1 Dirichlet: ipdgoperator.hh:186,357)
  double cpen;     // sigma * max(p-,p+)^2 (ipdgoperator.hh:129-131) or sigma p^2 on the boundary (:310)
  double A1, A2, A3;
};

// base offset (inside a block of n1^dim doubles) of the DoF line normal to direction d through face node `node`
__device__ __forceinline__ int line_base(int dim, int n1, int d, int node) {
  int rem = node, base = 0, st = 1;
  for (int dd = 0; dd < dim; dd++) {
    if (dd != d) { base += (rem % n1) * st; rem /= n1; }
    st *= n1;
  }
  return base;
}


// sum_k a[k*sa] * b[k*sb] for k < n, n compile-time for the common degrees so the loads pipeline
template <int NN>
__device__ __forceinline__ void dot2_n(const double* __restrict__ c0, const double* __restrict__ c1, int sc, const double* __restrict__ v,
                                       int sv, double& r0, double& r1) {
  double a = 0, b = 0;
#pragma unroll
  for (int k = 0; k < NN; k++) { const double x = v[k * sv]; a = fma(c0[k * sc], x, a); b = fma(c1[k * sc], x, b); }
  r0 = a; r1 = b;
}
__device__ __forceinline__ void dot2(int n, const double* __restrict__ c0, const double* __restrict__ c1, int sc,
                                     const double* __restrict__ v, int sv, double& r0, double& r1) {
  switch (n) {
    case 1: dot2_n<1>(c0, c1, sc, v, sv, r0, r1); break;
    case 2: dot2_n<2>(c0, c1, sc, v, sv, r0, r1); break;
    case 3: dot2_n<3>(c0, c1, sc, v, sv, r0, r1); break;
    case 4: dot2_n<4>(c0, c1, sc, v, sv, r0, r1); break;
    case 5: dot2_n<5>(c0, c1, sc, v, sv, r0, r1); break;
    case 6: dot2_n<6>(c0, c1, sc, v, sv, r0, r1); break;
    case 7: dot2_n<7>(c0, c1, sc, v, sv, r0, r1); break;
    case 8: dot2_n<8>(c0, c1, sc, v, sv, r0, r1); break;
    default: {
      double a = 0, b = 0;
      for (int k = 0; k < n; k++) { const double x = v[k * sv]; a = fma(c0[k * sc], x, a); b = fma(c1[k * sc], x, b); }
      r0 = a; r1 = b;
    }
  }
}
// one coefficient row applied to two vectors
template <int NN>
__device__ __forceinline__ void dotp_n(const double* __restrict__ c, const double* __restrict__ v0, const double* __restrict__ v1, int sv,
                                       double& r0, double& r1) {
  double a = 0, b = 0;
#pragma unroll
  for (int k = 0; k < NN; k++) { const double pv = c[k]; a = fma(pv, v0[k * sv], a); b = fma(pv, v1[k * sv], b); }
  r0 = a; r1 = b;
}
__device__ __forceinline__ void dotp(int n, const double* __restrict__ c, const double* __restrict__ v0, const double* __restrict__ v1,
                                     int sv, double& r0, double& r1) {
  switch (n) {
    case 1: dotp_n<1>(c, v0, v1, sv, r0, r1); break;
    case 2: dotp_n<2>(c, v0, v1, sv, r0, r1); break;
    case 3: dotp_n<3>(c, v0, v1, sv, r0, r1); break;
    case 4: dotp_n<4>(c, v0, v1, sv, r0, r1); break;
    case 5: dotp_n<5>(c, v0, v1, sv, r0, r1); break;
    case 6: dotp_n<6>(c, v0, v1, sv, r0, r1); break;
    case 7: dotp_n<7>(c, v0, v1, sv, r0, r1); break;
    case 8: dotp_n<8>(c, v0, v1, sv, r0, r1); break;
    default: {
      double a = 0, b = 0;
      for (int k = 0; k < n; k++) { const double pv = c[k]; a = fma(pv, v0[k * sv], a); b = fma(pv, v1[k * sv], b); }
      r0 = a; r1 = b;
    }
  }
}

template <int B, int E> struct CPow { static constexpr int v = B * CPow<B, E - 1>::v; };
template <int B> struct CPow<B, 0> { static constexpr int v = 1; };

// 1-D tables of the bucket's degree, passed by value (constant bank)
template <int N1> struct GenTab { double MinvS[N1 * N1], M[N1 * N1], mt[2][N1], mg[2][N1], g[2][N1], t[2][N1]; };

// One CTA handles `epc` elements of one degree bucket.  DIM and N1 = p + 1 are compile-time, so all index arithmetic on the
// elements' own blocks folds to constants; only the neighbours' degrees are run-time.
//   face phases : per face node, own trace and neighbour trace (tangentially L2-projected when the degrees differ) -> alpha, beta
//   line passes : one thread per DoF line (N1^(DIM-1) threads per element), the line in registers:
//                 X: w = T~_x u   Y: w += T~_y u   Z: w += T~_z u, w = M_z w   then M_y, then M_x -> global
template <int DIM, int N1>
__global__ void k_apply_generic(GenericParams P, GenTab<N1> T, int maxno1, long cnt, int epc, int per_elem) {
  extern __shared__ double sm_all[];
  constexpr int dim = DIM, nfaces = 2 * DIM;
  constexpr int pe = N1 - 1, n1 = N1;
  constexpr int ne = CPow<N1, DIM>::v;
  constexpr int nf = CPow<N1, DIM - 1>::v;   // face nodes == DoF lines per direction
  const int maxnfo = ipow_d(maxno1, dim - 1);
  const int maxtmp = n1 * maxno1;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long first = (long)blockIdx.x * epc;
  const int nel = (int)min((long)epc, cnt - first);   // elements of this CTA
  // per-element shared memory: su, sw, alpha, beta, rawD, rawV, tmpA, tmpB, face info
  auto SU = [&](int el) { return sm_all + (size_t)el * per_elem; };
  auto SW = [&](int el) { return SU(el) + ne; };
  auto AL = [&](int el) { return SW(el) + ne; };
  auto BE = [&](int el) { return AL(el) + nfaces * nf; };
  auto RD = [&](int el) { return BE(el) + nfaces * nf; };
  auto RV = [&](int el) { return RD(el) + nfaces * maxnfo; };
  auto TA = [&](int el) { return RV(el) + nfaces * maxnfo; };
  auto TB = [&](int el) { return TA(el) + nfaces * maxtmp; };
  auto FI = [&](int el) { return reinterpret_cast<FaceInfo*>(TB(el) + nfaces * maxtmp); };

  // ---- phase 0: load the blocks, face metadata ------------------------------------------------------------------
  for (int t = tid; t < nel * ne; t += nthr) {
    const int el = t / ne, i = t % ne;
    const long e = P.elist[P.ebegin + first + el];
    SU(el)[i] = P.x[P.off[e] + i];
  }
  for (int t = tid; t < nel * nfaces; t += nthr) {
    const int el = t / nfaces, f = t % nfaces, d = f / 2, s = f % 2;
    const long e = P.elist[P.ebegin + first + el];
    long r = e; int ijk[3];
    ijk[0] = (int)(r % P.n[0]); r /= P.n[0]; ijk[1] = (int)(r % P.n[1]); r /= P.n[1]; ijk[2] = (int)r;
    const int c = ijk[d] + (s ? 1 : -1);
    FaceInfo F;
    F.has_nb = (c >= 0 && c < P.n[d]);
    F.skip = (!F.has_nb && !P.dirichlet);
    double kappa = 1.0 / P.h[d];
    for (int dd = 0; dd < dim; dd++) if (dd != d) kappa *= P.h[dd];
    const double nu = s ? 1.0 : -1.0;
    F.po = pe; F.uo = 0;
    double w = 1.0;
    if (F.has_nb) {
      const long stride = d == 0 ? 1 : d == 1 ? P.n[0] : (long)P.n[0] * P.n[1];
      const long o = e + (s ? stride : -stride);
      F.po = P.deg[o]; F.uo = P.off[o];
      const int pm = max(P.pdeg[e], P.pdeg[o]);
      F.cpen = P.sigma * (double)pm * pm; w = 0.5;
    } else F.cpen = P.sigma * (double)P.pdeg[e] * P.pdeg[e];
    F.w_nu_k = w * nu * kappa;
    F.A1 = -0.5 * nu * kappa; F.A2 = -F.cpen; F.A3 = 0.5 * nu * kappa;
    FI(el)[f] = F;
  }
  __syncthreads();

  // ---- phase 1: own traces -> alpha/beta, raw neighbour traces; all elements and faces in parallel -----------------
  for (int t = tid; t < nel * nfaces * nf; t += nthr) {
    const int el = t / (nfaces * nf), rem = t % (nfaces * nf);
    const int f = rem / nf, node = rem % nf, d = f / 2, s = f % 2;
    const FaceInfo& F = FI(el)[f];
    const double* su = SU(el);
    double a = 0, b = 0;
    if (!F.skip) {
      const int base = line_base(dim, n1, d, node), sd = d == 0 ? 1 : d == 1 ? n1 : n1 * n1;
      double der = 0, val = 0;
#pragma unroll
      for (int k = 0; k < n1; k++) { const double v = su[base + k * sd]; der += T.g[s][k] * v; val += T.t[s][k] * v; }
      a = -F.w_nu_k * der + F.cpen * val;
      b = -F.w_nu_k * val;
    }
    AL(el)[f * nf + node] = a; BE(el)[f * nf + node] = b;
  }
  for (int t = tid; t < nel * nfaces * maxnfo; t += nthr) {
    const int el = t / (nfaces * maxnfo), rem = t % (nfaces * maxnfo);
    const int f = rem / maxnfo, node = rem % maxnfo, d = f / 2, s = f % 2;
    const FaceInfo& F = FI(el)[f];
    if (!F.has_nb) continue;
    const int no1 = F.po + 1, nfo = ipow_d(no1, dim - 1);
    if (node >= nfo) continue;
    const DegTable& To = P.tab[F.po];
    const double* uo = P.x + F.uo;
    const int base = line_base(dim, no1, d, node), sd = ipow_d(no1, d);
    double der, val;
    dot2(no1, To.g[1 - s], To.t[1 - s], 1, uo + base, sd, der, val);
    RD(el)[f * maxnfo + node] = der; RV(el)[f * maxnfo + node] = val;
  }
  __syncthreads();
  // ---- phase 2: same-degree faces add directly; mixed-degree faces: tangential L2 projection ---------------------
  if constexpr (DIM == 2) {
    for (int t = tid; t < nel * nfaces * nf; t += nthr) {
      const int el = t / (nfaces * nf), rem = t % (nfaces * nf), f = rem / nf, i = rem % nf;
      const FaceInfo& F = FI(el)[f];
      if (!F.has_nb) continue;
      const double* rawD = RD(el) + f * maxnfo; const double* rawV = RV(el) + f * maxnfo;
      double a, b;
      if (F.po == pe) { a = rawD[i]; b = rawV[i]; }
      else {
        const double* Pm = P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
        a = 0; b = 0;
        for (int k = 0; k <= F.po; k++) { const double pv = Pm[i * kMaxN + k]; a += pv * rawD[k]; b += pv * rawV[k]; }
      }
      AL(el)[f * nf + i] += F.A1 * a + F.A2 * b;
      BE(el)[f * nf + i] += F.A3 * b;
    }
  } else {
    const int slots = max(nf, maxtmp);
    for (int t = tid; t < nel * nfaces * slots; t += nthr) {
      const int el = t / (nfaces * slots), rem = t % (nfaces * slots), f = rem / slots, q = rem % slots;
      const FaceInfo& F = FI(el)[f];
      if (!F.has_nb) continue;
      const double* rawD = RD(el) + f * maxnfo; const double* rawV = RV(el) + f * maxnfo;
      if (F.po == pe) {
        if (q < nf) {
          AL(el)[f * nf + q] += F.A1 * rawD[q] + F.A2 * rawV[q];
          BE(el)[f * nf + q] += F.A3 * rawV[q];
        }
      } else {
        const int no1 = F.po + 1;
        if (q < n1 * no1) {  // tmp[i + n1*b] = sum_a P[i,a] raw[a + no1*b]
          const double* Pm = P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
          const int i = q % n1, b = q / n1;
          double a0, a1;
          dotp(no1, Pm + i * kMaxN, rawD + no1 * b, rawV + no1 * b, 1, a0, a1);
          TA(el)[f * maxtmp + q] = a0; TB(el)[f * maxtmp + q] = a1;
        }
      }
    }
    __syncthreads();
    for (int t = tid; t < nel * nfaces * nf; t += nthr) {
      const int el = t / (nfaces * nf), rem = t % (nfaces * nf), f = rem / nf, q = rem % nf;
      const FaceInfo& F = FI(el)[f];
      if (!F.has_nb || F.po == pe) continue;
      const int no1 = F.po + 1;
      const double* Pm = P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
      const int i = q % n1, j = q / n1;
      double a0, a1;
      dotp(no1, Pm + j * kMaxN, TA(el) + f * maxtmp + i, TB(el) + f * maxtmp + i, n1, a0, a1);
      AL(el)[f * nf + q] += F.A1 * a0 + F.A2 * a1;
      BE(el)[f * nf + q] += F.A3 * a1;
    }
  }
  __syncthreads();

  // ---- line passes ------------------------------------------------------------------------------------------------
  double kap[3] = {1, 1, 1};
#pragma unroll
  for (int d = 0; d < dim; d++) {
    double k = 1.0 / P.h[d];
    for (int dd = 0; dd < dim; dd++) if (dd != d) k *= P.h[dd];
    kap[d] = k;
  }
  const int lel = tid / nf, line = tid % nf;       // (element in CTA, line)
  const bool lact = lel < nel;
  // base offset and stride of line `line` along direction d (x-fastest local index)
  auto lbase = [&](int d) { return d == 0 ? n1 * line : (DIM == 2 ? line : (d == 1 ? (line % n1) + n1 * n1 * (line / n1) : line)); };
  auto lstride = [&](int d) { return d == 0 ? 1 : (d == 1 ? n1 : n1 * n1); };
#pragma unroll
  for (int d = 0; d < dim; d++) {
    if (lact) {
      const double* su = SU(lel); double* sw = SW(lel);
      const int base = lbase(d), sd = lstride(d);
      double v[N1], w[N1];
#pragma unroll
      for (int k = 0; k < n1; k++) v[k] = su[base + k * sd];
      const double a0 = AL(lel)[(2 * d) * nf + line], b0 = BE(lel)[(2 * d) * nf + line];
      const double a1 = AL(lel)[(2 * d + 1) * nf + line], b1 = BE(lel)[(2 * d + 1) * nf + line];
#pragma unroll
      for (int i = 0; i < n1; i++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < n1; k++) s = fma(T.MinvS[i * n1 + k], v[k], s);
        s *= kap[d];
        s = fma(T.mt[0][i], a0, s); s = fma(T.mg[0][i], b0, s);
        s = fma(T.mt[1][i], a1, s); s = fma(T.mg[1][i], b1, s);
        w[i] = (d == 0) ? s : s + sw[base + i * sd];
      }
      if (d == dim - 1) {  // last direction: its mass sweep acts on the same line
        double o[N1];
#pragma unroll
        for (int i = 0; i < n1; i++) {
          double s = 0;
#pragma unroll
          for (int k = 0; k < n1; k++) s = fma(T.M[i * n1 + k], w[k], s);
          o[i] = s;
        }
#pragma unroll
        for (int i = 0; i < n1; i++) w[i] = o[i];
      }
      if (dim == 1) { /* unreachable */ }
#pragma unroll
      for (int i = 0; i < n1; i++) sw[base + i * sd] = w[i];
    }
    __syncthreads();
  }
  // remaining mass sweeps, directions dim-2 .. 0; the last one (x lines, contiguous) writes to global
#pragma unroll
  for (int d = dim - 2; d >= 0; d--) {
    if (lact) {
      double* sw = SW(lel);
      const int base = lbase(d), sd = lstride(d);
      double w[N1];
#pragma unroll
      for (int k = 0; k < n1; k++) w[k] = sw[base + k * sd];
      const long e = P.elist[P.ebegin + first + lel];
      double* yo = P.y + P.off[e];
#pragma unroll
      for (int i = 0; i < n1; i++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < n1; k++) s = fma(T.M[i * n1 + k], w[k], s);
        if (d == 0) yo[base + i] = P.accum ? yo[base + i] + P.factor * s : P.factor * s;
        else sw[base + i * sd] = s;
      }
    }
    if (d > 0) __syncthreads();
  }
}

int launch_apply_generic(Ctx* ctx, Level& L, const double* x, double* y, double factor) {
  GenericParams P;
  P.dim = L.dim;
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.h[d] = L.h[d]; }
  P.sigma = ctx->sigma; P.dirichlet = ctx->dirichlet;
  P.deg = L.d_deg; P.pdeg = L.d_pdeg; P.off = L.d_off; P.elist = L.d_elist;
  P.tab = ctx->d_tab; P.P = ctx->d_P; P.x = x; P.y = y; P.factor = factor; P.accum = ctx->fuse_accum;
  // the degree buckets write disjoint rows of y: launch them on side streams so that small buckets overlap the large ones
  const size_t nb = L.bucket_p.size();
  const bool fork = nb > 1;
  if (fork) {
    if (!ctx->bucket_stream[0]) {
      for (int k = 0; k < 4; k++) HPDG_CUDA(cudaStreamCreateWithFlags(&ctx->bucket_stream[k], cudaStreamNonBlocking));
      for (int k = 0; k < 5; k++) HPDG_CUDA(cudaEventCreateWithFlags(&ctx->bucket_ev[k], cudaEventDisableTiming));
    }
    HPDG_CUDA(cudaEventRecord(ctx->bucket_ev[4], ctx->stream));
    for (int k = 0; k < 4; k++) HPDG_CUDA(cudaStreamWaitEvent(ctx->bucket_stream[k], ctx->bucket_ev[4], 0));
  }
  for (size_t b = 0; b < nb; b++) {
    long cnt = L.bucket_begin[b + 1] - L.bucket_begin[b];
    if (cnt == 0) continue;
    cudaStream_t lstream = fork ? ctx->bucket_stream[b % 4] : ctx->stream;
    int p = L.bucket_p[b], n1 = p + 1;
    int ne = 1, nf = 1;
    for (int d = 0; d < L.dim; d++) ne *= n1;
    for (int d = 0; d < L.dim - 1; d++) nf *= n1;
    const int maxno1 = L.maxp + 1;
    int maxnfo = 1;
    for (int d = 0; d < L.dim - 1; d++) maxnfo *= maxno1;
    const int nfaces = 2 * L.dim;
    const int per_elem = 2 * ne + 2 * nfaces * nf + 2 * nfaces * maxnfo + 2 * nfaces * n1 * maxno1 +
                         (int)((nfaces * sizeof(FaceInfo) + 7) / 8);
    P.ebegin = L.bucket_begin[b];
    // elements per CTA: aim at ~128 line threads, bounded by 96 KB of shared memory
    int epc = nf >= 64 ? 1 : (L.uniform ? 128 / nf : std::max(1, 64 / nf));  // mixed-degree meshes: face work dominates, fewer elements per CTA
    epc = (int)std::max<long>(1, std::min<long>(epc, (96 * 1024) / ((long)per_elem * 8)));
    epc = (int)std::min<long>(epc, cnt);
    const size_t smem_l = (size_t)epc * per_elem * sizeof(double);
    const unsigned grid = (unsigned)((cnt + epc - 1) / epc);
    const int threads = std::max(L.uniform ? 128 : 256, (epc * nf + 31) / 32 * 32);  // line passes use epc*nf threads, the face phases all of them
    const DegTable& HT = host_tables().deg[p];
#define HPDG_GEN_LAUNCH(D, NN)                                                                                            \
  do {                                                                                                                    \
    GenTab<NN> T;                                                                                                         \
    for (int i = 0; i < NN; i++) {                                                                                        \
      for (int j = 0; j < NN; j++) { T.MinvS[i * NN + j] = HT.MinvS[i * kMaxN + j]; T.M[i * NN + j] = HT.M[i * kMaxN + j]; } \
      for (int sd = 0; sd < 2; sd++) { T.mt[sd][i] = HT.mt[sd][i]; T.mg[sd][i] = HT.mg[sd][i]; T.g[sd][i] = HT.g[sd][i]; T.t[sd][i] = HT.t[sd][i]; } \
    }                                                                                                                     \
    if (smem_l + 1024 > 48 * 1024)                                                                                        \
      HPDG_CUDA(cudaFuncSetAttribute(k_apply_generic<D, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));  \
    k_apply_generic<D, NN><<<grid, threads, smem_l, lstream>>>(P, T, maxno1, cnt, epc, per_elem);                     \
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
    for (int k = 0; k < 4; k++) {
      HPDG_CUDA(cudaEventRecord(ctx->bucket_ev[k], ctx->bucket_stream[k]));
      HPDG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->bucket_ev[k], 0));
    }
  }
  return 0;
}

}  // namespace hpdg
