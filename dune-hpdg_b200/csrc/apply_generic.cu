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

// WARP: one warp per element (small blocks, several elements per CTA, __syncwarp instead of block barriers);
// otherwise one CTA per element.
template <int B, int E> struct CPow { static constexpr int v = B * CPow<B, E - 1>::v; };
template <int B> struct CPow<B, 0> { static constexpr int v = 1; };

// DIM and N1 = p_e + 1 of the bucket are compile-time so that every index computation on the element's own block
// folds to constants; only the neighbour degree stays a run-time quantity.
template <int DIM, int N1, bool WARP>
__global__ void k_apply_generic(GenericParams P, int maxno1, long cnt, int smem_per_group) {
  extern __shared__ double sm_all[];
  __shared__ FaceInfo finfo_all[8][6];
  const int grp = WARP ? threadIdx.x / 32 : 0;
  const int ltid = WARP ? threadIdx.x % 32 : threadIdx.x;
  const int gsize = WARP ? 32 : blockDim.x;
  const long slot = WARP ? (long)blockIdx.x * (blockDim.x / 32) + grp : blockIdx.x;
  if (slot >= cnt) return;   // whole group leaves together
  auto gsync = [&]() { if (WARP) __syncwarp(); else __syncthreads(); };
  double* sm = sm_all + (size_t)grp * smem_per_group;
  FaceInfo* finfo = finfo_all[grp];
  const long e = P.elist[P.ebegin + slot];
  constexpr int dim = DIM, nfaces = 2 * DIM;
  constexpr int pe = N1 - 1, n1 = N1;
  constexpr int ne = CPow<N1, DIM>::v;
  constexpr int nf = CPow<N1, DIM - 1>::v;
  const int maxnfo = ipow_d(maxno1, dim - 1);      // raw neighbour trace slots per face
  const int maxtmp = n1 * maxno1;                  // stage-1 projection slots per face (3-D)
  double* su = sm;
  double* sw = su + ne;
  double* st = sw + ne;
  double* alpha = st + ne;                 // [nfaces][nf]
  double* beta = alpha + nfaces * nf;
  double* rawD = beta + nfaces * nf;       // [nfaces][maxnfo]
  double* rawV = rawD + nfaces * maxnfo;
  double* tmpA = rawV + nfaces * maxnfo;   // [nfaces][maxtmp]
  double* tmpB = tmpA + nfaces * maxtmp;
  const DegTable& T = P.tab[pe];
  const double* ue = P.x + P.off[e];
  for (int i = ltid; i < ne; i += gsize) su[i] = ue[i];
  if (ltid < nfaces) {
    const int f = ltid, d = f / 2, s = f % 2;
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
    finfo[f] = F;
  }
  gsync();

  // ---- phase 1: own traces -> alpha/beta, raw neighbour traces, all faces in parallel -----------------
  {
    const int slots = max(nf, maxnfo);
    for (int t = ltid; t < nfaces * slots; t += gsize) {
      const int f = t / slots, node = t % slots, d = f / 2, s = f % 2;
      const FaceInfo& F = finfo[f];
      if (node < nf) {
        double a = 0, b = 0;
        if (!F.skip) {
          const int base = line_base(dim, n1, d, node), sd = d == 0 ? 1 : d == 1 ? n1 : n1 * n1;
          double der = 0, val = 0;
#pragma unroll
          for (int k = 0; k < n1; k++) { const double v = su[base + k * sd]; der += T.g[s][k] * v; val += T.t[s][k] * v; }
          a = -F.w_nu_k * der + F.cpen * val;
          b = -F.w_nu_k * val;
        }
        alpha[f * nf + node] = a; beta[f * nf + node] = b;
      }
      if (F.has_nb) {
        const int no1 = F.po + 1, nfo = ipow_d(no1, dim - 1);
        if (node < nfo) {
          const DegTable& To = P.tab[F.po];
          const double* uo = P.x + F.uo;
          const int base = line_base(dim, no1, d, node), sd = ipow_d(no1, d);
          double der = 0, val = 0;
          for (int k = 0; k < no1; k++) { const double v = __ldg(uo + base + k * sd); der += To.g[1 - s][k] * v; val += To.t[1 - s][k] * v; }
          rawD[f * maxnfo + node] = der; rawV[f * maxnfo + node] = val;
        }
      }
    }
  }
  gsync();
  // ---- phase 2: same-degree faces add directly; mixed-degree faces: tangential L2 projection, first direction ----
  if constexpr (DIM == 2) {
    for (int t = ltid; t < nfaces * nf; t += gsize) {
      const int f = t / nf, i = t % nf;
      const FaceInfo& F = finfo[f];
      if (!F.has_nb) continue;
      double a, b;
      if (F.po == pe) { a = rawD[f * maxnfo + i]; b = rawV[f * maxnfo + i]; }
      else {
        const double* Pm = P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
        a = 0; b = 0;
        for (int k = 0; k <= F.po; k++) { const double pv = Pm[i * kMaxN + k]; a += pv * rawD[f * maxnfo + k]; b += pv * rawV[f * maxnfo + k]; }
      }
      alpha[f * nf + i] += F.A1 * a + F.A2 * b;
      beta[f * nf + i] += F.A3 * b;
    }
  } else {
    const int slots = max(nf, maxtmp);
    for (int t = ltid; t < nfaces * slots; t += gsize) {
      const int f = t / slots, q = t % slots;
      const FaceInfo& F = finfo[f];
      if (!F.has_nb) continue;
      if (F.po == pe) {
        if (q < nf) {
          alpha[f * nf + q] += F.A1 * rawD[f * maxnfo + q] + F.A2 * rawV[f * maxnfo + q];
          beta[f * nf + q] += F.A3 * rawV[f * maxnfo + q];
        }
      } else {
        const int no1 = F.po + 1;
        if (q < n1 * no1) {  // tmp[i + n1*b] = sum_a P[i,a] raw[a + no1*b]
          const double* Pm = P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
          const int i = q % n1, b = q / n1;
          double a0 = 0, a1 = 0;
          for (int k = 0; k < no1; k++) { const double pv = Pm[i * kMaxN + k]; a0 += pv * rawD[f * maxnfo + k + no1 * b]; a1 += pv * rawV[f * maxnfo + k + no1 * b]; }
          tmpA[f * maxtmp + q] = a0; tmpB[f * maxtmp + q] = a1;
        }
      }
    }
    gsync();
    for (int t = ltid; t < nfaces * nf; t += gsize) {
      const int f = t / nf, q = t % nf;
      const FaceInfo& F = finfo[f];
      if (!F.has_nb || F.po == pe) continue;
      const int no1 = F.po + 1;
      const double* Pm = P.P + ((size_t)pe * (kMaxP + 1) + F.po) * kMaxN * kMaxN;
      const int i = q % n1, j = q / n1;
      double a0 = 0, a1 = 0;
      for (int k = 0; k < no1; k++) { const double pv = Pm[j * kMaxN + k]; a0 += pv * tmpA[f * maxtmp + i + n1 * k]; a1 += pv * tmpB[f * maxtmp + i + n1 * k]; }
      alpha[f * nf + q] += F.A1 * a0 + F.A2 * a1;
      beta[f * nf + q] += F.A3 * a1;
    }
  }
  gsync();
  double kap[3];
  for (int d = 0; d < dim; d++) {
    double k = 1.0 / P.h[d];
    for (int dd = 0; dd < dim; dd++) if (dd != d) k *= P.h[dd];
    kap[d] = k;
  }
  // w = sum_d [ kappa_d MinvS u + trace terms ]
  for (int idx = ltid; idx < ne; idx += gsize) {
    int a[3] = {0, 0, 0}, rem = idx;
#pragma unroll
    for (int d = 0; d < dim; d++) { a[d] = rem % n1; rem /= n1; }
    double acc = 0;
    int sd = 1;
#pragma unroll
    for (int d = 0; d < dim; d++) {
      const int base = idx - a[d] * sd;
      double s = 0;
#pragma unroll
      for (int k = 0; k < n1; k++) s += T.MinvS[a[d] * kMaxN + k] * su[base + k * sd];
      // tangential index of this dof on faces normal to d
      int ti = 0, ts = 1;
      for (int dd = 0; dd < dim; dd++) if (dd != d) { ti += a[dd] * ts; ts *= n1; }
      const double* al0 = alpha + (2 * d) * nf; const double* be0 = beta + (2 * d) * nf;
      const double* al1 = al0 + nf; const double* be1 = be0 + nf;
      acc += kap[d] * s + T.mt[0][a[d]] * al0[ti] + T.mg[0][a[d]] * be0[ti] + T.mt[1][a[d]] * al1[ti] + T.mg[1][a[d]] * be1[ti];
      sd *= n1;
    }
    sw[idx] = acc;
  }
  gsync();
  // y = factor * M_z M_y M_x w
  double* src = sw; double* dst = st;
  int sd = 1;
#pragma unroll
  for (int d = 0; d < dim; d++) {
    const bool last = (d == dim - 1);
    for (int idx = ltid; idx < ne; idx += gsize) {
      int ad = (idx / sd) % n1;
      const int base = idx - ad * sd;
      double s = 0;
#pragma unroll
      for (int k = 0; k < n1; k++) s += T.M[ad * kMaxN + k] * src[base + k * sd];
      if (last) P.y[P.off[e] + idx] = P.accum ? P.y[P.off[e] + idx] + P.factor * s : P.factor * s; else dst[idx] = s;
    }
    gsync();
    double* t = src; src = dst; dst = t;
    sd *= n1;
  }
}

int launch_apply_generic(Ctx* ctx, Level& L, const double* x, double* y, double factor) {
  GenericParams P;
  P.dim = L.dim;
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.h[d] = L.h[d]; }
  P.sigma = ctx->sigma; P.dirichlet = ctx->dirichlet;
  P.deg = L.d_deg; P.pdeg = L.d_pdeg; P.off = L.d_off; P.elist = L.d_elist;
  P.tab = ctx->d_tab; P.P = ctx->d_P; P.x = x; P.y = y; P.factor = factor; P.accum = ctx->fuse_accum;
  for (size_t b = 0; b < L.bucket_p.size(); b++) {
    long cnt = L.bucket_begin[b + 1] - L.bucket_begin[b];
    if (cnt == 0) continue;
    int p = L.bucket_p[b], n1 = p + 1;
    int ne = 1, nf = 1;
    for (int d = 0; d < L.dim; d++) ne *= n1;
    for (int d = 0; d < L.dim - 1; d++) nf *= n1;
    const int maxno1 = L.maxp + 1;
    int maxnfo = 1;
    for (int d = 0; d < L.dim - 1; d++) maxnfo *= maxno1;
    const int nfaces = 2 * L.dim;
    size_t smem = sizeof(double) * (3 * (size_t)ne + 2 * (size_t)nfaces * nf + 2 * (size_t)nfaces * maxnfo +
                                    2 * (size_t)nfaces * n1 * maxno1);
    P.ebegin = L.bucket_begin[b];
    const int per_group = (int)(smem / sizeof(double));
    const bool warp_mode = (ne <= 64 && smem * 8 <= 96 * 1024);
    const size_t smem_l = warp_mode ? smem * 8 : smem;
    const unsigned grid = warp_mode ? (unsigned)((cnt + 7) / 8) : (unsigned)cnt;
    const int threads = warp_mode ? 256 : (ne <= 128 ? 128 : 256);
#define HPDG_GEN_LAUNCH(D, NN, W)                                                                                         \
  do {                                                                                                                    \
    if (smem_l + 4096 > 48 * 1024)                                                                                        \
      HPDG_CUDA(cudaFuncSetAttribute(k_apply_generic<D, NN, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l)); \
    k_apply_generic<D, NN, W><<<grid, threads, smem_l, ctx->stream>>>(P, maxno1, cnt, per_group);                         \
  } while (0)
#define HPDG_GEN_CASE(NN)                                                                   \
  case NN:                                                                                  \
    if (L.dim == 2) { if (warp_mode) HPDG_GEN_LAUNCH(2, NN, true); else HPDG_GEN_LAUNCH(2, NN, false); } \
    else { if (warp_mode) HPDG_GEN_LAUNCH(3, NN, true); else HPDG_GEN_LAUNCH(3, NN, false); }            \
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
  return 0;
}

}  // namespace hpdg
