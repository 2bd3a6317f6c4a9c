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
};

__device__ __forceinline__ int ipow_d(int b, int e) { int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }

// Compute, for face (d,s) of element e, alpha/beta at the element's n^(dim-1) face nodes.
//   su: element coefficients in shared memory; rawD/rawV/tmpA/tmpB: scratch of size >= kMaxN^2
__device__ void face_traces(const GenericParams& P, long e, int pe, int d, int s, const double* su,
                            double* alpha, double* beta, double* rawD, double* rawV, double* tmpA,
                            double* tmpB) {
  const int dim = P.dim, ne1 = pe + 1;
  const int nf = ipow_d(ne1, dim - 1);
  const DegTable& Te = P.tab[pe];
  // element coordinates
  long r = e;
  int ijk[3];
  ijk[0] = (int)(r % P.n[0]); r /= P.n[0];
  ijk[1] = (int)(r % P.n[1]); r /= P.n[1];
  ijk[2] = (int)r;
  const int c = ijk[d] + (s ? 1 : -1);
  const bool has_nb = (c >= 0 && c < P.n[d]);
  double kappa = 1.0 / P.h[d];
  for (int dd = 0; dd < dim; dd++) if (dd != d) kappa *= P.h[dd];
  const double nu = s ? 1.0 : -1.0;
  if (!has_nb && !P.dirichlet) {  // natural boundary: no face term (ipdgoperator.hh:97-105)
    for (int i = threadIdx.x; i < nf; i += blockDim.x) { alpha[i] = 0; beta[i] = 0; }
    return;
  }
  double w, cpen;
  long o = -1;
  int po = pe;
  if (has_nb) {
    long stride = d == 0 ? 1 : d == 1 ? P.n[0] : (long)P.n[0] * P.n[1];
    o = e + (s ? stride : -stride);
    po = P.deg[o];
    int pm = max(P.pdeg[e], P.pdeg[o]);
    cpen = P.sigma * (double)pm * pm;  // ipdgoperator.hh:129-131
    w = 0.5;
  } else {
    cpen = P.sigma * (double)P.pdeg[e] * P.pdeg[e];  // ipdgoperator.hh:310
    w = 1.0;                                           // :357 "no 0.5 here"
  }
  // strides of the tangential directions inside a block of size n1^dim
  // own part
  {
    int sd = ipow_d(ne1, d);
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
      // tangential multi-index -> base offset
      int rem = i, base = 0, st = 1;
      for (int dd = 0; dd < dim; dd++) {
        if (dd != d) { base += (rem % ne1) * st; rem /= ne1; }
        st *= ne1;
      }
      double der = 0, val = 0;
      for (int k = 0; k < ne1; k++) { double v = su[base + k * sd]; der += Te.g[s][k] * v; val += Te.t[s][k] * v; }
      alpha[i] = -w * nu * kappa * der + cpen * val;
      beta[i] = -w * nu * kappa * val;
    }
  }
  if (!has_nb) return;
  // neighbour part: raw traces at the neighbour's face nodes (its side 1-s)
  const int no1 = po + 1;
  const int nfo = ipow_d(no1, dim - 1);
  const DegTable& To = P.tab[po];
  const double* uo = P.x + P.off[o];
  {
    int sd = ipow_d(no1, d);
    for (int i = threadIdx.x; i < nfo; i += blockDim.x) {
      int rem = i, base = 0, st = 1;
      for (int dd = 0; dd < dim; dd++) {
        if (dd != d) { base += (rem % no1) * st; rem /= no1; }
        st *= no1;
      }
      double der = 0, val = 0;
      for (int k = 0; k < no1; k++) { double v = __ldg(uo + base + k * sd); der += To.g[1 - s][k] * v; val += To.t[1 - s][k] * v; }
      rawD[i] = der; rawV[i] = val;
    }
  }
  __syncthreads();
  const double A1 = -0.5 * nu * kappa, A2 = -cpen, A3 = 0.5 * nu * kappa;
  if (po == pe) {
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
      alpha[i] += A1 * rawD[i] + A2 * rawV[i];
      beta[i] += A3 * rawV[i];
    }
    __syncthreads();
    return;
  }
  // project (no1)^(dim-1) -> (ne1)^(dim-1) with Pm = (M^{ee})^-1 M^{eo}, one tangential direction at a time
  const double* Pm = P.P + ((size_t)pe * (kMaxP + 1) + po) * kMaxN * kMaxN;
  if (dim == 2) {
    for (int i = threadIdx.x; i < ne1; i += blockDim.x) {
      double a = 0, b = 0;
      for (int k = 0; k < no1; k++) { double pv = Pm[i * kMaxN + k]; a += pv * rawD[k]; b += pv * rawV[k]; }
      alpha[i] += A1 * a + A2 * b;
      beta[i] += A3 * b;
    }
  } else {
    // first tangential direction (fast index): tmp[i + ne1*b] = sum_a P[i,a] raw[a + no1*b]
    for (int t = threadIdx.x; t < ne1 * no1; t += blockDim.x) {
      int i = t % ne1, b = t / ne1;
      double a0 = 0, a1 = 0;
      for (int k = 0; k < no1; k++) { double pv = Pm[i * kMaxN + k]; a0 += pv * rawD[k + no1 * b]; a1 += pv * rawV[k + no1 * b]; }
      tmpA[t] = a0; tmpB[t] = a1;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nf; t += blockDim.x) {
      int i = t % ne1, j = t / ne1;
      double a0 = 0, a1 = 0;
      for (int k = 0; k < no1; k++) { double pv = Pm[j * kMaxN + k]; a0 += pv * tmpA[i + ne1 * k]; a1 += pv * tmpB[i + ne1 * k]; }
      alpha[t] += A1 * a0 + A2 * a1;
      beta[t] += A3 * a1;
    }
  }
  __syncthreads();
}

__global__ void k_apply_generic(GenericParams P) {
  extern __shared__ double sm[];
  const long e = P.elist[P.ebegin + blockIdx.x];
  const int dim = P.dim;
  const int pe = P.deg[e], n1 = pe + 1;
  const int ne = ipow_d(n1, dim);
  const int nf = ipow_d(n1, dim - 1);
  double* su = sm;
  double* sw = su + ne;
  double* st = sw + ne;
  double* alpha = st + ne;            // [2*dim][nf]
  double* beta = alpha + 2 * dim * nf;
  double* rawD = beta + 2 * dim * nf;  // kMaxN^2 each
  double* rawV = rawD + kMaxN * kMaxN;
  double* tmpA = rawV + kMaxN * kMaxN;
  double* tmpB = tmpA + kMaxN * kMaxN;
  const DegTable& T = P.tab[pe];
  const double* ue = P.x + P.off[e];
  for (int i = threadIdx.x; i < ne; i += blockDim.x) su[i] = ue[i];
  __syncthreads();
  for (int f = 0; f < 2 * dim; f++)
    face_traces(P, e, pe, f / 2, f % 2, su, alpha + f * nf, beta + f * nf, rawD, rawV, tmpA, tmpB);
  __syncthreads();
  double kap[3];
  for (int d = 0; d < dim; d++) {
    double k = 1.0 / P.h[d];
    for (int dd = 0; dd < dim; dd++) if (dd != d) k *= P.h[dd];
    kap[d] = k;
  }
  // w = sum_d [ kappa_d MinvS u + trace terms ]
  for (int idx = threadIdx.x; idx < ne; idx += blockDim.x) {
    int a[3] = {0, 0, 0}, rem = idx;
    for (int d = 0; d < dim; d++) { a[d] = rem % n1; rem /= n1; }
    double acc = 0;
    int sd = 1;
    for (int d = 0; d < dim; d++) {
      const int base = idx - a[d] * sd;
      double s = 0;
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
  __syncthreads();
  // y = factor * M_z M_y M_x w
  double* src = sw; double* dst = st;
  int sd = 1;
  for (int d = 0; d < dim; d++) {
    const bool last = (d == dim - 1);
    for (int idx = threadIdx.x; idx < ne; idx += blockDim.x) {
      int ad = (idx / sd) % n1;
      const int base = idx - ad * sd;
      double s = 0;
      for (int k = 0; k < n1; k++) s += T.M[ad * kMaxN + k] * src[base + k * sd];
      if (last) P.y[P.off[e] + idx] = P.factor * s; else dst[idx] = s;
    }
    __syncthreads();
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
  P.tab = ctx->d_tab; P.P = ctx->d_P; P.x = x; P.y = y; P.factor = factor;
  for (size_t b = 0; b < L.bucket_p.size(); b++) {
    long cnt = L.bucket_begin[b + 1] - L.bucket_begin[b];
    if (cnt == 0) continue;
    int p = L.bucket_p[b], n1 = p + 1;
    int ne = 1, nf = 1;
    for (int d = 0; d < L.dim; d++) ne *= n1;
    for (int d = 0; d < L.dim - 1; d++) nf *= n1;
    size_t smem = sizeof(double) * (3 * (size_t)ne + 4 * (size_t)L.dim * nf + 4 * kMaxN * kMaxN);
    int threads = ne <= 32 ? 32 : ne <= 64 ? 64 : ne <= 128 ? 128 : 256;
    if (smem > 48 * 1024) HPDG_CUDA(cudaFuncSetAttribute(k_apply_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P.ebegin = L.bucket_begin[b];
    k_apply_generic<<<(unsigned)cnt, threads, smem, ctx->stream>>>(P);
    ctx->launches++;
    HPDG_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace hpdg
