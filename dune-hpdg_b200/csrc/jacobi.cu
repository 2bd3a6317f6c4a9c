// Block-Jacobi smoother: c = damping * sum_e P_e^T D_e^-1 P_e r.
//
// Replaces the reference's matrix-free block Jacobi, IPDGBlockJacobi driven by Operator::apply
// (matrix-free/localoperators/ipdgblockjacobi.hh:58-178, matrix-free/operator.hh:41-56), which
// re-assembles the diagonal block D_e = A_ee of every element on every sweep and hands it to a
// local solver.  Here D_e^-1 is precomputed once per level, in one of two forms:
//   dense : explicit n_e x n_e inverses, stored bucket by bucket (batched by block size), applied as
//           a batched dense mat-vec (HBM bound: 8 n_e^2 + 16 n_e bytes per element);
//   fd    : fast diagonalisation.  On the axis-parallel mesh D_e = sum_d M x .. x D_d x .. x M is a
//           Kronecker sum, so with the 1-D generalised eigenpairs D_d V_d = M V_d L_d, V_d^T M V_d = I
//           D_e^-1 = (Vx x Vy x Vz) diag(1/(lx+ly+lz)) (Vx x Vy x Vz)^T exactly; 16 B/DoF.
// Both are the exact inverse of the same block the reference assembles (penalty with the max of
// the two degrees, avg factor 1/2 interior and 1 on Dirichlet faces: ipdgblockjacobi.hh:69-86).
#include <cmath>
#include <cstdio>
#include <map>
#include <tuple>

#include "ctx.hpp"

namespace hpdg {

struct BlkParams {
  int dim;
  int n[3];
  double h[3];
  double sigma;
  int dirichlet;
  const int* deg;
  const int* pdeg;
  const long* off;
  const int* elist;
  long ebegin;
  const DegTable* tab;
};

__device__ __forceinline__ int ipw(int b, int e) { int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }

// 1-D factor D_d of element e (n1 x n1, stride n1) into out[]; executed by the whole CTA.
__device__ void build_dir_matrix(const BlkParams& P, long e, int d, double* out) {
  const int pe = P.deg[e], n1 = pe + 1;
  const DegTable& T = P.tab[pe];
  long r = e; int ijk[3];
  ijk[0] = (int)(r % P.n[0]); r /= P.n[0]; ijk[1] = (int)(r % P.n[1]); r /= P.n[1]; ijk[2] = (int)r;
  double kappa = 1.0 / P.h[d];
  for (int dd = 0; dd < P.dim; dd++) if (dd != d) kappa *= P.h[dd];
  double w[2], c[2];
  for (int s = 0; s < 2; s++) {
    const int cc = ijk[d] + (s ? 1 : -1);
    if (cc >= 0 && cc < P.n[d]) {
      long stride = d == 0 ? 1 : d == 1 ? P.n[0] : (long)P.n[0] * P.n[1];
      long o = e + (s ? stride : -stride);
      int pm = max(P.pdeg[e], P.pdeg[o]);
      w[s] = 0.5; c[s] = P.sigma * (double)pm * pm;
    } else if (P.dirichlet) { w[s] = 1.0; c[s] = P.sigma * (double)P.pdeg[e] * P.pdeg[e]; }
    else { w[s] = 0.0; c[s] = 0.0; }
  }
  for (int t = threadIdx.x; t < n1 * n1; t += blockDim.x) {
    const int i = t / n1, j = t % n1;
    double v = kappa * T.S[i * kMaxN + j];
    for (int s = 0; s < 2; s++) {
      const double nu = s ? 1.0 : -1.0;
      v += -w[s] * nu * kappa * (T.t[s][i] * T.g[s][j] + T.g[s][i] * T.t[s][j]) + c[s] * T.t[s][i] * T.t[s][j];
    }
    out[t] = v;
  }
}

// D_e (ne x ne row-major) into A (global or shared), CTA-wide.
__device__ void build_block(const BlkParams& P, long e, double* sD /* 3*kMaxN^2 smem */, double* A) {
  const int pe = P.deg[e], n1 = pe + 1, dim = P.dim;
  const int ne = ipw(n1, dim);
  const DegTable& T = P.tab[pe];
  for (int d = 0; d < dim; d++) build_dir_matrix(P, e, d, sD + d * kMaxN * kMaxN);
  __syncthreads();
  for (long t = threadIdx.x; t < (long)ne * ne; t += blockDim.x) {
    int a = (int)(t / ne), b = (int)(t % ne);
    int ai[3] = {0, 0, 0}, bi[3] = {0, 0, 0};
    for (int d = 0; d < dim; d++) { ai[d] = a % n1; a /= n1; bi[d] = b % n1; b /= n1; }
    double m[3] = {1, 1, 1};
    for (int d = 0; d < dim; d++) m[d] = T.M[ai[d] * kMaxN + bi[d]];
    double v = 0;
    for (int d = 0; d < dim; d++) {
      double f = sD[d * kMaxN * kMaxN + ai[d] * n1 + bi[d]];
      for (int dd = 0; dd < dim; dd++) if (dd != d) f *= m[dd];
      v += f;
    }
    A[t] = v;
  }
  __syncthreads();
}

__global__ void k_build_blocks(BlkParams P, double* out, int ne) {
  __shared__ double sD[3 * kMaxN * kMaxN];
  const long e = P.elist[P.ebegin + blockIdx.x];
  build_block(P, e, sD, out + (size_t)blockIdx.x * ne * ne);
}

// In-place Gauss-Jordan inversion without pivoting (SPD blocks), one CTA per matrix, in global memory.
__global__ void k_invert_blocks(double* A_all, int n) {
  double* A = A_all + (size_t)blockIdx.x * n * n;
  __shared__ double piv;
  extern __shared__ double colk[];  // column k copy (n) + row k copy (n)
  double* rowk = colk + n;
  for (int k = 0; k < n; k++) {
    if (threadIdx.x == 0) piv = 1.0 / A[(size_t)k * n + k];
    for (int i = threadIdx.x; i < n; i += blockDim.x) { colk[i] = A[(size_t)i * n + k]; rowk[i] = A[(size_t)k * n + i]; }
    __syncthreads();
    const double p = piv;
    for (long t = threadIdx.x; t < (long)n * n; t += blockDim.x) {
      const int i = (int)(t / n), j = (int)(t % n);
      double v;
      if (i == k) v = (j == k) ? p : rowk[j] * p;
      else if (j == k) v = -colk[i] * p;
      else v = A[t] - colk[i] * rowk[j] * p;
      A[t] = v;
    }
    __syncthreads();
  }
}

// c_e = damping * Dinv_e r_e; thread (slot, i) accumulates output row i using column access of the
// symmetric inverse (coalesced over i).
__global__ void k_jacobi_dense(const double* __restrict__ inv, const int* __restrict__ elist, long ebegin, long cnt,
                               const long* __restrict__ off, int ne, int eper, const double* __restrict__ r,
                               double* __restrict__ c, double damping, double* __restrict__ xacc) {
  extern __shared__ double sr[];  // eper * ne
  const int slot = threadIdx.x / ne, i = threadIdx.x % ne;
  const long b = (long)blockIdx.x * eper + slot;
  const bool act = slot < eper && b < cnt;
  long e = 0;
  if (act) { e = elist[ebegin + b]; sr[slot * ne + i] = r[off[e] + i]; }
  __syncthreads();
  if (!act) return;
  const double* A = inv + (size_t)b * ne * ne;
  const double* rr = sr + slot * ne;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int j = 0;
  for (; j + 3 < ne; j += 4) {
    a0 = fma(__ldcs(A + (size_t)j * ne + i), rr[j], a0);
    a1 = fma(__ldcs(A + (size_t)(j + 1) * ne + i), rr[j + 1], a1);
    a2 = fma(__ldcs(A + (size_t)(j + 2) * ne + i), rr[j + 2], a2);
    a3 = fma(__ldcs(A + (size_t)(j + 3) * ne + i), rr[j + 3], a3);
  }
  for (; j < ne; j++) a0 = fma(__ldcs(A + (size_t)j * ne + i), rr[j], a0);
  const double cv = damping * ((a0 + a1) + (a2 + a3));
  c[off[e] + i] = cv;
  if (xacc) xacc[off[e] + i] += cv;
}

static BlkParams make_blk(Ctx* ctx, Level& L) {
  BlkParams P;
  P.dim = L.dim;
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.h[d] = L.h[d]; }
  P.sigma = ctx->sigma; P.dirichlet = ctx->dirichlet;
  P.deg = L.d_deg; P.pdeg = L.d_pdeg; P.off = L.d_off; P.elist = L.d_elist; P.ebegin = 0;
  P.tab = ctx->d_tab;
  return P;
}

int jacobi_setup_dense(Ctx* ctx, Level& L) {
  if (L.nc) { ctx->err = "not available on non-conforming meshes (operator apply only)"; return 1; }
  if (L.jd.ready) return 0;
  if (ctx->nranks > 1 && &L == &ctx->levels.back()) {
    // neighbour penalty degrees across ranks are uniform in the distributed path, so the blocks
    // on rank boundaries are interior blocks; build_dir_matrix would treat them as domain boundary.
    ctx->err = "dense block-Jacobi setup is single-rank only; use the fd form"; return 1;
  }
  size_t total = 0;
  L.jd.bucket_off.assign(L.bucket_p.size() + 1, 0);
  for (size_t b = 0; b < L.bucket_p.size(); b++) {
    int n1 = L.bucket_p[b] + 1; size_t ne = 1;
    for (int d = 0; d < L.dim; d++) ne *= n1;
    L.jd.bucket_off[b] = total;
    total += (size_t)(L.bucket_begin[b + 1] - L.bucket_begin[b]) * ne * ne;
  }
  L.jd.bucket_off[L.bucket_p.size()] = total;
  size_t freeb = 0, totb = 0;
  HPDG_CUDA(cudaMemGetInfo(&freeb, &totb));
  if (total * sizeof(double) > freeb * 0.9) {
    char buf[256];
    snprintf(buf, sizeof buf, "dense block-Jacobi inverses need %.2f GB but only %.2f GB are free; use the fd form",
             total * 8e-9, freeb * 1e-9);
    ctx->err = buf; return 1;
  }
  HPDG_CUDA(cudaMalloc(&L.jd.d_inv, total * sizeof(double)));
  L.jd.bytes = total * sizeof(double);
  BlkParams P = make_blk(ctx, L);
  for (size_t b = 0; b < L.bucket_p.size(); b++) {
    long cnt = L.bucket_begin[b + 1] - L.bucket_begin[b];
    if (!cnt) continue;
    int n1 = L.bucket_p[b] + 1, ne = 1;
    for (int d = 0; d < L.dim; d++) ne *= n1;
    P.ebegin = L.bucket_begin[b];
    double* out = L.jd.d_inv + L.jd.bucket_off[b];
    int threads = ne * ne >= 4096 ? 256 : 128;
    k_build_blocks<<<(unsigned)cnt, threads, 0, ctx->stream>>>(P, out, ne);
    HPDG_CUDA(cudaGetLastError());
    int ithreads = ne >= 200 ? 1024 : ne >= 64 ? 512 : 128;
    k_invert_blocks<<<(unsigned)cnt, ithreads, 2 * ne * sizeof(double), ctx->stream>>>(out, ne);
    HPDG_CUDA(cudaGetLastError());
    ctx->launches += 2;
  }
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream));
  L.jd.ready = true;
  return 0;
}

int jacobi_apply_dense(Ctx* ctx, Level& L, const double* r, double* c, double damping) {
  if (!L.jd.ready) { ctx->err = "hpdg_jacobi_setup(dense) has not been called for this level"; return 1; }
  for (size_t b = 0; b < L.bucket_p.size(); b++) {
    long cnt = L.bucket_begin[b + 1] - L.bucket_begin[b];
    if (!cnt) continue;
    int n1 = L.bucket_p[b] + 1, ne = 1;
    for (int d = 0; d < L.dim; d++) ne *= n1;
    int eper = ne >= 256 ? 1 : 256 / ne;
    int threads = eper * ne;
    threads = (threads + 31) / 32 * 32;
    long blocks = (cnt + eper - 1) / eper;
    k_jacobi_dense<<<(unsigned)blocks, threads, (size_t)eper * ne * sizeof(double), ctx->stream>>>(
        L.jd.d_inv + L.jd.bucket_off[b], L.d_elist, L.bucket_begin[b], cnt, L.d_off, ne, eper, r, c, damping, ctx->fuse_xacc);
    ctx->launches++;
    HPDG_CUDA(cudaGetLastError());
  }
  return 0;
}

int diag_block_device(Ctx* ctx, Level& L, long e, double* d_out) {
  if (L.nc) { ctx->err = "not available on non-conforming meshes (operator apply only)"; return 1; }
  // single-element variant of k_build_blocks through a one-entry element list
  int* d_one = nullptr;
  int ei = (int)e;
  HPDG_CUDA(cudaMalloc(&d_one, sizeof(int)));
  HPDG_CUDA(cudaMemcpyAsync(d_one, &ei, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  BlkParams P = make_blk(ctx, L);
  P.elist = d_one; P.ebegin = 0;
  int n1 = L.deg[e] + 1, ne = 1;
  for (int d = 0; d < L.dim; d++) ne *= n1;
  k_build_blocks<<<1, 256, 0, ctx->stream>>>(P, d_out, ne);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_one);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fast-diagonalisation form
// ---------------------------------------------------------------------------------------------
struct FDParams {
  int dim;
  const int* deg;
  const long* off;
  const int* elist;
  long ebegin;
  const double* fac;
  const int* idx;
  const double* r;
  double* c;
  double damping;
  double* xacc;
};
constexpr int kFacStride = kMaxN * kMaxN + kMaxN;

__global__ void k_jacobi_fd(FDParams P) {
  extern __shared__ double sm[];
  const long e = P.elist[P.ebegin + blockIdx.x];
  const int dim = P.dim, n1 = P.deg[e] + 1;
  const int ne = ipw(n1, dim);
  double* a = sm; double* b = sm + ne;
  const double* re = P.r + P.off[e];
  for (int i = threadIdx.x; i < ne; i += blockDim.x) a[i] = re[i];
  __syncthreads();
  const double* F[3];
  for (int d = 0; d < dim; d++) F[d] = P.fac + (size_t)P.idx[e * 3 + d] * kFacStride;
  // forward: t = (Vx^T x Vy^T x Vz^T) r ; V row-major n1 x n1 (stride n1), eigenvectors in columns
  int sd = 1;
  double* src = a; double* dst = b;
  for (int d = 0; d < dim; d++) {
    const double* V = F[d];
    for (int idx = threadIdx.x; idx < ne; idx += blockDim.x) {
      const int ad = (idx / sd) % n1, base = idx - ad * sd;
      double s = 0;
      for (int k = 0; k < n1; k++) s += V[k * n1 + ad] * src[base + k * sd];
      dst[idx] = s;
    }
    __syncthreads();
    double* t = src; src = dst; dst = t; sd *= n1;
  }
  for (int idx = threadIdx.x; idx < ne; idx += blockDim.x) {
    int rem = idx; double lam = 0;
    for (int d = 0; d < dim; d++) { lam += F[d][kMaxN * kMaxN + rem % n1]; rem /= n1; }
    src[idx] = src[idx] / lam;
  }
  __syncthreads();
  sd = 1;
  for (int d = 0; d < dim; d++) {
    const double* V = F[d];
    const bool last = d == dim - 1;
    for (int idx = threadIdx.x; idx < ne; idx += blockDim.x) {
      const int ad = (idx / sd) % n1, base = idx - ad * sd;
      double s = 0;
      for (int k = 0; k < n1; k++) s += V[ad * n1 + k] * src[base + k * sd];
      if (last) { const double cv = P.damping * s; P.c[P.off[e] + idx] = cv; if (P.xacc) P.xacc[P.off[e] + idx] += cv; } else dst[idx] = s;
    }
    __syncthreads();
    double* t = src; src = dst; dst = t; sd *= n1;
  }
}

static int jacobi_build_fd_generic(Ctx* ctx, Level& L);

int jacobi_setup_fd(Ctx* ctx, Level& L) {
  if (L.nc) { ctx->err = "not available on non-conforming meshes (operator apply only)"; return 1; }
  if (L.jf.ready) return 0;
  // uniform-degree 3-D levels use the tiled kernel (jacobi_uniform.cu) whose 1-D factors live in its parameter block
  if (!uniform_supported(ctx, L)) { if (jacobi_build_fd_generic(ctx, L)) return 1; }
  L.jf.ready = true;
  return 0;
}

static int jacobi_build_fd_generic(Ctx* ctx, Level& L) {
  if (L.jf.d_fac) return 0;
  if (ctx->hp_distributed && ctx->nranks > 1 && hp_ghost_setup(ctx, L)) return 1;   // neighbour degrees across rank boundaries
  const HostTables& H = host_tables();
  typedef std::tuple<int, int, long long, long long, int, int> Key;  // dir-kappa id, p, c0, c1 (bits), w0, w1 (x2)
  std::map<Key, int> seen;
  std::vector<double> fac;
  std::vector<int> idx((size_t)L.nelem * 3, 0);
  for (long e = 0; e < L.nelem; e++) {
    long r = e; int ijk[3];
    ijk[0] = (int)(r % L.n[0]); r /= L.n[0]; ijk[1] = (int)(r % L.n[1]); r /= L.n[1]; ijk[2] = (int)r;
    const int p = L.deg[e], n1 = p + 1;
    for (int d = 0; d < L.dim; d++) {
      double kappa = 1.0 / L.h[d];
      for (int dd = 0; dd < L.dim; dd++) if (dd != d) kappa *= L.h[dd];
      double w[2], c[2];
      for (int s = 0; s < 2; s++) {
        const int cc = ijk[d] + (s ? 1 : -1);
        if (cc >= 0 && cc < L.n[d]) {
          long stride = d == 0 ? 1 : d == 1 ? L.n[0] : (long)L.n[0] * L.n[1];
          long o = e + (s ? stride : -stride);
          int pm = std::max(L.pdeg[e], L.pdeg[o]);
          w[s] = 0.5; c[s] = ctx->sigma * (double)pm * pm;
        } else if (ctx->bnd_is_rank[2 * d + s]) {
          // rank boundary = interior face; the neighbour's degree comes from the ghost layer on hp bricks
          int pm = L.pdeg[e];
          if (ctx->hp_distributed) {
            const int f = 2 * d + s, ta = d == 0 ? 1 : 0, tb = d == 2 ? 1 : 2;
            const long fe = ijk[ta] + (long)L.n[ta] * (L.dim == 3 ? ijk[tb] : 0);
            pm = std::max(pm, L.hpg.h_pdeg[f][fe]);
          }
          w[s] = 0.5; c[s] = ctx->sigma * (double)pm * pm;
        }
        else if (ctx->dirichlet) { w[s] = 1.0; c[s] = ctx->sigma * (double)L.pdeg[e] * L.pdeg[e]; }
        else { w[s] = 0.0; c[s] = 0.0; }
      }
      long long c0b, c1b;
      memcpy(&c0b, &c[0], 8); memcpy(&c1b, &c[1], 8);
      Key key(d, p, c0b, c1b, (int)(w[0] * 2), (int)(w[1] * 2));
      auto it = seen.find(key);
      int id;
      if (it == seen.end()) {
        id = (int)seen.size();
        seen[key] = id;
        const DegTable& T = H.deg[p];
        double D[kMaxN * kMaxN], M[kMaxN * kMaxN], V[kMaxN * kMaxN], lam[kMaxN];
        for (int i = 0; i < n1; i++) for (int j = 0; j < n1; j++) {
          double v = kappa * T.S[i * kMaxN + j];
          for (int s = 0; s < 2; s++) {
            const double nu = s ? 1.0 : -1.0;
            v += -w[s] * nu * kappa * (T.t[s][i] * T.g[s][j] + T.g[s][i] * T.t[s][j]) + c[s] * T.t[s][i] * T.t[s][j];
          }
          D[i * n1 + j] = v; M[i * n1 + j] = T.M[i * kMaxN + j];
        }
        gen_eig(n1, D, M, V, lam);
        fac.resize((size_t)(id + 1) * kFacStride, 0.0);
        double* out = &fac[(size_t)id * kFacStride];
        for (int i = 0; i < n1 * n1; i++) out[i] = V[i];
        for (int i = 0; i < n1; i++) out[kMaxN * kMaxN + i] = lam[i];
      } else id = it->second;
      idx[(size_t)e * 3 + d] = id;
    }
  }
  L.jf.nfac = (int)seen.size();
  HPDG_CUDA(cudaMalloc(&L.jf.d_fac, fac.size() * sizeof(double)));
  HPDG_CUDA(cudaMemcpy(L.jf.d_fac, fac.data(), fac.size() * sizeof(double), cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMalloc(&L.jf.d_idx, idx.size() * sizeof(int)));
  HPDG_CUDA(cudaMemcpy(L.jf.d_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

int jacobi_apply_fd(Ctx* ctx, Level& L, const double* r, double* c, double damping) {
  if (!L.jf.ready) { ctx->err = "hpdg_jacobi_setup(fd) has not been called for this level"; return 1; }
  {
    const int rc = jacobi_apply_fd_uniform(ctx, L, r, c, damping);
    if (rc >= 0) return rc;
  }
  if (jacobi_build_fd_generic(ctx, L)) return 1;
  FDParams P;
  P.dim = L.dim; P.deg = L.d_deg; P.off = L.d_off; P.elist = L.d_elist; P.fac = L.jf.d_fac; P.idx = L.jf.d_idx;
  P.r = r; P.c = c; P.damping = damping; P.xacc = ctx->fuse_xacc;
  for (size_t b = 0; b < L.bucket_p.size(); b++) {
    long cnt = L.bucket_begin[b + 1] - L.bucket_begin[b];
    if (!cnt) continue;
    int n1 = L.bucket_p[b] + 1, ne = 1;
    for (int d = 0; d < L.dim; d++) ne *= n1;
    int threads = ne <= 32 ? 32 : ne <= 64 ? 64 : ne <= 128 ? 128 : 256;
    size_t smem = 2 * (size_t)ne * sizeof(double);
    if (smem > 48 * 1024) HPDG_CUDA(cudaFuncSetAttribute(k_jacobi_fd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P.ebegin = L.bucket_begin[b];
    k_jacobi_fd<<<(unsigned)cnt, threads, smem, ctx->stream>>>(P);
    ctx->launches++;
    HPDG_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace hpdg
