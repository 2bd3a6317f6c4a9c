// Assembled-matrix side of the path: the SIPG matrix in the reference's DynamicBCRSMatrix layout, its mat-vec, and the
// reference's default smoother on top of it.
//
//  * hpdg_assemble_bcrs  replaces BuildingBlocks::laplace / dynamicStiffnessMatrix (buildingblocks/matrices.hh:29-89,
//    test/testobjects.hh:20-81): block pattern = element + face neighbours (ascending block column), every block a dense
//    row-major rowMap[i] x colMap[j] window into ONE contiguous array, blocks in (block row, ascending block column) order
//    (common/dynamicbcrs.hh:178-199, common/matrixwindow.hh:109).  Blocks are written directly from their Kronecker form
//    (DESIGN.md section 3): A_ee = sum_d M x..x D_d x..x M, A_eo = C_{d,s} x M^{eo} x M^{eo}  -- no quadrature loop.
//  * hpdg_bcrs_mv        replaces BCRSMatrix<MatrixWindow>::mv (common/matrixwindow.hh:196-209).
//  * hpdg_blockgs_iterate replaces DynamicBlockGS::iterate with the GSCore local solver (iterationsteps/dynamicblockgs.hh:17-40,
//    94-126).  The reference sweeps block rows in ascending order; rows on a hyperplane ix+iy+iz = const are mutually
//    independent (they are never face neighbours) and depend only on rows of lower hyperplanes, so sweeping hyperplane by
//    hyperplane reproduces the sequential sweep exactly (same operands for every row), in parallel inside a hyperplane.
#include <algorithm>
#include <cstdio>

#include "ctx.hpp"

namespace hpdg {

struct AsmParams {
  int dim;
  int n[3];
  double h[3];
  double sigma;
  int dirichlet;
  const int* deg;
  const int* pdeg;
  const DegTable* tab;
  const double* Mab;   // rectangular masses [(a*(kMaxP+1)+b)][i*kMaxN+j]
  const long* rowptr;
  const int* col;
  const long* boff;
  const int* brow;     // block row of block k
  double* val;
};

__device__ __forceinline__ int ipwa(int b, int e) { int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }

// one CTA per block
__global__ void k_assemble_blocks(AsmParams P) {
  __shared__ double sN[3][kMaxN * kMaxN];  // normal-direction factors (diagonal block: one per direction)
  const long k = blockIdx.x;
  const long e = P.brow[k], o = P.col[k];
  const int dim = P.dim;
  const int pe = P.deg[e], po = P.deg[o], ne1 = pe + 1, no1 = po + 1;
  const int nre = ipwa(ne1, dim), nco = ipwa(no1, dim);
  const DegTable& Te = P.tab[pe];
  const DegTable& To = P.tab[po];
  double* out = P.val + P.boff[k];
  long r = e; int ijk[3];
  ijk[0] = (int)(r % P.n[0]); r /= P.n[0]; ijk[1] = (int)(r % P.n[1]); r /= P.n[1]; ijk[2] = (int)r;
  int fd = -1, fs = 0;  // face (direction, side) that o lies across, or -1 for the diagonal block
  if (o != e) {
    long ro = o; int ojk[3];
    ojk[0] = (int)(ro % P.n[0]); ro /= P.n[0]; ojk[1] = (int)(ro % P.n[1]); ro /= P.n[1]; ojk[2] = (int)ro;
    for (int d = 0; d < 3; d++) if (ojk[d] != ijk[d]) { fd = d; fs = ojk[d] > ijk[d]; }
  }
  double kap[3];
  for (int d = 0; d < dim; d++) { double kk = 1.0 / P.h[d]; for (int dd = 0; dd < dim; dd++) if (dd != d) kk *= P.h[dd]; kap[d] = kk; }
  if (fd < 0) {
    for (int d = 0; d < dim; d++) {
      double w[2], c[2];
      for (int s = 0; s < 2; s++) {
        const int cc = ijk[d] + (s ? 1 : -1);
        if (cc >= 0 && cc < P.n[d]) {
          const long stride = d == 0 ? 1 : d == 1 ? P.n[0] : (long)P.n[0] * P.n[1];
          const long nb = e + (s ? stride : -stride);
          const int pm = max(P.pdeg[e], P.pdeg[nb]);
          w[s] = 0.5; c[s] = P.sigma * (double)pm * pm;
        } else if (P.dirichlet) { w[s] = 1.0; c[s] = P.sigma * (double)P.pdeg[e] * P.pdeg[e]; }
        else { w[s] = 0.0; c[s] = 0.0; }
      }
      for (int t = threadIdx.x; t < ne1 * ne1; t += blockDim.x) {
        const int i = t / ne1, j = t % ne1;
        double v = kap[d] * Te.S[i * kMaxN + j];
        for (int s = 0; s < 2; s++) {
          const double nu = s ? 1.0 : -1.0;
          v += -w[s] * nu * kap[d] * (Te.t[s][i] * Te.g[s][j] + Te.g[s][i] * Te.t[s][j]) + c[s] * Te.t[s][i] * Te.t[s][j];
        }
        sN[d][t] = v;
      }
    }
  } else {
    // M12-type coupling block seen from e across its face (fd, fs): variableipdg.hh:336-340 (e inside) / :348-352 (e outside)
    const double nu = fs ? 1.0 : -1.0;
    const int pm = max(P.pdeg[e], P.pdeg[o]);
    const double c = P.sigma * (double)pm * pm;
    for (int t = threadIdx.x; t < ne1 * no1; t += blockDim.x) {
      const int i = t / no1, j = t % no1;
      sN[0][t] = Te.t[fs][i] * (-0.5 * nu * kap[fd] * To.g[1 - fs][j] - c * To.t[1 - fs][j]) +
                 Te.g[fs][i] * (0.5 * nu * kap[fd] * To.t[1 - fs][j]);
    }
  }
  __syncthreads();
  const double* Mee = P.Mab + ((size_t)pe * (kMaxP + 1) + pe) * kMaxN * kMaxN;
  const double* Meo = P.Mab + ((size_t)pe * (kMaxP + 1) + po) * kMaxN * kMaxN;
  for (long t = threadIdx.x; t < (long)nre * nco; t += blockDim.x) {
    int a = (int)(t / nco), b = (int)(t % nco);
    int ai[3] = {0, 0, 0}, bi[3] = {0, 0, 0};
    for (int d = 0; d < dim; d++) { ai[d] = a % ne1; a /= ne1; bi[d] = b % no1; b /= no1; }
    double v = 0;
    if (fd < 0) {
      for (int d = 0; d < dim; d++) {
        double f = sN[d][ai[d] * ne1 + bi[d]];
        for (int dd = 0; dd < dim; dd++) if (dd != d) f *= Mee[ai[dd] * kMaxN + bi[dd]];
        v += f;
      }
    } else {
      v = sN[0][ai[fd] * no1 + bi[fd]];
      for (int dd = 0; dd < dim; dd++) if (dd != fd) v *= Meo[ai[dd] * kMaxN + bi[dd]];
    }
    out[t] = v;
  }
}

// y_i = sum_k A_ik x_k : one CTA per block row, thread per row of the block row
__global__ void k_bcrs_mv(const long* __restrict__ rowptr, const int* __restrict__ col, const long* __restrict__ boff,
                          const long* __restrict__ off, const double* __restrict__ val, const double* __restrict__ x,
                          double* __restrict__ y, const int* __restrict__ rows /* optional list */, int mode,
                          const double* __restrict__ b /* mode 1: y = b - A x */) {
  extern __shared__ double xs[];
  const long i = rows ? rows[blockIdx.x] : blockIdx.x;
  const int nr = (int)(off[i + 1] - off[i]);
  double acc[4] = {0, 0, 0, 0};  // up to 4 rows per thread (n_e <= 4*blockDim)
  for (long k = rowptr[i]; k < rowptr[i + 1]; k++) {
    const int j = col[k];
    const int nc = (int)(off[j + 1] - off[j]);
    __syncthreads();
    for (int t = threadIdx.x; t < nc; t += blockDim.x) xs[t] = x[off[j] + t];
    __syncthreads();
    const double* B = val + boff[k];
    for (int q = 0, a = threadIdx.x; a < nr; a += blockDim.x, q++) {
      double s = 0;
      for (int c = 0; c < nc; c++) s = fma(B[(size_t)a * nc + c], xs[c], s);
      acc[q] += s;
    }
  }
  for (int q = 0, a = threadIdx.x; a < nr; a += blockDim.x, q++) y[off[i] + a] = mode ? b[off[i] + a] - acc[q] : acc[q];
}

// GSCore on the block rows of one hyperplane: res holds r_i = b_i - sum_j A_ij x_j; x_i += (L_ii + D_ii)^-1 r_i
// reg != nullptr: the L1Smoother's local solver (iterationsteps/l1smoother.hh:127-145): the diagonal is D_aa + reg_a
__global__ void k_gscore_update(const long* __restrict__ rowptr, const int* __restrict__ col, const long* __restrict__ boff,
                                const long* __restrict__ off, const double* __restrict__ val, const double* __restrict__ res,
                                double* __restrict__ x, const int* __restrict__ rows, const double* __restrict__ reg) {
  extern __shared__ double sc[];  // corr (n)
  const long i = rows[blockIdx.x];
  const int n = (int)(off[i + 1] - off[i]);
  const double* D = nullptr;
  for (long k = rowptr[i]; k < rowptr[i + 1]; k++) if (col[k] == i) D = val + boff[k];
  const int lane = threadIdx.x;  // one warp
  for (int a = 0; a < n; a++) {
    double s = 0;
    for (int c = lane; c < a; c += 32) s = fma(D[(size_t)a * n + c], sc[c], s);
    for (int w = 16; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
    if (lane == 0) {
      const double d = D[(size_t)a * n + a];
      const double dr = reg ? d + reg[off[i] + a] : d;              // l1smoother.hh:144
      sc[a] = (fabs(d) == 0.) ? 0.0 : (res[off[i] + a] - s) / dr;  // dynamicblockgs.hh:26-27,36
    }
    __syncwarp();
  }
  for (int a = lane; a < n; a += 32) x[off[i] + a] += sc[a];
}

int bcrs_build(Ctx* ctx, Level& L) {
  if (L.nc) { ctx->err = "not available on non-conforming meshes (operator apply only)"; return 1; }
  Bcrs& A = L.bcrs;
  if (A.ready) return 0;
  const long ne = L.nelem;
  A.rowptr.assign(ne + 1, 0);
  const long stride[3] = {1, L.n[0], (long)L.n[0] * L.n[1]};
  std::vector<int> brow;
  for (long e = 0; e < ne; e++) {
    long r = e; int ijk[3];
    ijk[0] = (int)(r % L.n[0]); r /= L.n[0]; ijk[1] = (int)(r % L.n[1]); r /= L.n[1]; ijk[2] = (int)r;
    long cols[7]; int c = 0;
    cols[c++] = e;
    for (int d = 0; d < L.dim; d++) for (int s = 0; s < 2; s++) {
      const int cc = ijk[d] + (s ? 1 : -1);
      if (cc >= 0 && cc < L.n[d]) cols[c++] = e + (s ? stride[d] : -stride[d]);
    }
    std::sort(cols, cols + c);
    for (int a = 0; a < c; a++) { A.col.push_back((int)cols[a]); brow.push_back((int)e); }
    A.rowptr[e + 1] = (long)A.col.size();
  }
  const long nb = (long)A.col.size();
  A.boff.assign(nb + 1, 0);
  for (long k = 0; k < nb; k++) {
    const long i = brow[k], j = A.col[k];
    A.boff[k + 1] = A.boff[k] + (L.off[i + 1] - L.off[i]) * (L.off[j + 1] - L.off[j]);
  }
  const size_t nent = (size_t)A.boff[nb];
  size_t freeb = 0, totb = 0;
  HPDG_CUDA(cudaMemGetInfo(&freeb, &totb));
  if (nent * sizeof(double) > freeb * 0.9) {
    char buf[200];
    snprintf(buf, sizeof buf, "assembled matrix needs %.2f GB, only %.2f GB free (use the matrix-free operator)", nent * 8e-9, freeb * 1e-9);
    ctx->err = buf; A = Bcrs(); return 1;
  }
  HPDG_CUDA(cudaMalloc(&A.d_rowptr, sizeof(long) * (ne + 1)));
  HPDG_CUDA(cudaMalloc(&A.d_col, sizeof(int) * nb));
  HPDG_CUDA(cudaMalloc(&A.d_brow, sizeof(int) * nb));
  HPDG_CUDA(cudaMalloc(&A.d_boff, sizeof(long) * (nb + 1)));
  HPDG_CUDA(cudaMalloc(&A.d_val, sizeof(double) * std::max<size_t>(nent, 1)));
  HPDG_CUDA(cudaMemcpy(A.d_rowptr, A.rowptr.data(), sizeof(long) * (ne + 1), cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMemcpy(A.d_col, A.col.data(), sizeof(int) * nb, cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMemcpy(A.d_brow, brow.data(), sizeof(int) * nb, cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMemcpy(A.d_boff, A.boff.data(), sizeof(long) * (nb + 1), cudaMemcpyHostToDevice));
  if (!ctx->d_Mab) {
    const HostTables& H = host_tables();
    HPDG_CUDA(cudaMalloc(&ctx->d_Mab, sizeof(double) * H.Mab.size()));
    HPDG_CUDA(cudaMemcpy(ctx->d_Mab, H.Mab.data(), sizeof(double) * H.Mab.size(), cudaMemcpyHostToDevice));
  }
  AsmParams P;
  P.dim = L.dim;
  for (int d = 0; d < 3; d++) { P.n[d] = L.n[d]; P.h[d] = L.h[d]; }
  P.sigma = ctx->sigma; P.dirichlet = ctx->dirichlet; P.deg = L.d_deg; P.pdeg = L.d_pdeg; P.tab = ctx->d_tab; P.Mab = ctx->d_Mab;
  P.rowptr = A.d_rowptr; P.col = A.d_col; P.boff = A.d_boff; P.brow = A.d_brow; P.val = A.d_val;
  k_assemble_blocks<<<(unsigned)nb, 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  // hyperplane (wavefront) lists for the block Gauss-Seidel sweep
  const int nw = L.n[0] + L.n[1] + L.n[2] - 2;
  std::vector<std::vector<int>> waves(nw);
  for (long e = 0; e < ne; e++) {
    long r = e; const int ix = (int)(r % L.n[0]); r /= L.n[0]; const int iy = (int)(r % L.n[1]); r /= L.n[1];
    waves[ix + iy + (int)r].push_back((int)e);
  }
  A.wave_begin.assign(nw + 1, 0);
  std::vector<int> flat;
  for (int w = 0; w < nw; w++) { flat.insert(flat.end(), waves[w].begin(), waves[w].end()); A.wave_begin[w + 1] = (long)flat.size(); }
  HPDG_CUDA(cudaMalloc(&A.d_wave, sizeof(int) * ne));
  HPDG_CUDA(cudaMemcpy(A.d_wave, flat.data(), sizeof(int) * ne, cudaMemcpyHostToDevice));
  HPDG_CUDA(cudaMalloc(&A.d_res, sizeof(double) * L.ndof));
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream));
  A.ready = true;
  return 0;
}

static int maxblock(const Level& L) { int n1 = L.maxp + 1, ne = 1; for (int d = 0; d < L.dim; d++) ne *= n1; return ne; }

int bcrs_mv(Ctx* ctx, Level& L, const double* x, double* y) {
  Bcrs& A = L.bcrs;
  if (!A.ready) { ctx->err = "hpdg_assemble_bcrs has not been called for this level"; return 1; }
  const int ne = maxblock(L);
  const int threads = std::min(1024, std::max(32, (ne + 3) / 4 > 256 ? (ne + 3) / 4 : std::min(ne, 256)));
  const int thr = (threads + 31) / 32 * 32;
  if ((long)thr * 4 < ne) { ctx->err = "block size too large for the assembled mat-vec kernel"; return 1; }
  k_bcrs_mv<<<(unsigned)L.nelem, thr, ne * sizeof(double), ctx->stream>>>(A.d_rowptr, A.d_col, A.d_boff, L.d_off, A.d_val, x, y, nullptr, 0, nullptr);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

// L1Smoother::preprocess (iterationsteps/l1smoother.hh:31-57): reg_r[j] = sum over the ghost block columns g != r of block row r
// (each counted as often as it appears in the ghost list) of the l1 norm of row j of A[r][g].  One CTA per block row.
__global__ void k_l1_regularization(const long* __restrict__ rowptr, const int* __restrict__ col, const long* __restrict__ boff,
                                    const long* __restrict__ off, const double* __restrict__ val, const int* __restrict__ mult,
                                    double* __restrict__ reg) {
  const long r = blockIdx.x;
  const int nr = (int)(off[r + 1] - off[r]);
  for (int j = threadIdx.x; j < nr; j += blockDim.x) {
    double s = 0;
    for (long k = rowptr[r]; k < rowptr[r + 1]; k++) {
      const long g = col[k];
      if (g == r || mult[g] == 0) continue;
      const int nc = (int)(off[g + 1] - off[g]);
      const double* B = val + boff[k] + (size_t)j * nc;
      double t = 0;
      for (int c = 0; c < nc; c++) t += fabs(B[c]);
      s += mult[g] * t;
    }
    reg[off[r] + j] = s;
  }
}

int l1_setup(Ctx* ctx, Level& L, const long* ghosts, long nghost) {
  Bcrs& A = L.bcrs;
  if (!A.ready) { ctx->err = "hpdg_assemble_bcrs has not been called for this level"; return 1; }
  std::vector<int> mult((size_t)L.nelem, 0);
  for (long q = 0; q < nghost; q++) {
    if (ghosts[q] < 0 || ghosts[q] >= L.nelem) { ctx->err = "ghost block index out of range"; return 1; }
    mult[(size_t)ghosts[q]]++;
  }
  int* d_mult = nullptr;
  HPDG_CUDA(cudaMalloc(&d_mult, sizeof(int) * mult.size()));
  HPDG_CUDA(cudaMemcpy(d_mult, mult.data(), sizeof(int) * mult.size(), cudaMemcpyHostToDevice));
  if (!A.d_l1reg) HPDG_CUDA(cudaMalloc(&A.d_l1reg, sizeof(double) * L.ndof));
  k_l1_regularization<<<(unsigned)L.nelem, 64, 0, ctx->stream>>>(A.d_rowptr, A.d_col, A.d_boff, L.d_off, A.d_val, d_mult, A.d_l1reg);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  HPDG_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_mult);
  A.l1_ready = true;
  return 0;
}

// l1 != 0: L1Smoother::iterate (l1smoother.hh:63-113) -- the same sweep with the regularised local solver
int blockgs_iterate(Ctx* ctx, Level& L, const double* b, double* x, int l1) {
  Bcrs& A = L.bcrs;
  if (!A.ready) { ctx->err = "hpdg_assemble_bcrs has not been called for this level"; return 1; }
  if (l1 && !A.l1_ready) { ctx->err = "hpdg_l1_setup has not been called for this level"; return 1; }
  const int ne = maxblock(L);
  const int thr = (std::min(std::max(32, (ne + 3) / 4 > 256 ? (ne + 3) / 4 : std::min(ne, 256)), 1024) + 31) / 32 * 32;
  const int nw = (int)A.wave_begin.size() - 1;
  for (int w = 0; w < nw; w++) {
    const long cnt = A.wave_begin[w + 1] - A.wave_begin[w];
    if (!cnt) continue;
    const int* rows = A.d_wave + A.wave_begin[w];
    // r_i = b_i - sum_j A_ij x_j over the whole row including the diagonal, with the current x (dynamicblockgs.hh:108-111)
    k_bcrs_mv<<<(unsigned)cnt, thr, ne * sizeof(double), ctx->stream>>>(A.d_rowptr, A.d_col, A.d_boff, L.d_off, A.d_val, x, A.d_res, rows, 1, b);
    k_gscore_update<<<(unsigned)cnt, 32, ne * sizeof(double), ctx->stream>>>(A.d_rowptr, A.d_col, A.d_boff, L.d_off, A.d_val, A.d_res, x, rows, l1 ? A.d_l1reg : nullptr);
    ctx->launches += 2;
  }
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hpdg
