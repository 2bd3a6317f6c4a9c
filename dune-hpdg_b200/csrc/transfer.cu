// p-transfer between DG spaces of different order on the same mesh, and the BLAS-1 helpers of the
// V-cycle.  Replaces DGOrderTransfer::restrict / prolong (transferoperators/ordertransfer.hh:91-119)
// whose block-diagonal matrix holds, per element, the Kronecker product of the 1-D interpolation
// matrix T[i][j] = l^coarse_j(x^fine_i) (transferoperators/dynamicordertransfer.hh:48-73), or the
// identity for elements already at or below the coarse order (ordertransfer.hh:73-77).  The
// Kronecker structure is applied by sum factorisation, never formed.
#include "ctx.hpp"

namespace hpdg {

struct XferParams {
  int dim;
  const int* degf;
  const int* degc;
  const long* offf;
  const long* offc;
  const double* T;  // [(c*(kMaxP+1)+f)][i*kMaxN+j], i fine node, j coarse function
  const double* in;
  double* out;
};

__device__ __forceinline__ int ipw3(int b, int e) { int r = 1; for (int i = 0; i < e; i++) r *= b; return r; }

// One CTA per element.  RESTRICT: out_c = (T x T x T)^T in_f ; else out_f = (T x T x T) in_c.
template <bool RESTRICT>
__global__ void k_transfer(XferParams P) {
  extern __shared__ double sm[];
  const long e = blockIdx.x;
  const int dim = P.dim;
  const int pf = P.degf[e], pc = P.degc[e];
  const int nf = pf + 1, nc = pc + 1;
  const int nin = RESTRICT ? nf : nc, nout = RESTRICT ? nc : nf;
  const double* in = P.in + (RESTRICT ? P.offf[e] : P.offc[e]);
  double* out = P.out + (RESTRICT ? P.offc[e] : P.offf[e]);
  const int tin = ipw3(nin, dim);
  if (pf == pc) {
    for (int i = threadIdx.x; i < tin; i += blockDim.x) out[i] = in[i];
    return;
  }
  const double* T = P.T + ((size_t)pc * (kMaxP + 1) + pf) * kMaxN * kMaxN;
  const int mx = ipw3(nf, dim);
  double* a = sm; double* b = sm + mx;
  for (int i = threadIdx.x; i < tin; i += blockDim.x) a[i] = in[i];
  __syncthreads();
  // contract one direction at a time; extents ext[d] switch from nin to nout
  int ext[3] = {1, 1, 1};
  for (int d = 0; d < dim; d++) ext[d] = nin;
  double* src = a; double* dst = b;
  for (int d = 0; d < dim; d++) {
    int oext[3] = {ext[0], ext[1], ext[2]};
    oext[d] = nout;
    const int tot = oext[0] * oext[1] * oext[2];
    const int sin_d = d == 0 ? 1 : d == 1 ? ext[0] : ext[0] * ext[1];
    const bool last = d == dim - 1;
    for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
      int i0 = idx % oext[0], i1 = (idx / oext[0]) % oext[1], i2 = idx / (oext[0] * oext[1]);
      int id[3] = {i0, i1, i2};
      const int o = id[d];
      id[d] = 0;
      const int base = id[0] + ext[0] * (id[1] + ext[1] * id[2]);
      double s = 0;
      for (int k = 0; k < nin; k++) {
        const double t = RESTRICT ? T[k * kMaxN + o] : T[o * kMaxN + k];
        s += t * src[base + k * sin_d];
      }
      if (last) out[idx] = s; else dst[idx] = s;
    }
    __syncthreads();
    ext[d] = nout;
    double* t = src; src = dst; dst = t;
  }
}

// Uniform level -> uniform level in 3-D: a warp per element, all extents compile-time, 4 elements per CTA.
template <int NF, int NC, bool RESTRICT>
__global__ void __launch_bounds__(128) k_transfer_uniform(const double* __restrict__ T /* [i fine][j coarse], stride kMaxN */,
                                                          const double* __restrict__ in, double* __restrict__ out, long nelem) {
  constexpr int NIN = RESTRICT ? NF : NC, NOUT = RESTRICT ? NC : NF;
  constexpr int F3 = NF * NF * NF;
  __shared__ double sT[NF * NC];
  __shared__ double buf[4][2][F3];
  for (int t = threadIdx.x; t < NF * NC; t += blockDim.x) sT[t] = T[(t / NC) * kMaxN + (t % NC)];
  __syncthreads();
  const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long e = (long)blockIdx.x * 4 + w;
  if (e >= nelem) return;
  constexpr int TIN = NIN * NIN * NIN, TOUT = NOUT * NOUT * NOUT;
  const double* ein = in + e * TIN;
  double* eout = out + e * TOUT;
  double* a = buf[w][0]; double* b = buf[w][1];
  for (int i = lane; i < TIN; i += 32) a[i] = ein[i];
  __syncwarp();
  // x: (NIN, NIN, NIN) -> (NOUT, NIN, NIN)
  for (int idx = lane; idx < NOUT * NIN * NIN; idx += 32) {
    const int o = idx % NOUT, r = idx / NOUT;
    double s = 0;
#pragma unroll
    for (int k = 0; k < NIN; k++) s = fma(RESTRICT ? sT[k * NC + o] : sT[o * NC + k], a[k + NIN * r], s);
    b[idx] = s;
  }
  __syncwarp();
  // y: (NOUT, NIN, NIN) -> (NOUT, NOUT, NIN)
  for (int idx = lane; idx < NOUT * NOUT * NIN; idx += 32) {
    const int i0 = idx % NOUT, o = (idx / NOUT) % NOUT, i2 = idx / (NOUT * NOUT);
    double s = 0;
#pragma unroll
    for (int k = 0; k < NIN; k++) s = fma(RESTRICT ? sT[k * NC + o] : sT[o * NC + k], b[i0 + NOUT * (k + NIN * i2)], s);
    a[idx] = s;
  }
  __syncwarp();
  // z: (NOUT, NOUT, NIN) -> (NOUT, NOUT, NOUT)
  for (int idx = lane; idx < TOUT; idx += 32) {
    const int r = idx % (NOUT * NOUT), o = idx / (NOUT * NOUT);
    double s = 0;
#pragma unroll
    for (int k = 0; k < NIN; k++) s = fma(RESTRICT ? sT[k * NC + o] : sT[o * NC + k], a[r + NOUT * NOUT * k], s);
    eout[idx] = s;
  }
}

template <int NF, int NC>
static int xfer_uniform(Ctx* ctx, long nelem, const double* in, double* out, bool restrict_) {
  const double* T = ctx->d_T + ((size_t)(NC - 1) * (kMaxP + 1) + (NF - 1)) * kMaxN * kMaxN;
  const unsigned grid = (unsigned)((nelem + 3) / 4);
  if (restrict_) k_transfer_uniform<NF, NC, true><<<grid, 128, 0, ctx->stream>>>(T, in, out, nelem);
  else k_transfer_uniform<NF, NC, false><<<grid, 128, 0, ctx->stream>>>(T, in, out, nelem);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

static XferParams make_x(Ctx* ctx, Level& fine, Level& coarse, const double* in, double* out) {
  XferParams P;
  P.dim = fine.dim; P.degf = fine.d_deg; P.degc = coarse.d_deg; P.offf = fine.d_off; P.offc = coarse.d_off;
  P.T = ctx->d_T; P.in = in; P.out = out;
  return P;
}

static int xfer_launch(Ctx* ctx, Level& fine, Level& coarse, const double* in, double* out, bool restrict_) {
  if (fine.dim == 3 && fine.uniform && coarse.uniform && !ctx->force_generic) {
    const int key = (fine.p_uni + 1) * 16 + (coarse.p_uni + 1);
    switch (key) {
      case 3 * 16 + 2: return xfer_uniform<3, 2>(ctx, fine.nelem, in, out, restrict_);
      case 4 * 16 + 2: return xfer_uniform<4, 2>(ctx, fine.nelem, in, out, restrict_);
      case 5 * 16 + 3: return xfer_uniform<5, 3>(ctx, fine.nelem, in, out, restrict_);
      case 6 * 16 + 3: return xfer_uniform<6, 3>(ctx, fine.nelem, in, out, restrict_);
      case 7 * 16 + 4: return xfer_uniform<7, 4>(ctx, fine.nelem, in, out, restrict_);
      default: break;
    }
  }
  XferParams P = make_x(ctx, fine, coarse, in, out);
  int n1 = fine.maxp + 1, mx = 1;
  for (int d = 0; d < fine.dim; d++) mx *= n1;
  size_t smem = 2 * (size_t)mx * sizeof(double);
  int threads = mx <= 32 ? 32 : mx <= 64 ? 64 : mx <= 128 ? 128 : 256;
  if (restrict_) {
    if (smem > 48 * 1024) HPDG_CUDA(cudaFuncSetAttribute(k_transfer<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_transfer<true><<<(unsigned)fine.nelem, threads, smem, ctx->stream>>>(P);
  } else {
    if (smem > 48 * 1024) HPDG_CUDA(cudaFuncSetAttribute(k_transfer<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_transfer<false><<<(unsigned)fine.nelem, threads, smem, ctx->stream>>>(P);
  }
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

int launch_restrict(Ctx* ctx, Level& fine, Level& coarse, const double* xf, double* xc) {
  return xfer_launch(ctx, fine, coarse, xf, xc, true);
}
int launch_prolong(Ctx* ctx, Level& fine, Level& coarse, const double* xc, double* xf) {
  return xfer_launch(ctx, fine, coarse, xc, xf, false);
}

}  // namespace hpdg
