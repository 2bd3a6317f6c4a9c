// BLAS-1 on the flat DynamicBlockVector storage (reference: common/dynamicbvector.hh:185-314: operator=, *=, +=, -=, axpy,
// operator* (dot), two_norm) and the vector updates of the Krylov loop with DEVICE-resident scalars, so that a whole
// preconditioned-CG iteration is enqueued without a host round trip (csrc/api.cu: pcg_device).
// All kernels are HBM bound: grid-stride loops, 8 CTAs per SM; the dot product is a deterministic two-stage reduction
// (fixed grid, fixed tree) whose scratch lives in the context (= per device).
#include "ctx.hpp"

namespace hpdg {

static int grid_for(long n) { long b = (n + 255) / 256; return (int)(b < 148 * 8 ? (b < 1 ? 1 : b) : 148 * 8); }

__global__ void k_axpy(long n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] += a * x[i];
}
__global__ void k_sub(long n, const double* __restrict__ b, const double* __restrict__ ax, double* __restrict__ r) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) r[i] = b[i] - ax[i];
}
__global__ void k_scale(long n, double a, double* __restrict__ x) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] *= a;
}
// CG: alpha = s[num] / s[den];  x += alpha p;  r -= alpha q
__global__ void k_cg_update(long n, const double* __restrict__ s, int num, int den, const double* __restrict__ p,
                            const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r) {
  const double alpha = s[num] / s[den];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    r[i] = fma(-alpha, q[i], r[i]);
  }
}
// CG: beta = s[num] / s[den];  p = z + beta p
__global__ void k_cg_direction(long n, const double* __restrict__ s, int num, int den, const double* __restrict__ z,
                               double* __restrict__ p) {
  const double beta = s[num] / s[den];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}

constexpr int kDotBlocks = 592, kDotThreads = 256;
// the same update with the new residual's r . r folded in (one pass over r instead of two): per-CTA partial sums over the fixed
// grid of the two-stage dot product, so the result does not depend on anything but n
__global__ void k_cg_update_rr(long n, const double* __restrict__ s, int num, int den, const double* __restrict__ p,
                               const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, double* __restrict__ part) {
  __shared__ double sh[kDotThreads];
  const double alpha = s[num] / s[den];
  double acc = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double rn = fma(-alpha, q[i], r[i]);
    r[i] = rn;
    acc = fma(rn, rn, acc);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int w = kDotThreads / 2; w > 0; w >>= 1) { if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void k_dot1(long n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ part) {
  __shared__ double sh[kDotThreads];
  double s = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s = fma(x[i], y[i], s);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = kDotThreads / 2; w > 0; w >>= 1) { if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void k_dot2(const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double sh[1024];
  double s = 0;
  for (int i = threadIdx.x; i < kDotBlocks; i += blockDim.x) s += part[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) { if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = sh[0];
}

int blas_scratch(Ctx* ctx) {
  if (!ctx->d_partial) HPDG_CUDA(cudaMalloc(&ctx->d_partial, kDotBlocks * sizeof(double)));
  if (!ctx->d_scalar) {
    HPDG_CUDA(cudaMalloc(&ctx->d_scalar, kScalarSlots * sizeof(double)));
    HPDG_CUDA(cudaMemset(ctx->d_scalar, 0, kScalarSlots * sizeof(double)));
  }
  return 0;
}

int launch_axpy(Ctx* ctx, long n, double a, const double* x, double* y) {
  k_axpy<<<grid_for(n), 256, 0, ctx->stream>>>(n, a, x, y);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}
int launch_xpay_sub(Ctx* ctx, long n, const double* b, const double* ax, double* r) {
  k_sub<<<grid_for(n), 256, 0, ctx->stream>>>(n, b, ax, r);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}
int launch_scale(Ctx* ctx, long n, double a, double* x) {
  k_scale<<<grid_for(n), 256, 0, ctx->stream>>>(n, a, x);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}
int launch_dot(Ctx* ctx, long n, const double* x, const double* y, double* d_result) {
  if (blas_scratch(ctx)) return 1;
  k_dot1<<<kDotBlocks, kDotThreads, 0, ctx->stream>>>(n, x, y, ctx->d_partial);
  k_dot2<<<1, 1024, 0, ctx->stream>>>(ctx->d_partial, d_result);
  ctx->launches += 2;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}
int launch_cg_update(Ctx* ctx, long n, int num, int den, const double* p, const double* q, double* x, double* r) {
  k_cg_update<<<grid_for(n), 256, 0, ctx->stream>>>(n, ctx->d_scalar, num, den, p, q, x, r);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}
// x += a p; r -= a q; *d_rr = r . r (this rank's part) in one pass
int launch_cg_update_rr(Ctx* ctx, long n, int num, int den, const double* p, const double* q, double* x, double* r, double* d_rr) {
  if (blas_scratch(ctx)) return 1;
  k_cg_update_rr<<<kDotBlocks, kDotThreads, 0, ctx->stream>>>(n, ctx->d_scalar, num, den, p, q, x, r, ctx->d_partial);
  k_dot2<<<1, 1024, 0, ctx->stream>>>(ctx->d_partial, d_rr);
  ctx->launches += 2;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}
int launch_cg_direction(Ctx* ctx, long n, int num, int den, const double* z, double* p) {
  k_cg_direction<<<grid_for(n), 256, 0, ctx->stream>>>(n, ctx->d_scalar, num, den, z, p);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hpdg
