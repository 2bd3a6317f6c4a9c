// BLAS-1 on the flat DynamicBlockVector storage (reference: common/dynamicbvector.hh:185-314: operator=, *=, +=, -=, axpy,
// operator* (dot), two_norm) and the vector updates of the Krylov loop with DEVICE-resident scalars, so that a whole
// preconditioned-CG iteration is enqueued without a host round trip (csrc/api.cu: pcg_device).
// All kernels are HBM bound: grid-stride loops, 8 CTAs per SM; the dot product is a deterministic two-stage reduction
// (fixed grid, fixed tree) whose scratch lives in the context (= per device).
#include <cstdint>

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
// The Krylov-loop kernels move 16-byte words when the vectors allow it (cudaMalloc'ed vectors always do): half the load / store
// instructions per byte.  VEC = 2: i counts double2 words, the odd tail element (if any) is handled by the last thread.
template <int VEC> struct WordOf { using type = double; };
template <> struct WordOf<2> { using type = double2; };
__device__ __forceinline__ double2 fma2(double a, double2 x, double2 y) { return make_double2(fma(a, x.x, y.x), fma(a, x.y, y.y)); }
__device__ __forceinline__ double fma2(double a, double x, double y) { return fma(a, x, y); }
__device__ __forceinline__ double sq2(double2 v, double acc) { return fma(v.y, v.y, fma(v.x, v.x, acc)); }
__device__ __forceinline__ double sq2(double v, double acc) { return fma(v, v, acc); }
__device__ __forceinline__ double dot2w(double2 a, double2 b, double acc) { return fma(a.y, b.y, fma(a.x, b.x, acc)); }
__device__ __forceinline__ double dot2w(double a, double b, double acc) { return fma(a, b, acc); }
static bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
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
template <int VEC>
__global__ void k_cg_direction(long n, const double* __restrict__ s, int num, int den, const double* __restrict__ z_,
                               double* __restrict__ p_) {
  using W = typename WordOf<VEC>::type;
  const W* __restrict__ z = reinterpret_cast<const W*>(z_);
  W* __restrict__ p = reinterpret_cast<W*>(p_);
  const double beta = s[num] / s[den];
  const long nw = n / VEC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nw; i += (long)gridDim.x * blockDim.x) p[i] = fma2(beta, p[i], z[i]);
  if (VEC == 2 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p_[n - 1] = fma(beta, p_[n - 1], z_[n - 1]);
}

constexpr int kDotBlocks = 592, kDotThreads = 256;
// the same update with the new residual's r . r folded in (one pass over r instead of two): per-CTA partial sums over the fixed
// grid of the two-stage dot product, so the result does not depend on anything but n
template <int VEC>
__global__ void k_cg_update_rr(long n, const double* __restrict__ s, int num, int den, const double* __restrict__ p_,
                               const double* __restrict__ q_, double* __restrict__ x_, double* __restrict__ r_, double* __restrict__ part) {
  using W = typename WordOf<VEC>::type;
  const W* __restrict__ p = reinterpret_cast<const W*>(p_);
  const W* __restrict__ q = reinterpret_cast<const W*>(q_);
  W* __restrict__ x = reinterpret_cast<W*>(x_);
  W* __restrict__ r = reinterpret_cast<W*>(r_);
  __shared__ double sh[kDotThreads];
  const double alpha = s[num] / s[den];
  double acc = 0;
  const long nw = n / VEC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nw; i += (long)gridDim.x * blockDim.x) {
    x[i] = fma2(alpha, p[i], x[i]);
    const W rn = fma2(-alpha, q[i], r[i]);
    r[i] = rn;
    acc = sq2(rn, acc);
  }
  if (VEC == 2 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    x_[n - 1] = fma(alpha, p_[n - 1], x_[n - 1]);
    const double rn = fma(-alpha, q_[n - 1], r_[n - 1]);
    r_[n - 1] = rn;
    acc = fma(rn, rn, acc);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int w = kDotThreads / 2; w > 0; w >>= 1) { if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
template <int VEC>
__global__ void k_dot1(long n, const double* __restrict__ x_, const double* __restrict__ y_, double* __restrict__ part) {
  using W = typename WordOf<VEC>::type;
  const W* __restrict__ x = reinterpret_cast<const W*>(x_);
  const W* __restrict__ y = reinterpret_cast<const W*>(y_);
  __shared__ double sh[kDotThreads];
  double s = 0;
  const long nw = n / VEC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nw; i += (long)gridDim.x * blockDim.x) s = dot2w(x[i], y[i], s);
  if (VEC == 2 && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) s = fma(x_[n - 1], y_[n - 1], s);
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
  if (aligned16(x, y)) k_dot1<2><<<kDotBlocks, kDotThreads, 0, ctx->stream>>>(n, x, y, ctx->d_partial);
  else k_dot1<1><<<kDotBlocks, kDotThreads, 0, ctx->stream>>>(n, x, y, ctx->d_partial);
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
  if (aligned16(p, q, x, r)) k_cg_update_rr<2><<<kDotBlocks, kDotThreads, 0, ctx->stream>>>(n, ctx->d_scalar, num, den, p, q, x, r, ctx->d_partial);
  else k_cg_update_rr<1><<<kDotBlocks, kDotThreads, 0, ctx->stream>>>(n, ctx->d_scalar, num, den, p, q, x, r, ctx->d_partial);
  k_dot2<<<1, 1024, 0, ctx->stream>>>(ctx->d_partial, d_rr);
  ctx->launches += 2;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}
int launch_cg_direction(Ctx* ctx, long n, int num, int den, const double* z, double* p) {
  if (aligned16(z, p)) k_cg_direction<2><<<grid_for(n / 2), 256, 0, ctx->stream>>>(n, ctx->d_scalar, num, den, z, p);
  else k_cg_direction<1><<<grid_for(n), 256, 0, ctx->stream>>>(n, ctx->d_scalar, num, den, z, p);
  ctx->launches++;
  HPDG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hpdg
