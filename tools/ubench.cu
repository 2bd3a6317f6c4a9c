// Micro-benchmarks that size the design space of the SIPG apply kernel on B200:
// FP64 FMA rate, FP64 tensor (DMMA m8n8k4) rate, both mixed, L2 and HBM read bandwidth,
// shared-memory LDS.64 bandwidth, SHFL rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = fma(r[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__constant__ double cc[64];
__global__ void k_dfma_const(double* out, int iters) {
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = fma(r[i], cc[i], cc[16 + i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double* out, int iters) {
  double d[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { d[i][0] = threadIdx.x; d[i][1] = i; }
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: per iteration 8 DMMA (8*256 FMA per warp) + 16 DFMA warp instrs (16*32 FMA per warp)
__global__ void k_mixed(double* out, int iters, double fa, double fb) {
  double d[8][2]; double r[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { d[i][0] = threadIdx.x; d[i][1] = i; }
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 1e-3 + i;
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { dmma884(d[i][0], d[i][1], a, b); r[2*i] = fma(r[2*i], fa, fb); r[2*i+1] = fma(r[2*i+1], fa, fb); }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1];
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_read(const double2* __restrict__ in, size_t n2, double* out, int reps) {
  double s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 v = __ldg(in + i); s += v.x + v.y;
    }
  if (s == 1.2345) out[0] = s;
}
__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ o, size_t n2) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += stride) o[i] = in[i];
}

__global__ void k_lds(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0; int idx = threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) { s += sm[(idx + j * 256) & 4095]; }
    idx = (idx + 32) & 4095;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shfl(double* out, int iters) {
  int v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = threadIdx.x + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = __shfl_xor_sync(0xffffffffu, v[j], 1) + 1;
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  double hc[64]; for (int i = 0; i < 64; i++) hc[i] = 1.0 + i * 1e-9;
  CK(cudaMemcpyToSymbol(cc, hc, sizeof(hc)));
  int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32; int blocks = sms * 2;
    float ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf(", \"dfma_tflops_w%d\": %.2f\n", warps * 2, fl / ms * 1e-9);
    ms = timeit([&] { k_dfma_const<<<blocks, threads>>>(out, iters); });
    printf(", \"dfma_const_tflops_w%d\": %.2f\n", warps * 2, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma<<<blocks, threads>>>(out, iters); });
    double flm = 2.0 * 256 * 8 * iters * (double)blocks * warps;
    printf(", \"dmma_tflops_w%d\": %.2f\n", warps * 2, flm / ms * 1e-9);
    ms = timeit([&] { k_mixed<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double flx = (2.0 * 256 * 8 + 2.0 * 32 * 16) * iters * (double)blocks * warps;
    printf(", \"mixed_tflops_w%d\": %.2f\n", warps * 2, flx / ms * 1e-9);
  }
  // memory
  size_t big = (size_t)1 << 30;  // 1 GiB
  double2 *a, *b; CK(cudaMalloc(&a, big)); CK(cudaMalloc(&b, big));
  CK(cudaMemset(a, 0, big)); CK(cudaMemset(b, 0, big));
  {
    float ms = timeit([&] { k_copy<<<sms * 16, 512>>>(a, b, big / 16); });
    printf(", \"hbm_copy_gbs\": %.1f\n", 2.0 * big / ms * 1e-6);
    ms = timeit([&] { k_read<<<sms * 16, 512>>>(a, big / 16, out, 1); });
    printf(", \"hbm_read_gbs\": %.1f\n", 1.0 * big / ms * 1e-6);
  }
  for (size_t mb : {16, 32, 64, 96}) {
    size_t bytes = mb << 20; int reps = 40;
    float ms = timeit([&] { k_read<<<sms * 8, 512>>>(a, bytes / 16, out, reps); });
    printf(", \"l2_read_gbs_%zuMB\": %.1f\n", mb, (double)bytes * reps / ms * 1e-6);
  }
  {
    int threads = 512, blocks = sms * 2; int it2 = 8192;
    CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    float ms = timeit([&] { k_lds<<<blocks, threads, 32768>>>(out, it2); });
    double bytes = 8.0 * 8 * it2 * (double)blocks * threads;
    printf(", \"lds64_tbs\": %.2f, \"lds64_B_per_clk_sm_at_nominal\": %.1f\n", bytes / ms * 1e-9, bytes / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    ms = timeit([&] { k_shfl<<<blocks, threads>>>(out, it2); });
    double sh = 8.0 * it2 * (double)blocks * threads / 32;
    printf(", \"shfl_warpinstr_per_clk_sm_at_nominal\": %.3f\n", sh / (ms * 1e-3) / sms / (p.clockRate * 1e3));
  }
  printf("}\n");
  return 0;
}
