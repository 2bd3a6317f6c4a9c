// Do FP64 FMAs and shared-memory accesses overlap on B200?  Per loop iteration a warp issues NF independent DFMAs and NL
// 128-bit shared-memory loads (4 wavefronts each) whose results are folded with 4 integer XORs; 3 CTAs x 256 threads per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_overlap ubench_overlap.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template <int NF, int NL>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
  extern __shared__ uint4 sm4[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm4[i] = make_uint4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 1e-3 + i;
  unsigned acc = 0;
  const volatile uint4* p = sm4 + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NF; i++) r[i % 16] = fma(r[i % 16], a, b);
#pragma unroll
    for (int j = 0; j < NL; j++) { uint4 v; v.x = p[j * 256].x; v.y = p[j * 256].y; v.z = p[j * 256].z; v.w = p[j * 256].w; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  }
  double s = acc;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
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

template <int NF, int NL> void run(double* out, int sms) {
  const int iters = 4096, blocks = sms * 3, threads = 256;
  CK(cudaFuncSetAttribute(k_mix<NF, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  float ms = timeit([&] { k_mix<NF, NL><<<blocks, threads, 65536>>>(out, iters, 1.0000001, 1e-9); });
  // SM cycles per iteration (all 24 warps of the SM do one iteration each) at 1965 MHz
  printf("NF=%2d DFMA  NL=%d LDS.128 (%3d wavefronts/SM/iter)  %.3f ms  %.0f clk/iter/SM\n", NF, NL, NL * 4 * 24, ms, ms * 1e-3 * 1.965e9 / iters);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 3 * 256));
  run<16, 0>(out, sms); run<0, 1>(out, sms); run<0, 2>(out, sms); run<0, 4>(out, sms);
  run<16, 1>(out, sms); run<16, 2>(out, sms); run<16, 4>(out, sms); run<32, 2>(out, sms);
  return 0;
}
