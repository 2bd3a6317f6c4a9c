// Do FP64 FMAs and shared-memory accesses overlap on B200?  Per loop iteration a warp issues NF independent DFMAs and NL
// LDS.64 (+ NS STS.64); the time of the mix is compared with the time of each part alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_overlap ubench_overlap.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template <int NF, int NL, int NS>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-9;
  __syncthreads();
  double r[16], l[8];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 8; i++) l[i] = 0;
  int idx = threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NF; i++) r[i % 16] = fma(r[i % 16], a, b);
#pragma unroll
    for (int j = 0; j < NL; j++) l[j % 8] += 0 * sm[(idx + j * 256) & 4095] , l[j % 8] = sm[(idx + j * 256) & 4095];
#pragma unroll
    for (int j = 0; j < NS; j++) sm[(idx + j * 256 + 128) & 4095] = r[j % 16];
    idx = (idx + 32) & 4095;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
#pragma unroll
  for (int i = 0; i < 8; i++) s += l[i];
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

template <int NF, int NL, int NS> void run(const char* name, double* out, int sms) {
  const int iters = 4096, blocks = sms * 3, threads = 256;
  CK(cudaFuncSetAttribute(k_mix<NF, NL, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  float ms = timeit([&] { k_mix<NF, NL, NS><<<blocks, threads, 65536>>>(out, iters, 1.0000001, 1e-9); });
  // cycles per iteration per SM sub-partition (6 warps each) at 1965 MHz
  printf("%-28s NF=%2d NL=%2d NS=%2d  %.3f ms   %.1f clk/iter/warp-slot\n", name, NF, NL, NS, ms, ms * 1e-3 * 1.965e9 / iters / 6);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 3 * 256));
  run<16, 0, 0>("dfma only", out, sms);
  run<0, 4, 0>("lds only", out, sms);
  run<0, 8, 0>("lds only", out, sms);
  run<0, 0, 4>("sts only", out, sms);
  run<16, 4, 0>("dfma + lds", out, sms);
  run<16, 8, 0>("dfma + lds", out, sms);
  run<16, 0, 4>("dfma + sts", out, sms);
  run<16, 4, 2>("dfma + lds + sts", out, sms);
  run<32, 8, 4>("dfma + lds + sts", out, sms);
  return 0;
}
