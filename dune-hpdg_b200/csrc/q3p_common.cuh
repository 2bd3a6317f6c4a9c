// Device helpers shared by the persistent Q3 tile kernels (apply_uniform_q3p.cuh, apply_uniform_q3e.cuh, jacobi_uniform_q3p.cuh):
// mbarrier / bulk-copy wrappers, compile-time loops, the opaque thread index and the half-CTA named barrier.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>

namespace hpdg {

__device__ __forceinline__ uint32_t q3p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void q3p_mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(q3p_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void q3p_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(q3p_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool q3p_mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(q3p_smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void q3p_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(q3p_smem_u32(dst)), "l"(src), "r"(bytes), "r"(q3p_smem_u32(b)) : "memory");
}
// shared -> global bulk store, committed as this thread's own bulk group (wait with cp.async.bulk.wait_group.read)
__device__ __forceinline__ void q3p_bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(q3p_smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}

// shared -> global bulk reduction dst += src (FP64 add performed at the L2), committed as this thread's own bulk group
__device__ __forceinline__ void q3p_bulk_s2g_add(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;\n" ::"l"(dst), "r"(q3p_smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}

template <int... I, class F>
__device__ __forceinline__ void q3p_for_impl(std::integer_sequence<int, I...>, F f) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void q3p_for(F f) { q3p_for_impl(std::make_integer_sequence<int, N>{}, f); }

// threadIdx.x re-read as an opaque value: the role indices derived from it are recomputed where a pass needs them
// instead of being kept alive (and spilled) across the whole tile loop
__device__ __forceinline__ int q3p_tid() {
  int t;
  asm volatile("mov.u32 %0, %%tid.x;\n" : "=r"(t));
  return t;
}

__device__ __forceinline__ void q3p_bar_half(int half) {
  if (half) asm volatile("bar.sync 2, 128;\n" ::: "memory");
  else asm volatile("bar.sync 1, 128;\n" ::: "memory");
}

}  // namespace hpdg
