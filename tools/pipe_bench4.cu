// Latency of probing an already-completed mbarrier phase: try_wait vs test_wait (one warp, dependent probes).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(long long* out, int iters) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");   // completes phase 0
  __syncthreads();
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t ok;
    if (MODE == 0)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(acc & 0u) : "memory");
    else
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(acc & 0u) : "memory");
    acc += ok;      // dependent chain: the next probe's operand depends on this result
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int threads : {32, 512}) {
      if (mode == 0) k<0><<<1, threads>>>(d, iters); else k<1><<<1, threads>>>(d, iters);
      cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%s, %3d threads: %.1f clk per dependent probe of a completed phase (ok count %lld)\n", mode ? "test_wait" : "try_wait ", threads, (double)h[0] / iters, h[1]);
    }
  }
  return 0;
}
