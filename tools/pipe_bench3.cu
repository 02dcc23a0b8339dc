// FFMA vs FFMA2 (fma.rn.f32x2) throughput and their mix on one SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int N2, int N1>   // per iteration: N2 packed + N1 scalar independent chains
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  uint64_t a2[N2 > 0 ? N2 : 1];
  float a1[N1 > 0 ? N1 : 1];
  for (int i = 0; i < N2; ++i) a2[i] = ((uint64_t)__float_as_uint(seed + i) << 32) | __float_as_uint(seed + threadIdx.x);
  for (int i = 0; i < N1; ++i) a1[i] = seed + i + threadIdx.x;
  const uint64_t m2 = ((uint64_t)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (N2 > N1 ? N2 : N1); ++i) {
      if (i < N2) a2[i] = fma2(a2[i], m2, m2);
      if (i < N1) a1[i] = fma1(a1[i], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < N2; ++i) s += __uint_as_float((uint32_t)a2[i]) + __uint_as_float((uint32_t)(a2[i] >> 32));
  for (int i = 0; i < N1; ++i) s += a1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int N2, int N1>
void run(int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<N2, N1><<<148, threads>>>(out, cyc, iters, 0.001f);
  k<N2, N1><<<148, threads>>>(out, cyc, iters, 0.001f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("FFMA2 x%d + FFMA x%d per iter, threads=%4d: %.1f FMA/clk/SM, %.2f instr/clk/SMSP\n", N2, N1, threads,
         (double)threads * iters * (2 * N2 + N1) / c, (double)threads / 128 * iters * (N2 + N1) / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int threads : {256, 512, 1024}) {
    run<0, 8>(threads);
    run<8, 0>(threads);
    run<8, 8>(threads);
    run<8, 4>(threads);
    run<4, 8>(threads);
  }
  return 0;
}
