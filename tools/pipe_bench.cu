// Throughput of the instruction mixes used by the attention softmax on one SM (lanes per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);                                   // MUFU only
      if (MODE == 1) { acc ^= pack(a[i], a[(i + 1) & 7]); a[i] += 1.0f; }   // F2FP + FADD
      if (MODE == 2) a[i] = fmaf(a[i], 1.0001f, 0.5f);                   // FFMA only
      if (MODE == 3) { float w = a[i] * a[i]; float q = fmaf(w, 0.1f, 0.2f); q = fmaf(w, q, 1.4f); a[i] = ex2(a[i] * q); }   // poly + ex2
      if (MODE == 4) { float w = a[i] * a[i]; float q = fmaf(w, 0.1f, 0.2f); q = fmaf(w, q, 1.4f); float e = ex2(a[i] * q); acc ^= pack(e, e); a[i] = e; }
      if (MODE == 5) { a[i] = fmaxf(fabsf(a[i]), fabsf(a[(i + 3) & 7])) + 1.0f; }   // FMNMX + FADD
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.001f);
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.001f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-28s threads=%4d  %.2f elements/clk/SM  (%.1f clk per warp-element per SMSP)\n", name, threads, (double)threads * iters * 8 / c,
         c / ((double)threads / 128 * iters * 8));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 512, 1024}) {
    run<0>("MUFU.EX2", threads);
    run<1>("F2FP.pack + FADD", threads);
    run<2>("FFMA", threads);
    run<3>("poly(4 fma) + EX2", threads);
    run<4>("poly + EX2 + F2FP", threads);
    run<5>("FMNMX + FADD", threads);
  }
  return 0;
}
