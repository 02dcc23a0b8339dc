// Cycles one warp needs for the softmax block of csrc/attention.cu (32 logits -> 16 packed bf16 pairs), by warps per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench2.bin pipe_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

struct ClampPoly { float c0, c1, c2, c3, c4, wmax, clamp, wlo, ex_a, ex_b; };
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <bool HI>
__device__ __forceinline__ void exp_block(const uint32_t (&v)[32], const ClampPoly& cp, uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float z = __uint_as_float(v[2 * i + e]);
      const float w = z * z;
      float q;
      if (HI) { q = fmaf(w, cp.c4, cp.c3); q = fmaf(w, q, cp.c2); q = fmaf(w, q, cp.c1); } else { q = fmaf(w, cp.c2, cp.c1); }
      q = fmaf(w, q, cp.c0);
      a[e] = z * q;
    }
    pk[i] = pack_bf16(ex2_approx(a[0]), ex2_approx(a[1]));
  }
}

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <bool HI>
__device__ __forceinline__ void exp_block2(const uint32_t (&v)[32], const ClampPoly& cp, uint32_t (&pk)[16]) {
  const uint64_t c0 = pk2(cp.c0, cp.c0), c1 = pk2(cp.c1, cp.c1), c2 = pk2(cp.c2, cp.c2), c3 = pk2(cp.c3, cp.c3), c4 = pk2(cp.c4, cp.c4);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t z = pk2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    const uint64_t w = mul2(z, z);
    uint64_t q;
    if (HI) { q = fma2(w, c4, c3); q = fma2(w, q, c2); q = fma2(w, q, c1); } else { q = fma2(w, c2, c1); }
    q = fma2(w, q, c0);
    const uint64_t a = mul2(z, q);
    float a0, a1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
    pk[i] = pack_bf16(ex2_approx(a0), ex2_approx(a1));
  }
}

template <int MODE>
__global__ void kreg(uint4* out, long long* cyc, int iters, ClampPoly cp) {
  extern __shared__ uint4 sm[];
  uint32_t v[32], pk[16];
  uint4* src = sm;                       // [8][threads]
  uint4* dst = sm + 8 * blockDim.x;      // [4][threads]
  for (int i = threadIdx.x; i < 8 * blockDim.x; i += blockDim.x) src[i] = make_uint4(__float_as_uint(0.001f * i), __float_as_uint(0.002f * i), __float_as_uint(0.5f), __float_as_uint(-0.3f));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u = src[i * blockDim.x + threadIdx.x];
      v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
    }
    if (MODE == 0) exp_block<false>(v, cp, pk);
    if (MODE == 1) exp_block2<false>(v, cp, pk);
    if (MODE == 2) {      // half the ex2
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float z0 = __uint_as_float(v[2 * i]), z1 = __uint_as_float(v[2 * i + 1]);
        const float w0 = z0 * z0, w1 = z1 * z1;
        const float a0 = z0 * fmaf(w0, fmaf(w0, cp.c2, cp.c1), cp.c0), a1 = z1 * fmaf(w1, fmaf(w1, cp.c2, cp.c1), cp.c0);
        pk[i] = pack_bf16(ex2_approx(a0), a1);
      }
    }
    if (MODE == 3) {      // no pack
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float z0 = __uint_as_float(v[2 * i]), z1 = __uint_as_float(v[2 * i + 1]);
        const float w0 = z0 * z0, w1 = z1 * z1;
        const float a0 = z0 * fmaf(w0, fmaf(w0, cp.c2, cp.c1), cp.c0), a1 = z1 * fmaf(w1, fmaf(w1, cp.c2, cp.c1), cp.c0);
        pk[i] = __float_as_uint(ex2_approx(a0) + ex2_approx(a1));
      }
    }
    if (MODE == 4) {      // ex2 only
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = __float_as_uint(ex2_approx(__uint_as_float(v[2 * i])) + ex2_approx(__uint_as_float(v[2 * i + 1])));
    }
    if (MODE == 5) {      // no ex2
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float z0 = __uint_as_float(v[2 * i]), z1 = __uint_as_float(v[2 * i + 1]);
        const float w0 = z0 * z0, w1 = z1 * z1;
        const float a0 = z0 * fmaf(w0, fmaf(w0, cp.c2, cp.c1), cp.c0), a1 = z1 * fmaf(w1, fmaf(w1, cp.c2, cp.c1), cp.c0);
        pk[i] = pack_bf16(a0, a1);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q * blockDim.x + threadIdx.x] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = dst[(threadIdx.x * 7) & 3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void runreg(const char* name, int threads) {
  uint4* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 16); cudaMalloc(&cyc, 148 * 8);
  ClampPoly cp{1.4427f, -1.9e-4f, 6e-8f, -1e-11f, 1e-15f, 625.f, 50.f, 8.f, 0.0577f, 72.1f};
  const int iters = 256;
  cudaFuncSetAttribute(kreg<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 192);
  for (int rep = 0; rep < 2; ++rep) kreg<MODE><<<148, threads, threads * 192>>>(out, cyc, iters, cp);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("REG %-30s warps/SMSP=%d  %.0f clk per block-iteration  (%.1f clk per warp-block per SMSP) %s\n", name, threads / 128, c / iters,
         c / iters / (threads / 128), cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

template <int MODE>
__global__ void k(const uint4* in, uint4* out, long long* cyc, int iters, ClampPoly cp) {
  extern __shared__ uint4 sm[];
  uint32_t v[32], pk[16];
  long long t0 = 0, t1 = 0;
  uint4* dst = sm + threadIdx.x * 4;
  uint4* src = sm + blockDim.x * 4;
  for (int i = threadIdx.x; i < 16 * blockDim.x; i += blockDim.x) src[i] = in[i];
  __syncthreads();
  for (int it = 0; it < iters + 1; ++it) {
    if (it == 1) { __syncthreads(); t0 = clock64(); }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 u = src[((it & 1) * 8 + i) * blockDim.x + threadIdx.x];
      v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
    }
    if (MODE == 2) {
      float wm = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) wm = fmaxf(wm, fabsf(__uint_as_float(v[i])));
      if (!__any_sync(0xffffffffu, wm > cp.wlo)) exp_block2<false>(v, cp, pk); else exp_block2<true>(v, cp, pk);
    } else if (MODE >= 1) {
      float wm = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) wm = fmaxf(wm, fabsf(__uint_as_float(v[i])));
      if (!__any_sync(0xffffffffu, wm > cp.wlo)) exp_block<false>(v, cp, pk); else exp_block<true>(v, cp, pk);
    } else {
      exp_block<false>(v, cp, pk);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
  t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = dst[(threadIdx.x * 7) & 3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, const uint4* in) {
  uint4* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 16); cudaMalloc(&cyc, 148 * 8);
  ClampPoly cp{1.4427f, -1.9e-4f, 6e-8f, -1e-11f, 1e-15f, 625.f, 50.f, 8.f, 0.0577f, 72.1f};
  const int iters = 64;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 200);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, threads, threads * (64 + 256)>>>(in, out, cyc, iters, cp);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-34s warps/SMSP=%d  %.0f clk per block-iteration  (%.1f clk per warp-block per SMSP) %s\n", name, threads / 128, c / iters,
         c / iters / (threads / 128), cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  uint4* in;
  const size_t n = (size_t)65 * 8 * 1024;
  cudaMalloc(&in, n * 16);
  float* h = new float[n * 4];
  for (size_t i = 0; i < n * 4; ++i) h[i] = ((i * 2654435761u) % 2000) / 1000.f - 1.f;
  cudaMemcpy(in, h, n * 16, cudaMemcpyHostToDevice);
  for (int threads : {128, 256, 512}) {
    run<0>("ldg + exp_block<lo> + sts", threads, in);
    run<1>("ldg + max/vote + exp_block + sts", threads, in);
    run<2>("same with f32x2 polynomial", threads, in);
  }
  for (int threads : {128, 256, 512}) {
    runreg<0>("exp_block<lo>", threads);
    runreg<1>("exp_block2<lo> (f32x2)", threads);
    runreg<2>("half the ex2", threads);
    runreg<3>("no pack", threads);
    runreg<4>("ex2 only", threads);
    runreg<5>("no ex2", threads);
  }
  return 0;
}
