// What can a streaming kernel reach on this part for a given read : write mix?  (The HBM-side kernels of the path -- depthwise conv,
// out-projection epilogues, guided Euler -- write more than they read; MEASURED_PEAKS.json only has the 1 : 1 copy figure.)
// Plain coalesced float4 accesses, grid = 148 x 16 CTAs of 256 threads, 1 Gi elements; GB/s of read + written bytes.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__global__ void k_read(const float4* __restrict__ a, size_t n4, float* sink) {
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = a[i];
    s += v.x + v.y + v.z + v.w;
  }
  if (s == 123.456f) *sink = s;
}
__global__ void k_write(float4* __restrict__ a, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) a[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
// 4 B read : 6 B written per element (fp32 in, fp32 + bf16 out): the depthwise conv / residual epilogue mix
__global__ void k_mix(const float4* __restrict__ a, float4* __restrict__ b, uint2* __restrict__ c, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = a[i];
    b[i] = v;
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    c[i] = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
  }
}
// in place: 4 B read + 4 B written at the same address (+ 2 B bf16 copy): the residual stream
__global__ void k_rmw(float4* __restrict__ a, uint2* __restrict__ c, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = a[i];
    v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
    a[i] = v;
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    c[i] = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
  }
}

template <typename F>
static void run(const char* name, double bytes, F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  float best = 1e9f;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-46s %8.1f us  %7.1f GB/s\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9);
}

int main() {
  const size_t n = 1ull << 28, n4 = n / 4;         // 1 GiB of fp32 per buffer (>> the 126 MB L2)
  float4 *a, *b; uint2* c; float* sink;
  cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&c, n * 2); cudaMalloc(&sink, 4);
  cudaMemset(a, 0, n * 4);
  for (int mult : {8, 16, 32}) {
    const int grid = 148 * mult;
    printf("grid = 148 x %d CTAs of 256 threads\n", mult);
    run("read only (4 B / element)", n * 4.0, [&] { k_read<<<grid, 256>>>(a, n4, sink); });
    run("write only (4 B / element)", n * 4.0, [&] { k_write<<<grid, 256>>>(b, n4); });
    run("copy (4 B read : 4 B written)", n * 8.0, [&] { k_copy<<<grid, 256>>>(a, b, n4); });
    run("fp32 in, fp32 + bf16 out (4 read : 6 written)", n * 10.0, [&] { k_mix<<<grid, 256>>>(a, b, c, n4); });
    run("in-place fp32 update + bf16 copy (4 : 4 + 2)", n * 10.0, [&] { k_rmw<<<grid, 256>>>(a, c, n4); });
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
