// The depthwise conv's memory skeleton without its arithmetic: TMA boxes [ROWS x CH channels] of fp32 in (3-D map over [batch, N, C],
// optional 30-row halo), every element written back once as fp32 (+ optionally bf16), persistent CTAs, NBUF tile buffers.
// Which part of the access shape keeps the real kernel at 3.5-3.9 TB/s when a flat streaming kernel reaches 5.7-6.2?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I video-to-audio-and-piano-rp_b200/csrc -o tools/hbm_conv_shape_bench.bin tools/hbm_conv_shape_bench.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace e2b;

template <int T, int CH, int HALO, int NBUF, bool BF16, bool ROWMAJOR>
__global__ void __launch_bounds__(256, 2) skel(const __grid_constant__ CUtensorMap tmx, float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
                                               int batch, int N, int C) {
  constexpr int ROWS = T + HALO;
  extern __shared__ __align__(128) float buf[];
  uint64_t* full = reinterpret_cast<uint64_t*>(buf + NBUF * ROWS * CH);
  const int cchunks = C / CH, rtiles = (N + T - 1) / T, per_chunk = batch * rtiles, total = per_chunk * cchunks;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto decode = [&](int tile, int& cc, int& b, int& rt) {
    int rem;
    if (ROWMAJOR) { cc = tile % cchunks; rem = tile / cchunks; } else { cc = tile / per_chunk; rem = tile - cc * per_chunk; }
    b = rem / rtiles; rt = rem - b * rtiles;
  };
  auto issue = [&](int tile, int s) {
    int cc, b, rt;
    decode(tile, cc, b, rt);
    mbar_arrive_expect_tx(&full[s], ROWS * CH * 4);
    tma_load_3d(buf + (size_t)s * ROWS * CH, &tmx, &full[s], cc * CH, rt * T - HALO / 2, b);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < NBUF - 1; ++i)
      if ((int)blockIdx.x + i * (int)gridDim.x < total) issue(blockIdx.x + i * gridDim.x, i);
  constexpr int TPR = CH / 4;                 // threads per row (float4 each)
  constexpr int RPP = 256 / TPR;              // rows per pass
  const int c4 = threadIdx.x % TPR, rr = threadIdx.x / TPR;
  int it = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int s = it % NBUF;
    if (threadIdx.x == 0 && tile + (NBUF - 1) * (int)gridDim.x < total) issue(tile + (NBUF - 1) * gridDim.x, (it + NBUF - 1) % NBUF);
    int cc, b, rt;
    decode(tile, cc, b, rt);
    mbar_wait(&full[s], (it / NBUF) & 1);
    const float* xs = buf + (size_t)s * ROWS * CH + (HALO / 2) * CH;
#pragma unroll 4
    for (int r = rr; r < T; r += RPP) {
      const int row = rt * T + r;
      if (row < N) {
        const float4 v = *reinterpret_cast<const float4*>(xs + r * CH + c4 * 4);
        const size_t off = ((size_t)b * N + row) * C + cc * CH + c4 * 4;
        *reinterpret_cast<float4*>(y + off) = v;
        if (BF16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
          *reinterpret_cast<uint2*>(yb + off) = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
        }
      }
    }
    __syncthreads();
  }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int T, int CH, int HALO, int NBUF, bool BF16, bool ROWMAJOR>
static void run(const char* name, PFN_enc enc, float* x, float* y, __nv_bfloat16* yb, int batch, int N, int C) {
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)batch}, strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)N * C * 4};
  cuuint32_t box[3] = {CH, T + HALO, 1}, es[3] = {1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("%s: encode failed\n", name); return; }
  const int smem = NBUF * (T + HALO) * CH * 4 + 64;
  auto k = skel<T, CH, HALO, NBUF, BF16, ROWMAJOR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(e0);
    k<<<296, 256, smem>>>(tm, y, yb, batch, N, C);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  const double bytes = (double)batch * N * C * (BF16 ? 10.0 : 8.0);
  printf("%-86s %7.1f us  %7.1f GB/s  (%s)\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode entry point\n"); return 1; }
  PFN_enc enc = reinterpret_cast<PFN_enc>(p);
  const int batch = 128, N = 782, C = 1024;
  const size_t n = (size_t)batch * N * C;
  float *x, *y; __nv_bfloat16* yb;
  cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&yb, n * 2);
  cudaMemset(x, 0, n * 4);
  printf("128 x 782 rows x 1024 channels fp32; GB/s of algorithmic bytes (8 or 10 B / element); 296 CTAs of 256 threads, 2 per SM\n");
  run<64, 128, 30, 2, false, false>("conv shape: 64+30 rows x 128 ch, 2 buffers, chunk-major, fp32 out", enc, x, y, yb, batch, N, C);
  run<64, 128, 30, 2, true, false>("conv shape + bf16 out", enc, x, y, yb, batch, N, C);
  run<64, 128, 0, 2, false, false>("no halo: 64 rows x 128 ch", enc, x, y, yb, batch, N, C);
  run<64, 128, 30, 2, false, true>("conv shape, channel chunk fastest (row-major tile order)", enc, x, y, yb, batch, N, C);
  run<64, 256, 30, 1, false, false>("64+30 rows x 256 ch, 1 buffer (wider rows: 1 KB segments)", enc, x, y, yb, batch, N, C);
  run<32, 256, 30, 2, false, false>("32+30 rows x 256 ch, 2 buffers", enc, x, y, yb, batch, N, C);
  run<32, 128, 30, 3, false, false>("32+30 rows x 128 ch, 3 buffers", enc, x, y, yb, batch, N, C);
  run<32, 128, 0, 4, false, false>("no halo: 32 rows x 128 ch, 4 buffers", enc, x, y, yb, batch, N, C);
  run<16, 128, 0, 4, false, true>("no halo: 16 rows x 128 ch, 4 buffers, row-major order", enc, x, y, yb, batch, N, C);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
