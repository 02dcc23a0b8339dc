// HBM-bound glue kernels of the E2 transformer forward and the sampler epilogue (sm_100a).
//   rmsnorm        x-transformers RMSNorm / AdaptiveRMSNorm  (F.normalize eps 1e-12, * sqrt(d) * scale)
//   dwconv         DepthwiseConv (e2_tts_crossatt3.py:495-528) + the caller's residual add (:1082,1097,1122)
//   time_mlp/gemv  time_cond_mlp (:555-564,793-797) and every AdaptiveRMSNorm.to_gamma / AdaLNZero.to_gamma (:532-551)
//   init_stream    register-token prepend + condition drop (:975-997, 2040-2044)
//   guided_euler   cfg combine (+APG projection :162-173, 2106-2113), K-pass mix and the Euler update
#include <math.h>

#include "kernels.h"
#include "prof.h"
#include "ptx.cuh"

namespace e2b {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// x * sigmoid(x) as two MUFU ops and three FMA-pipe ops, branch-free (__fdividef carries a range check per call)
__device__ __forceinline__ float silu(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

// ------------------------------------------------------------------------------------------------ rmsnorm
constexpr int NORM_MAX_V4 = 16;   // C <= 2048

// NV4 = float4 per lane the row may need (C <= 128 NV4): sizing the register tile to the row (8 for C = 1024, 10 for 1280, 4 for
// 512) instead of the maximum keeps the kernel at <= 64 registers, i.e. 32 resident warps per SM instead of 16 -- the kernel
// is latency bound (ncu: DRAM 60-68 %, long-scoreboard 6.5-7.7 stalls per issue), so bytes in flight are what it needs.
template <int OUT_MODE, int NV4>   // OUT_MODE: 0 bf16, 1 fp32, 2 bf16 (hi, lo) pair with lo at column + C
__global__ void __launch_bounds__(256, NV4 <= 10 ? 4 : 2) rmsnorm_kernel(const float* __restrict__ x, int ldx, void* __restrict__ yv, int ldy,
                                                      const float* __restrict__ scale, int scale_bstride, int rows_out_total,
                                                      int rows_per_batch, int skip_rows, int C, float sqrt_c) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows_out_total) return;
  const int rpo = rows_per_batch - skip_rows;
  const int b = warp / rpo, i = warp % rpo;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)(b * rows_per_batch + skip_rows + i) * ldx);
  const int nv = C >> 2;
  float4 v[NV4];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV4; ++k) {
    const int idx = lane + 32 * k;
    if (idx < nv) {
      v[k] = xr[idx];
      ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
  }
  ss = warp_sum(ss);
  const float inv = sqrt_c / fmaxf(sqrtf(ss), 1e-12f);
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * scale_bstride);
#pragma unroll
  for (int k = 0; k < NV4; ++k) {
    const int idx = lane + 32 * k;
    if (idx < nv) {
      const float4 s = __ldg(sc + idx);
      const float4 o = make_float4(v[k].x * inv * s.x, v[k].y * inv * s.y, v[k].z * inv * s.z, v[k].w * inv * s.w);
      if constexpr (OUT_MODE == 1) {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(yv) + (size_t)warp * ldy)[idx] = o;
      } else {
        __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(yv) + (size_t)warp * ldy;
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
        uint2 hi;
        hi.x = *reinterpret_cast<const uint32_t*>(&h0);
        hi.y = *reinterpret_cast<const uint32_t*>(&h1);
        reinterpret_cast<uint2*>(yr)[idx] = hi;
        if constexpr (OUT_MODE == 2) {
          const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
          uint2 lo;
          lo.x = pack_bf16(o.x - f0.x, o.y - f0.y);
          lo.y = pack_bf16(o.z - f1.x, o.w - f1.y);
          reinterpret_cast<uint2*>(yr + C)[idx] = lo;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ dwconv
// y = x + mask * silu(conv31(mask * x) + bias), channels-last.  TMA-fed: a persistent CTA streams [64+30 rows x 128
// channels] fp32 tiles of x into a double-buffered shared-memory ring (3-D tensor map over [batch, N, C]: rows outside
// the sequence are zero-filled by the TMA unit, which is exactly the conv's zero padding).  One thread = one channel;
// it slides an 8-output register window down the tile, so every x element is read from shared memory once and feeds
// 31 FMAs with 8 independent accumulators.  (The first version -- global loads into a register ring -- was latency
// bound: ncu 8 % DRAM, long-scoreboard 4.3 stalls/issue.)
constexpr int DW_T = 64;            // output rows per tile
constexpr int DW_C = 128;           // channels per tile
constexpr int DW_G = 16;            // outputs per register-window step (two steps per thread)

int make_tmap_generic(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box);

// Sum over the 32 lanes of a warp of 16 per-lane values at once (transposing butterfly, 16 shuffles): on return lane l holds
// the warp total of v[(l >> 1) & 15].  Fixed exchange pattern: deterministic.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
  float t8[8], t4[4], t2[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool hi = lane & 16;
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? v[i] : v[i + 8], 16);
    t8[i] = (hi ? v[i + 8] : v[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane & 8;
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? t8[i] : t8[i + 4], 8);
    t4[i] = (hi ? t8[i + 4] : t8[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = lane & 4;
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? t4[i] : t4[i + 2], 4);
    t2[i] = (hi ? t4[i + 2] : t4[i]) + recv;
  }
  const bool hi = lane & 2;
  const float recv = __shfl_xor_sync(0xffffffffu, hi ? t2[0] : t2[1], 2);
  const float one = (hi ? t2[1] : t2[0]) + recv;
  return one + __shfl_xor_sync(0xffffffffu, one, 1);
}

// One thread: 32 consecutive outputs of one channel.  FULL = no row of the tile (halo included) is masked, which is every
// tile except the last one of a sequence: the per-element length tests disappear from the hot loop.
// NORM: the conv output feeds an RMSNorm and then a GEMM (e2_tts_crossatt3.py:1082-1084, 1122-1126): also emit the bf16 A operand
// bf16(y * gain[c]) (lane pairs exchange values so that a lane stores two adjacent channels) and, per warp, the sums of y^2 over its
// 32 channels for every row (wpart[g + row], written by the even lanes) -- the GEMM scales its accumulator rows by sqrt(C) / ||y||.
template <int KS, bool FULL, bool NORM>
__device__ __forceinline__ void dwconv_rows(const float* __restrict__ xs, float* __restrict__ yp, const float (&w)[KS], float bs, int C,
                                            int row0 /*first input row of xs[0]*/, int out0 /*first output row*/, int len, int N,
                                            __nv_bfloat16* __restrict__ ybp /*row out0, channel c & ~1*/, float gain, float* __restrict__ wpart,
                                            int lane) {
  constexpr int HALO = KS - 1, HALF = KS / 2;
  auto ldx = [&](int j) -> float {
    if (FULL) return xs[(size_t)j * DW_C];
    return (row0 + j < len) ? xs[(size_t)j * DW_C] : 0.f;
  };
  float win[HALO + DW_G];
#pragma unroll
  for (int j = 0; j < HALO; ++j) win[j] = ldx(j);
#pragma unroll
  for (int g = 0; g < DW_T / 2; g += DW_G) {
    if (!FULL && out0 + g >= N) break;
#pragma unroll
    for (int o = 0; o < DW_G; ++o) win[HALO + o] = ldx(HALO + g + o);
    float acc[DW_G];
#pragma unroll
    for (int o = 0; o < DW_G; ++o) acc[o] = bs;
#pragma unroll
    for (int i = 0; i < KS; ++i)
#pragma unroll
      for (int o = 0; o < DW_G; ++o) acc[o] = fmaf(w[i], win[o + i], acc[o]);
#pragma unroll
    for (int o = 0; o < DW_G; ++o) {
      if (FULL) {
        acc[o] = win[HALF + o] + silu(acc[o]);
        yp[(size_t)(g + o) * C] = acc[o];
      } else {
        const int r = out0 + g + o;
        // the residual uses the UNMASKED x (e2_tts_crossatt3.py:1082: conv(x, mask) + x)
        acc[o] = (r < len) ? win[HALF + o] + silu(acc[o]) : xs[(size_t)(HALF + g + o) * DW_C];
        if (r < N) yp[(size_t)(g + o) * C] = acc[o];
      }
    }
    if constexpr (NORM) {
      // bf16 copy: the even lane of a channel pair stores rows 0..7 of the window, the odd lane rows 8..15 -- one exchange and one
      // full-warp 4-byte store per row pair, no divergence
      const bool odd = lane & 1;
#pragma unroll
      for (int o = 0; o < DW_G / 2; ++o) {
        const float lo = acc[o] * gain, hi = acc[o + DW_G / 2] * gain;
        const float recv = __shfl_xor_sync(0xffffffffu, odd ? lo : hi, 1);
        const int ro = odd ? o + DW_G / 2 : o;
        const uint32_t pk = odd ? pack_bf16(recv, hi) : pack_bf16(lo, recv);
        if (FULL || out0 + g + ro < N) *reinterpret_cast<uint32_t*>(ybp + (size_t)(g + ro) * C) = pk;
      }
#pragma unroll
      for (int o = 0; o < DW_G; ++o) acc[o] *= acc[o];
      const float tot = warp_reduce16(acc, lane);
      if (!(lane & 1)) wpart[g + ((lane >> 1) & 15)] = tot;
    }
#pragma unroll
    for (int j = 0; j < HALO; ++j) win[j] = win[j + DW_G];
  }
}

template <int KS, bool NORM>
__global__ void __launch_bounds__(2 * DW_C, 2) dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmx, float* __restrict__ y,
                                                                 const float* __restrict__ wt, const float* __restrict__ bias,
                                                                 const int* __restrict__ lens, int batch, int N, int C,
                                                                 __nv_bfloat16* __restrict__ yb, const float* __restrict__ gain,
                                                                 float* __restrict__ rss, int rss_ld) {
  constexpr int HALO = KS - 1, HALF = KS / 2, ROWS = DW_T + HALO;
  // no static shared memory: the dynamic window starts at offset 0 (128-byte alignment needed by the TMA destination) and
  // is used directly so loads stay LDS (a uintptr_t round-trip would make them generic LD.E)
  extern __shared__ __align__(128) float buf[];               // 2 x [ROWS][DW_C], then two mbarriers, then (NORM) 2 x [8 warps][32 rows]
  uint64_t* full = reinterpret_cast<uint64_t*>(buf + 2 * ROWS * DW_C);
  float* wparts = buf + 2 * ROWS * DW_C + 4;
  const int tid = threadIdx.x & (DW_C - 1);       // channel inside the tile
  const int rhalf = threadIdx.x / DW_C;           // which 32-row half of the tile this thread produces
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cchunks = (C + DW_C - 1) / DW_C, rtiles = (N + DW_T - 1) / DW_T;
  const int per_chunk = batch * rtiles;           // tiles of one channel chunk: chunk-major order, so a CTA keeps its taps
  const int total = per_chunk * cchunks;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmx);
  }
  __syncthreads();
  auto issue = [&](int tile, int s) {
    const int cc = tile / per_chunk, rem = tile - cc * per_chunk, b = rem / rtiles, rt = rem - b * rtiles;
    mbar_arrive_expect_tx(&full[s], ROWS * DW_C * 4);
    tma_load_3d(buf + (size_t)s * ROWS * DW_C, &tmx, &full[s], cc * DW_C, rt * DW_T - HALF, b);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < total) issue(blockIdx.x, 0);

  float w[KS];
  float bs = 0.f, gn = 1.f;
  int wcc = -1;
  int it = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int s = it & 1;
    if (threadIdx.x == 0 && tile + (int)gridDim.x < total) issue(tile + gridDim.x, s ^ 1);
    const int cc = tile / per_chunk, rem = tile - cc * per_chunk, b = rem / rtiles, rt = rem - b * rtiles;
    const int c = cc * DW_C + tid;
    const int r0 = rt * DW_T;
    const int len = lens ? min(N, __ldg(lens + b)) : N;
    if (cc != wcc && c < C) {                     // taps change only with the channel chunk
#pragma unroll
      for (int i = 0; i < KS; ++i) w[i] = __ldg(wt + (size_t)i * C + c);
      bs = __ldg(bias + c);
      if (NORM) gn = gain ? __ldg(gain + c) : 1.f;
      wcc = cc;
    }
    mbar_wait(&full[s], (it >> 1) & 1);
    float* wpart = wparts + (s * 8 + warp) * 32;  // this warp's row sums for this tile (double-buffered across tiles)
    if (c < C) {                                  // whole warps: C is a multiple of 32 in NORM mode
      const int g0 = rhalf * (DW_T / 2);
      const float* xs = buf + (size_t)s * ROWS * DW_C + (size_t)g0 * DW_C + tid;      // xs[j * DW_C] = x[r0 + g0 - HALF + j]
      float* yp = y + ((size_t)b * N + r0 + g0) * C + c;
      __nv_bfloat16* ybp = NORM ? yb + ((size_t)b * N + r0 + g0) * C + (c & ~1) : nullptr;
      if (r0 + DW_T + HALF <= len) dwconv_rows<KS, true, NORM>(xs, yp, w, bs, C, r0 + g0 - HALF, r0 + g0, len, N, ybp, gn, wpart, lane);
      else dwconv_rows<KS, false, NORM>(xs, yp, w, bs, C, r0 + g0 - HALF, r0 + g0, len, N, ybp, gn, wpart, lane);
    } else if (NORM) {
      if (!(lane & 1)) { wpart[(lane >> 1) & 15] = 0.f; wpart[16 + ((lane >> 1) & 15)] = 0.f; }
    }
    __syncthreads();     // everyone is done with buffer s before it is refilled two tiles later; the warps' row sums are complete
    if (NORM && threadIdx.x < DW_T) {
      // the 128 channels of this tile, per row: the four channel warps of the row's half, in warp order (deterministic)
      const int rr = threadIdx.x, hf = rr >> 5, wi = rr & 31;
      const float* p = wparts + (s * 8 + hf * 4) * 32 + wi;
      const float tot = ((p[0] + p[32]) + p[64]) + p[96];
      if (r0 + rr < N) rss[(size_t)cc * rss_ld + (size_t)b * N + r0 + rr] = tot;
    }
  }
}

// ------------------------------------------------------------------------------------------------ time conditioning
__global__ void __launch_bounds__(256) time_mlp_kernel(const float* __restrict__ times, const float* __restrict__ fw,
                                                       const float* __restrict__ w1, const float* __restrict__ b1, int dim,
                                                       float* __restrict__ tcond) {
  extern __shared__ float emb[];   // dim + 1
  const int s = blockIdx.x;
  const float t = times[s];
  const int half = dim / 2;
  for (int i = threadIdx.x; i <= dim; i += blockDim.x) {
    float v;
    if (i == 0) v = t;
    else {
      const int j = (i - 1) % half;
      float f = t * fw[j];
      f = f * 2.0f;
      f = f * 3.14159265358979323846f;
      v = (i - 1 < half) ? sinf(f) : cosf(f);
    }
    emb[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < dim; j += nw) {
    const float* wr = w1 + (size_t)j * (dim + 1);
    float acc = 0.f;
    for (int k = lane; k <= dim; k += 32) acc = fmaf(wr[k], emb[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) tcond[(size_t)s * dim + j] = silu(acc + b1[j]);
  }
}

constexpr int GEMV_MAX_K = 64;    // dim <= 2048
__global__ void __launch_bounds__(256) time_gemv_kernel(const float* __restrict__ tcond, int nt, int dim, const float* const* __restrict__ w,
                                                        const float* const* __restrict__ bvec, const int* __restrict__ act, int nmat,
                                                        float* __restrict__ out) {
  const int m = blockIdx.y;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= dim) return;
  const float* wr = w[m] + (size_t)j * dim;
  float wreg[GEMV_MAX_K];
  const int nk = (dim + 31) >> 5;
#pragma unroll
  for (int k = 0; k < GEMV_MAX_K; ++k)
    if (k < nk) wreg[k] = (lane + 32 * k < dim) ? __ldg(wr + lane + 32 * k) : 0.f;
  const float bj = bvec[m] ? bvec[m][j] : 0.f;
  const int a = act[m];
  for (int s = 0; s < nt; ++s) {
    const float* tc = tcond + (size_t)s * dim;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < GEMV_MAX_K; ++k)
      if (k < nk) acc = fmaf(wreg[k], (lane + 32 * k < dim) ? __ldg(tc + lane + 32 * k) : 0.f, acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + bj;
      out[((size_t)s * nmat + m) * dim + j] = (a == 0) ? v + 1.0f : 1.0f / (1.0f + expf(-v));
    }
  }
}

// ------------------------------------------------------------------------------------------------ stream init
__global__ void __launch_bounds__(256) init_stream_kernel(float* __restrict__ dst, __nv_bfloat16* __restrict__ dst_b16,
                                                          const float* __restrict__ regs, const float* __restrict__ src, int src_batches,
                                                          const unsigned char* __restrict__ drop, const float* __restrict__ add_table,
                                                          int batch, int n, int R, int C, int regs_only, int ldb, int b16_split) {
  const int nv = C >> 2;
  const int rows = regs_only ? R : (R + n);
  const size_t total = (size_t)batch * rows * nv;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int v = idx % nv;
    const size_t rr = idx / nv;
    const int pos = rr % rows, b = rr / rows;
    float4 val;
    if (pos < R) val = __ldg(reinterpret_cast<const float4*>(regs + (size_t)pos * C) + v);
    else {
      if (src && !(drop && drop[b])) val = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(b % src_batches) * n + (pos - R)) * C) + v);
      else val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (add_table) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(add_table + (size_t)(pos - R) * C) + v);
        val.x += a.x; val.y += a.y; val.z += a.z; val.w += a.w;
      }
    }
    const size_t o = ((size_t)b * (R + n) + pos) * C;
    reinterpret_cast<float4*>(dst + o)[v] = val;
    if (dst_b16) {
      __nv_bfloat16* br = dst_b16 + ((size_t)b * (R + n) + pos) * ldb;
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(val.x, val.y), h1 = __floats2bfloat162_rn(val.z, val.w);
      uint2 u;
      u.x = *reinterpret_cast<const uint32_t*>(&h0);
      u.y = *reinterpret_cast<const uint32_t*>(&h1);
      reinterpret_cast<uint2*>(br)[v] = u;
      if (b16_split > 0) {
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        uint2 lo;
        lo.x = pack_bf16(val.x - f0.x, val.y - f0.y);
        lo.y = pack_bf16(val.z - f1.x, val.w - f1.y);
        reinterpret_cast<uint2*>(br + b16_split)[v] = lo;
      }
    }
  }
}

__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd,
                                                       size_t rows, int C) {
  const size_t total = rows * ldd;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = idx % ldd;
    const size_t r = idx / ldd;
    dst[idx] = __float2bfloat16_rn(c < C ? src[r * lds + c] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ guided Euler
constexpr int EULER_MAX_P = 8;
struct EulerW { float w[EULER_MAX_P]; };

// scratch[2*b] += <p0 - p1, p0>, scratch[2*b+1] += <p0, p0>  in fp64 (reference `project` runs in double)
__global__ void __launch_bounds__(256) apg_reduce_kernel(const float* __restrict__ pred, size_t pass_stride, size_t per_sample,
                                                         double* __restrict__ scratch) {
  const int b = blockIdx.y;
  const float4* p0 = reinterpret_cast<const float4*>(pred + (size_t)b * per_sample);
  const float4* p1 = reinterpret_cast<const float4*>(pred + pass_stride + (size_t)b * per_sample);
  double a = 0.0, c = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 u = p0[i], v = p1[i];
    a += (double)(u.x - v.x) * u.x + (double)(u.y - v.y) * u.y + (double)(u.z - v.z) * u.z + (double)(u.w - v.w) * u.w;
    c += (double)u.x * u.x + (double)u.y * u.y + (double)u.z * u.z + (double)u.w * u.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __shared__ double sa[8], sc[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sa[warp] = a; sc[warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tc = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ta += sa[i]; tc += sc[i]; }
    atomicAdd(scratch + 2 * b, ta);
    atomicAdd(scratch + 2 * b + 1, tc);
  }
}

__global__ void __launch_bounds__(256) guided_euler_kernel(float* __restrict__ y, const float* __restrict__ pred, int P, size_t pass_stride,
                                                           size_t per_sample, size_t total4, EulerW gw, float dt, int apg,
                                                           float keep_parallel, const double* __restrict__ scratch,
                                                           __nv_bfloat16* __restrict__ yb, int n_copies,
                                                           const float* __restrict__ inpaint, const int* __restrict__ inpaint_lens,
                                                           int row4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 p0 = reinterpret_cast<const float4*>(pred)[i];
    float v[4] = {p0.x, p0.y, p0.z, p0.w};
    const float q0[4] = {p0.x, p0.y, p0.z, p0.w};
    if (apg) {
      // upd = p0 - p1 ; par = (<upd,p0>/max(|p0|,1e-12)^2) p0 ; upd' = (upd - par) + par*keep   (fp64 like the reference)
      const int b = (int)((i * 4) / per_sample);
      const double a = scratch[2 * b];
      const double nrm = fmax(sqrt(scratch[2 * b + 1]), 1e-12);
      const double coef = a / (nrm * nrm);
      const float4 p1 = reinterpret_cast<const float4*>(pred + pass_stride)[i];
      const float q1[4] = {p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double upd = (double)(q0[e] - q1[e]);
        const double par = coef * (double)q0[e];
        const float orth = (float)(upd - par);
        const float upd2 = orth + (float)par * keep_parallel;
        v[e] = q0[e] + upd2 * gw.w[0];
      }
    } else {
      for (int k = 1; k < P; ++k) {
        const float4 pk = reinterpret_cast<const float4*>(pred + (size_t)k * pass_stride)[i];
        const float wk = gw.w[k - 1];
        v[0] += (q0[0] - pk.x) * wk;
        v[1] += (q0[1] - pk.y) * wk;
        v[2] += (q0[2] - pk.z) * wk;
        v[3] += (q0[3] - pk.w) * wk;
      }
    }
    float4 yy = reinterpret_cast<float4*>(y)[i];
    yy.x += dt * v[0]; yy.y += dt * v[1]; yy.z += dt * v[2]; yy.w += dt * v[3];
    if (inpaint) {
      // last update of an in-painting call: out = where(cond_mask, cond, out)  (e2_tts_crossatt3.py:2259-2260)
      const size_t ps4 = per_sample / 4;
      const int b = (int)(i / ps4);
      const int pos = (int)((i - (size_t)b * ps4) / row4);
      if (pos < __ldg(inpaint_lens + b)) yy = __ldg(reinterpret_cast<const float4*>(inpaint) + i);
    }
    reinterpret_cast<float4*>(y)[i] = yy;
    if (yb) {
      uint2 u;
      u.x = pack_bf16(yy.x, yy.y);
      u.y = pack_bf16(yy.z, yy.w);
      for (int cpy = 0; cpy < n_copies; ++cpy) reinterpret_cast<uint2*>(yb + (size_t)cpy * total4 * 4)[i] = u;
    }
  }
}

static int grid_for(size_t work_items, int block, int max_blocks = 148 * 16) {
  size_t g = (work_items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (size_t)max_blocks) g = max_blocks;
  return (int)g;
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("%s launch: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // namespace e2b

using namespace e2b;

extern "C" int e2b_rmsnorm_launch(const float* x, int ldx, void* y, int ldy, const float* scale, int scale_bstride, int batch,
                                  int rows_per_batch, int skip_rows, int C, int out_mode, cudaStream_t stream) {
  if (C % 4 || C > NORM_MAX_V4 * 128 || ldx % 4 || ldy % 4) { e2b_set_kernel_error("rmsnorm: C=%d unsupported", C); return -1; }
  if (out_mode == 2 && ldy < 2 * C) { e2b_set_kernel_error("rmsnorm: split output needs ldy >= 2C"); return -1; }
  const int rows_out = batch * (rows_per_batch - skip_rows);
  if (rows_out <= 0) return 0;
  const int blocks = (rows_out + 7) / 8;
  ProfScope ps(stream, "rmsnorm", rows_out, C, 0, 3.0 * rows_out * C, (double)rows_out * C * (out_mode == 1 ? 8.0 : (out_mode == 2 ? 8.0 : 6.0)));
  const float sq = sqrtf((float)C);
  const int nv4 = (C / 4 + 31) / 32;
#define E2B_NORM_LAUNCH(MODE_, NV_) \
  rmsnorm_kernel<MODE_, NV_><<<blocks, 256, 0, stream>>>(x, ldx, y, ldy, scale, scale_bstride, rows_out, rows_per_batch, skip_rows, C, sq)
#define E2B_NORM_MODE(MODE_)                      \
  do {                                            \
    if (nv4 <= 4) E2B_NORM_LAUNCH(MODE_, 4);      \
    else if (nv4 <= 8) E2B_NORM_LAUNCH(MODE_, 8); \
    else if (nv4 <= 10) E2B_NORM_LAUNCH(MODE_, 10); \
    else E2B_NORM_LAUNCH(MODE_, NORM_MAX_V4);     \
  } while (0)
  if (out_mode == 1) E2B_NORM_MODE(1);
  else if (out_mode == 2) E2B_NORM_MODE(2);
  else E2B_NORM_MODE(0);
#undef E2B_NORM_MODE
#undef E2B_NORM_LAUNCH
  return check_launch("rmsnorm");
}

extern "C" int e2b_dwconv_launch(const float* x, float* y, const float* w, const float* bias, const int* lens, int batch, int N,
                                 int C, int ksize, cudaStream_t stream) {
  return e2b_dwconv_norm_launch(x, y, w, bias, lens, batch, N, C, ksize, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int e2b_dwconv_norm_launch(const float* x, float* y, const float* w, const float* bias, const int* lens, int batch, int N,
                                      int C, int ksize, void* y_b16, const float* gain, float* row_ss, int row_ss_ld, cudaStream_t stream) {
  if (ksize != 31) { e2b_set_kernel_error("dwconv: kernel size %d unsupported (31 only)", ksize); return -1; }
  if (batch <= 0 || N <= 0) return 0;
  if (C % 4) { e2b_set_kernel_error("dwconv: C must be a multiple of 4"); return -1; }
  const bool norm = y_b16 != nullptr;
  if (norm && (C % 32 || !row_ss || row_ss_ld < batch * N || (reinterpret_cast<uintptr_t>(y_b16) & 3))) {
    e2b_set_kernel_error("dwconv: the normed-operand outputs need C %% 32 == 0, row_ss and row_ss_ld >= batch * N");
    return -1;
  }
  CUtensorMap tm;
  const uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)batch};
  const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)N * C * 4};
  const uint32_t box[3] = {DW_C, DW_T + 30, 1};
  if (make_tmap_generic(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box)) return -1;
  const int smem = 2 * (DW_T + 30) * DW_C * 4 + 16 + 2 * 8 * 32 * 4;
  static bool configured[E2B_MAX_DEVICES] = {false};
  bool& conf = configured[e2b_device_slot()];
  if (!conf) {
    if (cudaFuncSetAttribute(dwconv_tma_kernel<31, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(dwconv_tma_kernel<31, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      e2b_set_kernel_error("dwconv: shared memory attribute failed");
      return -1;
    }
    conf = true;
  }
  const int total = batch * ((N + DW_T - 1) / DW_T) * ((C + DW_C - 1) / DW_C);
  const int grid = total < 2 * 148 ? total : 2 * 148;
  ProfScope ps(stream, "dwconv", (long long)batch * N, C, ksize, 2.0 * batch * N * (double)C * ksize, (norm ? 10.0 : 8.0) * batch * N * (double)C);
  if (norm)
    dwconv_tma_kernel<31, true><<<grid, 2 * DW_C, smem, stream>>>(tm, y, w, bias, lens, batch, N, C, reinterpret_cast<__nv_bfloat16*>(y_b16), gain,
                                                                   row_ss, row_ss_ld);
  else
    dwconv_tma_kernel<31, false><<<grid, 2 * DW_C, smem, stream>>>(tm, y, w, bias, lens, batch, N, C, nullptr, nullptr, nullptr, 0);
  return check_launch("dwconv");
}

extern "C" int e2b_time_mlp_launch(const float* times, int nt, const float* fourier_w, const float* w1, const float* b1, int dim,
                                   float* tcond, cudaStream_t stream) {
  if (nt <= 0) return 0;
  time_mlp_kernel<<<nt, 256, (dim + 1) * sizeof(float), stream>>>(times, fourier_w, w1, b1, dim, tcond);
  return check_launch("time_mlp");
}

extern "C" int e2b_time_gemv_launch(const float* tcond, int nt, int dim, const float* const* w, const float* const* b, const int* act,
                                    int nmat, float* out, cudaStream_t stream) {
  if (nt <= 0 || nmat <= 0) return 0;
  if (dim > GEMV_MAX_K * 32) { e2b_set_kernel_error("time_gemv: dim %d too large", dim); return -1; }
  dim3 grid((dim + 7) / 8, nmat);
  time_gemv_kernel<<<grid, 256, 0, stream>>>(tcond, nt, dim, w, b, act, nmat, out);
  return check_launch("time_gemv");
}

extern "C" int e2b_init_stream_split_launch(float* dst, void* dst_b16, int ldb, int b16_split, const float* registers, const float* src,
                                            int src_batches, const unsigned char* drop, const float* add_table, int batch, int n, int R, int C,
                                            cudaStream_t stream) {
  if (C % 4 || ldb % 4 || b16_split % 4) { e2b_set_kernel_error("init_stream: C / ldb / split must be multiples of 4"); return -1; }
  const int regs_only = (src_batches < 0);
  const size_t total = (size_t)batch * (regs_only ? R : R + n) * (C / 4);
  if (!total) return 0;
  ProfScope ps(stream, "init_stream", (long long)batch * (regs_only ? R : R + n), C, 0, 0.0, (double)total * 16.0 * (dst_b16 ? 1.5 : 1.0));
  init_stream_kernel<<<grid_for(total, 256), 256, 0, stream>>>(dst, reinterpret_cast<__nv_bfloat16*>(dst_b16), registers, src,
                                                               src_batches > 0 ? src_batches : 1, drop, add_table, batch, n, R, C,
                                                               regs_only, ldb > 0 ? ldb : C, b16_split);
  return check_launch("init_stream");
}

extern "C" int e2b_init_stream_launch(float* dst, void* dst_b16, const float* registers, const float* src, int src_batches,
                                      const unsigned char* drop, const float* add_table, int batch, int n, int R, int C,
                                      cudaStream_t stream) {
  return e2b_init_stream_split_launch(dst, dst_b16, C, 0, registers, src, src_batches, drop, add_table, batch, n, R, C, stream);
}

namespace e2b {
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  const size_t total = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % cols, r = i / cols;
    dst[(size_t)c * rows + r] = src[i];
  }
}
}  // namespace e2b

extern "C" int e2b_transpose_launch(const float* src, float* dst, int rows, int cols, cudaStream_t stream) {
  const size_t total = (size_t)rows * cols;
  if (!total) return 0;
  e2b::transpose_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, dst, rows, cols);
  return check_launch("transpose");
}

namespace e2b {
__global__ void __launch_bounds__(256) cast_part_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd,
                                                        size_t rows, int C, int W, int part) {
  const size_t total = rows * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = idx % W;
    const size_t r = idx / W;
    float v = c < C ? src[r * lds + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    dst[r * ldd + c] = part == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}
}  // namespace e2b

extern "C" int e2b_cast_part_launch(const float* src, int lds, void* dst, int ldd, int rows, int C, int W, int part, cudaStream_t stream) {
  const size_t total = (size_t)rows * W;
  if (!total) return 0;
  e2b::cast_part_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, (size_t)rows, C, W, part);
  return check_launch("cast_part");
}

namespace e2b {
__global__ void __launch_bounds__(256) scale_cols_kernel(const float* __restrict__ src, const float* __restrict__ gain, float* __restrict__ dst,
                                                         size_t total, int cols) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i] * gain[i % cols];
}
}  // namespace e2b

extern "C" int e2b_scale_cols_launch(const float* src, const float* gain, float* dst, int rows, int cols, cudaStream_t stream) {
  const size_t total = (size_t)rows * cols;
  if (!total) return 0;
  e2b::scale_cols_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, gain, dst, total, cols);
  return check_launch("scale_cols");
}

extern "C" int e2b_cast_pad_launch(const float* src, int lds, void* dst, int ldd, int rows, int C, cudaStream_t stream) {
  const size_t total = (size_t)rows * ldd;
  if (!total) return 0;
  cast_pad_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, (size_t)rows, C);
  return check_launch("cast_pad");
}

extern "C" int e2b_guided_euler_launch(float* y, const float* pred, int P, int B, long long per_sample, const float* w, float dt,
                                       int apg, float keep_parallel, double* scratch, void* y_b16, int n_copies, cudaStream_t stream) {
  return e2b_guided_euler_inpaint_launch(y, pred, P, B, per_sample, w, dt, apg, keep_parallel, scratch, y_b16, n_copies, nullptr, nullptr, 0,
                                         stream);
}

extern "C" int e2b_guided_euler_inpaint_launch(float* y, const float* pred, int P, int B, long long per_sample, const float* w, float dt,
                                               int apg, float keep_parallel, double* scratch, void* y_b16, int n_copies,
                                               const float* inpaint, const int* inpaint_lens, int row_elems, cudaStream_t stream) {
  if (P < 1 || P > EULER_MAX_P) { e2b_set_kernel_error("guided_euler: P=%d out of range", P); return -1; }
  if (inpaint && (!inpaint_lens || row_elems <= 0 || row_elems % 4 || per_sample % row_elems)) {
    e2b_set_kernel_error("guided_euler: in-painting needs lengths and a row size that divides the sample size");
    return -1;
  }
  if (per_sample % 4) { e2b_set_kernel_error("guided_euler: per-sample size must be a multiple of 4"); return -1; }
  if (apg && (P != 2 || !scratch)) { e2b_set_kernel_error("guided_euler: APG needs exactly one guidance pass and scratch"); return -1; }
  const size_t pass_stride = (size_t)B * per_sample;
  EulerW gw;
  for (int k = 0; k < EULER_MAX_P; ++k) gw.w[k] = (k < P - 1) ? w[k] : 0.f;
  if (apg) {
    if (cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * B, stream) != cudaSuccess) { e2b_set_kernel_error("guided_euler: memset of the APG scratch failed"); return -1; }
    dim3 g(grid_for(per_sample / 4, 256, 64), B);
    apg_reduce_kernel<<<g, 256, 0, stream>>>(pred, pass_stride, per_sample, scratch);
    if (check_launch("apg_reduce")) return -1;
  }
  const size_t total4 = pass_stride / 4;
  ProfScope ps(stream, "guided_euler", P, B, per_sample, 2.0 * P * pass_stride, 4.0 * pass_stride * (P + 2) + 2.0 * pass_stride * n_copies);
  // one float4 per thread (no grid-stride loop): a 20-50 us kernel lives on memory-level parallelism, and the capped grid left
  // each thread walking 2-3 dependent load -> store rounds (2.4 TB/s at C2; tools/bench_hbm_kernels.py)
  guided_euler_kernel<<<grid_for(total4, 256, 1 << 22), 256, 0, stream>>>(y, pred, P, pass_stride, per_sample, total4, gw, dt, apg,
                                                                keep_parallel, scratch, reinterpret_cast<__nv_bfloat16*>(y_b16), n_copies,
                                                                inpaint, inpaint_lens, row_elems / 4);
  return check_launch("guided_euler");
}

namespace e2b {
// dst[b, i, :] = i < lens[b] ? src[b, i, :] : 0      (where(cond_mask, cond, 0), e2_tts_crossatt3.py:2228)
__global__ void __launch_bounds__(256) mask_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, const int* __restrict__ lens,
                                                        int n, int c4, size_t total4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / c4;
    const int b = (int)(row / n), pos = (int)(row - (size_t)b * n);
    reinterpret_cast<float4*>(dst)[i] = pos < __ldg(lens + b) ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
}  // namespace e2b

extern "C" int e2b_mask_rows_launch(const float* src, float* dst, const int* lens_dev, int B, int n, int C, cudaStream_t stream) {
  if (C % 4) { e2b_set_kernel_error("mask_rows: C must be a multiple of 4"); return -1; }
  const size_t total4 = (size_t)B * n * (C / 4);
  if (!total4) return 0;
  e2b::mask_rows_kernel<<<grid_for(total4, 256), 256, 0, stream>>>(src, dst, lens_dev, n, C / 4, total4);
  return check_launch("mask_rows");
}
