// tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulation in TMEM.
//
//   * persistent, warp-specialised CTA (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
//     warp 2 = TMEM allocator, warps 4-7 = epilogue (one TMEM lane quarter each)
//   * operands staged by TMA (SWIZZLE_128B boxes of 64 K-elements) through a STAGES-deep mbarrier ring
//   * 128 x BN accumulator tile, double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
//     MMAs of tile i+1
//   * A may be the K-concatenation of up to three tensors (the reference's `pack((audio, text, frames))`
//     e2_tts_crossatt3.py:693-695 and `torch.cat((x, skip))` :1116) -- the concat tensor is never materialised
//   * fused epilogues: bias, GEGLU (x-transformers FeedForward glu=True), residual + AdaLN-Zero gate + row mask
//     (e2_tts_crossatt3.py:546-551, 1128-1137), and the QKV epilogue (interleaved RoPE on q/k, q pre-scaling,
//     transposed V store for the attention kernel, sigmoid value-head gate).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "ptx.cuh"

namespace e2b {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 256;

struct GemmArgs {
  CUtensorMap tmA[3];
  CUtensorMap tmB;
  int kb_end[3];
  e2b_gemm_desc d;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;   // 512 or 256: both powers of two
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&f)[32], int ncols) {
  if (ncols >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(f[8 * i + 0], f[8 * i + 1]);
      u.y = pack_bf16(f[8 * i + 2], f[8 * i + 3]);
      u.z = pack_bf16(f[8 * i + 4], f[8 * i + 5]);
      u.w = pack_bf16(f[8 * i + 6], f[8 * i + 7]);
      p[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[i] = __float2bfloat16_rn(f[i]);
  }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&f)[32], int ncols) {
  if (ncols >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[i] = f[i];
  }
}
__device__ __forceinline__ void load_f32x32(const float* src, float (&f)[32], int ncols) {
  if (ncols >= 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4* p = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = p[i];
      f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = (i < ncols) ? src[i] : 0.f;
  }
}

// One 32-column chunk of one accumulator row.  `col` = global (packed) column of f[0].
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const e2b_gemm_desc& d, int row, int col, float (&f)[32], float (&g)[32]) {
  const int ncols = min(32, d.N - col);
  if constexpr (EPI == E2B_EPI_BF16) {
    if (d.bias) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] += (i < ncols) ? __ldg(d.bias + col + i) : 0.f;
    }
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)row * d.ldo + col, f, ncols);
  } else if constexpr (EPI == E2B_EPI_F32) {
    const int rin = d.rpb_in > 0 ? row % d.rpb_in : row;
    const int orow = d.rpb_in > 0 ? (row / d.rpb_in) * d.rpb_out + d.row_off + rin : row;
    if (d.bias) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] += (i < ncols) ? __ldg(d.bias + col + i) : 0.f;
    }
    if (d.add_table) {
      float t[32];
      load_f32x32(d.add_table + (size_t)rin * d.ld_add + col, t, ncols);
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] += t[i];
    }
    store_f32x32(reinterpret_cast<float*>(d.out) + (size_t)orow * d.ldo + col, f, ncols);
    if (d.out_b16) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out_b16) + (size_t)orow * d.ldo_b16 + col, f, ncols);
  } else if constexpr (EPI == E2B_EPI_GEGLU) {
    // f = value columns, g = gate columns of the same packed tile; `col` is the OUTPUT column (inner index)
    const int nc = min(32, d.N / 2 - col);
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = f[i] * gelu_erf(g[i]);
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)row * d.ldo + col, f, nc);
  } else if constexpr (EPI == E2B_EPI_RESID) {
    const int b = d.rows_per_batch > 0 ? row / d.rows_per_batch : 0;
    const int pos = d.rows_per_batch > 0 ? row % d.rows_per_batch : row;
    const bool valid = d.lens ? (pos < __ldg(d.lens + b)) : true;
    float r[32];
    load_f32x32(d.resid + (size_t)row * d.ldr + col, r, ncols);
    if (valid) {
      if (d.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] += (i < ncols) ? __ldg(d.bias + col + i) : 0.f;
      }
      if (d.gate) {
        const float* gp = d.gate + (size_t)b * d.gate_bstride + col;
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] *= (i < ncols) ? __ldg(gp + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] += f[i];
    }
    store_f32x32(reinterpret_cast<float*>(d.out) + (size_t)row * d.ldo + col, r, ncols);
    if (d.out_b16) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out_b16) + (size_t)row * d.ldo_b16 + col, r, ncols);
  } else if constexpr (EPI == E2B_EPI_QKV) {
    const int b = row / d.rows_per_batch;
    const int pos = row % d.rows_per_batch;
    if (col < d.k_end) {
      // interleaved RoPE (x-transformers rotate_half on adjacent pairs): (x0,x1) -> (x0 c - x1 s, x1 c + x0 s)
      const float sc = (col < d.q_end) ? d.q_scale : 1.0f;
      const float2* rp = reinterpret_cast<const float2*>(d.rope) + (size_t)(pos + d.pos_off) * 32 + ((col & 63) >> 1);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 cs = __ldg(rp + i);
        const float x0 = f[2 * i], x1 = f[2 * i + 1];
        f[2 * i] = (x0 * cs.x - x1 * cs.y) * sc;
        f[2 * i + 1] = (x1 * cs.x + x0 * cs.y) * sc;
      }
      store_bf16x32(reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)row * d.ldo + col, f, 32);
    } else if (col < d.v_end) {
      const int c = col - d.k_end;
      const int h = c >> 6, d0 = c & 63;
      __nv_bfloat16* vp = reinterpret_cast<__nv_bfloat16*>(d.vt) + ((size_t)(b * d.heads_v + h) * 64 + d0) * d.vt_ld + pos;
#pragma unroll
      for (int i = 0; i < 32; ++i) vp[(size_t)i * d.vt_ld] = __float2bfloat16_rn(f[i]);
    } else {
      const int c = col - d.v_end;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < ncols) d.hgate[(size_t)row * d.hgate_ld + c + i] = sigmoidf_(f[i] + __ldg(d.hgate_bias + c + i));
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const e2b_gemm_desc& d = args.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (d.M + BM - 1) / BM;
  const int n_tiles = (d.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int KB = d.K / BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < d.num_src; ++s) tma_prefetch_desc(&args.tmA[s]);
    tma_prefetch_desc(&args.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      int src = 0, kb0 = 0;
      for (int kb = 0; kb < KB; ++kb) {
        while (kb >= args.kb_end[src]) { kb0 = args.kb_end[src]; ++src; }
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
        tma_load_2d(sA + stage * Cfg::A_BYTES, &args.tmA[src], &full[stage], (kb - kb0) * BK, m0);
        tma_load_2d(sB + stage * Cfg::B_BYTES, &args.tmB, &full[stage], kb * BK, n0);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
        const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16_ss(tmem_d, da + k * UMMA_K_STEP_ENC, db + k * UMMA_K_STEP_ENC, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (kb == KB - 1) umma_commit(&tfull[as]);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (TMEM -> registers -> global)
    const int ew = warp - 4;   // == warp % 4: the TMEM lane quarter this warp may read
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      const int row = m0 + ew * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * BN;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if constexpr (EPI == E2B_EPI_GEGLU) {
        static_assert(BN == 256 || EPI != E2B_EPI_GEGLU, "GEGLU packs 128 value + 128 gate columns per tile");
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          uint32_t v[32], g[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld32(taddr + BN / 2 + c * 32, g);
          tmem_ld_wait();
          float f[32], gg[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { f[i] = __uint_as_float(v[i]); gg[i] = __uint_as_float(g[i]); }
          // packed bias: value bias at packed col, gate bias at packed col + BN/2
          if (d.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              f[i] += __ldg(d.bias + n0 + c * 32 + i);
              gg[i] += __ldg(d.bias + n0 + BN / 2 + c * 32 + i);
            }
          }
          if (row < d.M) epilogue_chunk<EPI>(d, row, n0 / 2 + c * 32, f, gg);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          if (n0 + c * 32 >= d.N) break;
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (row < d.M) epilogue_chunk<EPI>(d, row, n0 + c * 32, f, f);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = {64 cols, box_rows}, 128B swizzle.
int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { e2b_set_kernel_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) {
    e2b_set_kernel_error("tensor map: base %p / ld %llu not 16-byte aligned", base, (unsigned long long)ld);
    return -1;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    e2b_set_kernel_error("cuTensorMapEncodeTiled failed: %d (rows %llu cols %llu ld %llu box_rows %u)", (int)r,
                         (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return -1;
  }
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int EPI>
static int launch_t(const GemmArgs& a, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { e2b_set_kernel_error("gemm smem attribute: %s", cudaGetErrorString(e)); return -1; }
    configured = true;
  }
  const int tiles = ((a.d.M + BM - 1) / BM) * ((a.d.N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("gemm launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // namespace e2b

using namespace e2b;

extern "C" void e2b_set_kernel_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* e2b_kernel_last_error(void) { return g_err; }

extern "C" int e2b_gemm_launch(const e2b_gemm_desc* d, cudaStream_t stream) {
  if (d->M <= 0 || d->N <= 0) return 0;
  if (d->num_src < 1 || d->num_src > 3) { e2b_set_kernel_error("gemm: num_src %d", d->num_src); return -1; }
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.d = *d;
  int k = 0;
  for (int s = 0; s < 3; ++s) {
    if (s < d->num_src) {
      if (d->ka[s] <= 0 || d->ka[s] % BK) { e2b_set_kernel_error("gemm: ka[%d]=%d must be a positive multiple of 64", s, d->ka[s]); return -1; }
      if (make_tmap_bf16(&a.tmA[s], d->a[s], d->M, d->ka[s], d->lda[s], BM)) return -1;
      k += d->ka[s];
    }
    a.kb_end[s] = (s < d->num_src) ? k / BK : (1 << 30);
  }
  if (k != d->K) { e2b_set_kernel_error("gemm: sum(ka)=%d != K=%d", k, d->K); return -1; }
  // Narrow outputs use 128-wide tiles; GEGLU needs the 128+128 packed 256 tile.
  const bool bn256 = (d->epi == E2B_EPI_GEGLU) || (d->N % 256 == 0) || (d->N > 1024);
  if (d->epi == E2B_EPI_GEGLU && d->N % 256) { e2b_set_kernel_error("gemm: GEGLU needs N %% 256 == 0 (N=%d)", d->N); return -1; }
  if (d->epi == E2B_EPI_QKV && ((d->q_end | d->k_end | d->v_end) % 64 || d->rows_per_batch <= 0)) {
    e2b_set_kernel_error("gemm: QKV segment ends must be multiples of 64 and rows_per_batch > 0");
    return -1;
  }
  if (make_tmap_bf16(&a.tmB, d->w, d->N, d->K, d->ldw, bn256 ? 256 : 128)) return -1;
#define E2B_DISPATCH(BN_)                                                             \
  switch (d->epi) {                                                                   \
    case E2B_EPI_BF16: return launch_t<BN_, E2B_EPI_BF16>(a, stream);                 \
    case E2B_EPI_F32: return launch_t<BN_, E2B_EPI_F32>(a, stream);                   \
    case E2B_EPI_RESID: return launch_t<BN_, E2B_EPI_RESID>(a, stream);               \
    case E2B_EPI_QKV: return launch_t<BN_, E2B_EPI_QKV>(a, stream);                   \
    default: break;                                                                   \
  }
  if (bn256) {
    if (d->epi == E2B_EPI_GEGLU) return launch_t<256, E2B_EPI_GEGLU>(a, stream);
    E2B_DISPATCH(256)
  } else {
    E2B_DISPATCH(128)
  }
#undef E2B_DISPATCH
  e2b_set_kernel_error("gemm: unknown epilogue %d", d->epi);
  return -1;
}
