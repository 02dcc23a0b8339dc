// tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulation in TMEM.
//
//   * persistent, warp-specialised CTA: warps 0..EW-1 = epilogue (EW = 4 or 8; one or two per TMEM lane quarter), then one TMA
//     producer warp, one MMA issuer warp and one TMEM allocator warp (the single-thread roles run warp-uniform, the instruction
//     under elect_one)
//   * operands staged by TMA (SWIZZLE_128B boxes of 64 K-elements) through a STAGES-deep mbarrier ring
//   * 128 x BN accumulator tile (BN = 128 or 256), double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
//     MMAs of tile i+1; CG = 2: a CTA pair shares a 256 x 256 tile (tcgen05.mma.cta_group::2, half of the B tile per CTA)
//   * A may be the K-concatenation of up to nine tensors (the reference's `pack((audio, text, frames))` e2_tts_crossatt3.py:693-695,
//     `torch.cat((x, skip))` :1116, `proj_in + cond_proj_in`, and the (hi, lo, hi) sources of the fp32 mode) -- never materialised
//   * a ragged last column tile runs an MMA only as wide as its columns, with its own B box
//   * fused epilogues: bias, GEGLU (x-transformers FeedForward glu=True), residual + AdaLN-Zero gate + row mask
//     (e2_tts_crossatt3.py:546-551, 1128-1137), QKV (interleaved RoPE on q/k, q pre-scaling, V rows, sigmoid value-head gate).
//     Classic form: transpose each 32 x 32 chunk through shared memory for coalesced global accesses.  Row-per-lane forms
//     (rt_epilogue, qt_epilogue): a lane keeps the accumulator row tcgen05.ld gives it, residual / results move by TMA, and the
//     RMSNorm of the value just produced becomes row sums + a scaled bf16 copy for the consuming GEMM (norm as a row scale).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "prof.h"
#include "ptx.cuh"

namespace e2b {

constexpr int BM = 128;
constexpr int BK = 64;

// Internal epilogue id: E2B_EPI_RESID with the residual stream moved by the TMA unit (see rt_epilogue below).
constexpr int EPI_RESID_TMA = 5;
// Internal epilogue id: E2B_EPI_QKV (bf16 mode, V as rows) with row-per-lane arithmetic and TMA stores (see qt_epilogue below).
constexpr int EPI_QKV_TMA = 6;

struct GemmArgs {
  CUtensorMap tmA[E2B_MAX_SRC];
  CUtensorMap tmB;
  CUtensorMap tmR, tmO, tmO16;   // EPI_RESID_TMA: fp32 residual in, fp32 result out, bf16 copy out (boxes of 32 rows x 32 columns)
                                 // EPI_QKV_TMA: tmO = q | k rows [M, k_end], tmO16 = v rows [M, v_end - k_end] (bf16, 32 x 32 boxes)
  CUtensorMap tmBt;          // ragged last column tile: the same W with a box of tail_rows rows (no zero-filled rows through the pipe)
  int tail_rows;             // 0: N is a multiple of the tile width (or the tail is a full box); else valid columns of the last tile, rounded up to 16
  int pf_kb;                 // A-operand L2 prefetch distance in 64-wide K blocks (0 = off)
  int kb_end[E2B_MAX_SRC];
  e2b_gemm_desc d;
};

// EW = number of epilogue warps.  4: one per TMEM lane quarter, 4-stage operand ring (large K, MMA-bound).  8: two per
// quarter taking alternate 32-column chunks, 3-stage ring: for K <= 1024 the epilogue (one warp per scheduler, IPC ~0.3)
// is longer than the 8K-cycle mainloop of a tile, so thread-level parallelism in the epilogue is worth a pipeline stage.
// CG = 2: a CTA PAIR (cluster of two CTAs on the SMs of one TPC) works on one 256 x BN tile with tcgen05.mma.cta_group::2: each CTA
// stages its own 128 rows of A and HALF of the B tile (BN / 2 rows of W), the pair's tensor cores read both halves -- a third less
// L2 -> shared-memory traffic per FLOP and two thirds of the shared-memory operand reads of two independent 128 x BN tiles.
template <int BN, int EW, int EPI = 0, int CG = 1>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CG == 2 ? ((EPI == EPI_RESID_TMA && EW == 8) ? 4 : (EW == 8 ? 5 : 6)) : ((BN == 256) ? (EW == 8 ? 3 : 4) : 6);
  static constexpr int THREADS = 128 + 32 * EW;
  // EPI_RESID_TMA: per epilogue warp RT_NBUF residual / result tiles of 32 x 32 fp32 (4 KB, SWIZZLE_128B) and one 32 x 32 bf16
  // tile (2 KB, SWIZZLE_64B); then the gate and bias vectors of the current column tile (2 x BN floats)
  static constexpr bool RT = (EPI == EPI_RESID_TMA);
  static constexpr int RT_NBUF = (EW == 8) ? 2 : 1;
  static constexpr int RT_WARP_BYTES = RT_NBUF * 4096 + 2048;
  // EPI_QKV_TMA: per epilogue warp two 32 x 32 bf16 tiles (2 KB each, SWIZZLE_64B) used alternately
  static constexpr bool QT = (EPI == EPI_QKV_TMA);
  static constexpr int EPI_BYTES = RT ? EW * RT_WARP_BYTES + 2 * BN * 4 : (QT ? EW * 4096 : EW * (4608 /*epilogue transpose buffer*/ + 128 /*row scales*/));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/ + (RT ? 128 : 0) /*residual-tile barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;   // 512 or 256: both powers of two
};

// exact-erf GELU (x-transformers uses nn.GELU(), not the tanh form).  erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7,
// three orders below the bf16 rounding of the result): 2 MUFU + ~10 FMA-pipe ops instead of the ~30-instruction erff, which
// made the GEGLU epilogue the limiter for small K.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);           // erf(|x| / sqrt 2)
  return 0.5f * x + 0.5f * fabsf(x) * e;                   // 0.5 x (1 + sign(x) erf(|x|/sqrt 2))
}
// The same for a pair in packed arithmetic (fma.rn.f32x2): the GEGLU epilogue is issue bound for small K (frames stream:
// ~4000 issue cycles per tile against 2048 of MMA), and the packed form needs ~9.5 instructions per element instead of ~17.
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x) {
  const uint64_t ax = x & 0x7fffffff7fffffffull;
  const uint64_t z = f32x2_mul(ax, f32x2_pack(0.70710678118654752f, 0.70710678118654752f));
  float d0, d1;
  f32x2_unpack(f32x2_fma(f32x2_pack(0.3275911f, 0.3275911f), z, f32x2_pack(1.0f, 1.0f)), d0, d1);
  const uint64_t t = f32x2_pack(rcp_approx(d0), rcp_approx(d1));
  uint64_t p = f32x2_fma(t, f32x2_pack(1.061405429f, 1.061405429f), f32x2_pack(-1.453152027f, -1.453152027f));
  p = f32x2_fma(t, p, f32x2_pack(1.421413741f, 1.421413741f));
  p = f32x2_fma(t, p, f32x2_pack(-0.284496736f, -0.284496736f));
  p = f32x2_fma(t, p, f32x2_pack(0.254829592f, 0.254829592f));
  float a0, a1;
  f32x2_unpack(f32x2_mul(f32x2_mul(z, z), f32x2_pack(-1.4426950408889634f, -1.4426950408889634f)), a0, a1);
  const uint64_t ex = f32x2_pack(ex2_approx(a0), ex2_approx(a1));                                    // exp(-z^2)
  const uint64_t e = f32x2_fma(f32x2_mul(p, t) ^ 0x8000000080000000ull, ex, f32x2_pack(1.0f, 1.0f));  // erf(|x| / sqrt 2)
  const uint64_t half = f32x2_pack(0.5f, 0.5f);
  return f32x2_fma(f32x2_mul(half, ax), e, f32x2_mul(half, x));          // 0.5 x (1 + sign(x) erf(|x|/sqrt 2))
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------
// Epilogue.  tcgen05.ld hands every thread one accumulator ROW (32 columns per chunk); writing that layout straight to
// global memory touches 32 different 128-byte lines per warp instruction (ncu: 32 sectors/request, L1-bound).  Each
// epilogue warp therefore transposes the 32x32 chunk through a private padded shared-memory buffer, after which lane
// l owns 4 consecutive columns (cg = l % 8) of rows (l / 8) + 4k, k = 0..7: a warp-wide 16-byte access covers
// 4 rows x 128 contiguous bytes.  The per-row loop is kept branch-free and address-light: row base pointers are computed
// once per tile, operand presence is tested once per chunk (ncu on the first "guarded" version: 15 branches and 17
// integer ops per store -- instruction bound).  All column counts must be multiples of 4 (checked on the host).
// ---------------------------------------------------------------------------------------------------------------
constexpr int EPI_PITCH4 = 9;                         // float4 per staged row (36 floats: conflict-free 16-byte accesses)
constexpr int EPI_BUF_BYTES = 32 * EPI_PITCH4 * 16;   // per epilogue warp

__device__ __forceinline__ uint2 pack4_bf16(const float4& v) { return make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w)); }
// bf16 store of 4 values; split > 0 also stores the rounding residual lo = bf16(v - hi) `split` elements further (the
// error-compensated fp32 mode: consumers read (hi, lo, hi) against weights (W_hi, W_hi, W_lo))
__device__ __forceinline__ void st_bf16x4(char* p, const float4& v, int split) {
  const uint2 hi = pack4_bf16(v);
  *reinterpret_cast<uint2*>(p) = hi;
  if (split > 0) {
    const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&hi.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&hi.y);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    *reinterpret_cast<uint2*>(p + 2 * split) = make_uint2(pack_bf16(v.x - f0.x, v.y - f0.y), pack_bf16(v.z - f1.x, v.w - f1.y));
  }
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_one() { return make_float4(1.f, 1.f, 1.f, 1.f); }

__device__ __forceinline__ void stage_rows(float4* buf, int lane, const uint32_t (&v)[32]) {
  float4* w = buf + lane * EPI_PITCH4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    w[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
}

// Row base pointers of the 8 rows a lane owns after the transpose (byte pointers; out == nullptr <=> row >= M).
struct EpiRows {
  char* out[8];          // primary output row (bf16 or fp32)
  char* out2[8];         // EPI_F32 / EPI_RESID: bf16 copy row (or nullptr) ; EPI_QKV: head-gate row (fp32)
  const char* aux[8];    // EPI_RESID: residual row ; EPI_F32: add_table row (or nullptr) ; EPI_QKV: rope row of this position
  const char* gate[8];   // EPI_RESID with a per-batch gate: gate row of this row's batch item
  bool valid[8];         // EPI_RESID row mask (valid length)
};


// ---------------------------------------------------------------------------------------------------------------
// EPI_RESID_TMA: out = resid + valid(row) * gate[col] * (acc + bias[col]) with the fp32 residual stream moved by the TMA unit.
// The classic epilogue above spends its time in the load/store unit: per 32 x 32 chunk a lane issues 8 LDG.128 for the residual,
// 16 STS/LDS for the transpose and 16 STG for the fp32 result and its bf16 copy, and stalls on the residual loads (ncu: LSU 62-68 %,
// long-scoreboard the top stall; the K <= 1024 out-projections sit at 3.8 TB/s of algorithmic traffic, 58 % of the HBM peak).
// Here a lane keeps the accumulator layout tcgen05.ld gives it -- one ROW of 32 columns -- and never touches global memory:
//   * the residual chunk [32 rows x 32 columns fp32] arrives by a TMA load (one elected lane, issued a chunk ahead) in the 128-byte
//     swizzled layout, in which the row-per-lane LDS.128 / STS.128 of a quarter warp hit 8 different 16-byte columns: conflict free;
//   * the lane updates its row in place in shared memory, writes the bf16 copy of the row into a second (64-byte swizzled) tile,
//     and one lane hands both tiles to TMA stores (full 128-byte lines, asynchronous, no registers, no L1 sectors);
//   * gate and bias of the tile's columns are staged once per tile in shared memory and read as broadcasts;
//   * because a lane owns whole rows, the row sums of squares that a following RMSNorm needs are lane-local: with d.row_ss set the
//     epilogue also writes sum(out^2) of the columns it handled (one partial per column tile and epilogue-warp group) and, with
//     d.b16_scale, the bf16 copy becomes bf16(out * scale[col]) -- the A operand of the next GEMM, which multiplies its accumulator
//     rows by sqrt(C) / ||x|| (norm as a row scale: the rmsnorm launch and its 6 B / element round trip disappear).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int crd0, int crd1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(crd0), "r"(crd1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory"); }

// How a CTA walks the output tiles: tile index t0, t0 + tstep, ... < total; tile -> first row (tile / n_tiles) * mrows + moff.
// One CTA per tile: mrows = 128, moff = 0; a CTA pair: both CTAs walk the pair's 256-row tiles, moff = 128 * rank.
struct TileWalk { int t0, tstep, total, n_tiles, mrows, moff; bool remote_tempty; };
__device__ __forceinline__ int tile_m0(const TileWalk& w, int tile) { return (tile / w.n_tiles) * w.mrows + w.moff; }
// epilogue warp done with a TMEM accumulator buffer: tell the MMA issuer (in a CTA pair it lives in the leader CTA)
__device__ __forceinline__ void arrive_tempty(const TileWalk& w, uint64_t* bar) {
  if (w.remote_tempty) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(smem_u32(bar) & 0xFEFFFFFFu) : "memory");
  else mbar_arrive(bar);
}

template <int BN, int EW, int CG>
__device__ __forceinline__ void rt_epilogue(const GemmArgs& args, uint8_t* sEpi, uint64_t* rfull_all, uint64_t* tfull, uint64_t* tempty,
                                            uint32_t tmem_base, int warp, int lane, const TileWalk tw) {
  using Cfg = GemmCfg<BN, EW, EPI_RESID_TMA, CG>;
  const int total = tw.total, n_tiles = tw.n_tiles;
  constexpr int NBUF = Cfg::RT_NBUF;
  constexpr int CSTEP = EW / 4;
  const e2b_gemm_desc& d = args.d;
  const int ew = warp & 3;      // TMEM lane quarter
  // With 8 epilogue warps the two warps of a lane quarter split the tile's 32-column chunks into two CONTIGUOUS runs (not alternate
  // chunks): a warp then owns whole 128-column groups, and the row sums of squares can be written per 128 columns -- the same
  // partials, summed in the same order, whatever the tile width, warp count or row count of the launch (bit-identical results for a
  // clip whatever the batch it is sampled in).
  constexpr int CPW = BN / 32 / CSTEP;                              // chunks per warp
  const int cw = warp >> 2;
  const int c_first = cw * CPW;
  uint8_t* rbuf = sEpi + warp * Cfg::RT_WARP_BYTES;                 // NBUF x 4 KB fp32 tiles (1024-byte aligned)
  uint8_t* hbuf = rbuf + NBUF * 4096;                               // 2 KB bf16 tile
  float* gvec = reinterpret_cast<float*>(sEpi + EW * Cfg::RT_WARP_BYTES);   // gate of the tile's columns
  float* bvec = gvec + BN;                                                   // bias
  uint64_t* rfull = rfull_all + warp * 2;
  const bool per_batch_gate = d.gate && d.gate_bstride != 0;
  const int sw7 = lane & 7, sw3 = (lane >> 1) & 3;                  // swizzle phases of this lane's row in the fp32 / bf16 tile

  // running chunk sequence of this warp over all its tiles: chunk q lives in buffer q % NBUF, barrier phase (q / NBUF) & 1
  int q_issue = 0, q_use = 0;
  int it_tile = tw.t0, it_c = c_first;          // cursor of the next residual chunk to request
  auto cursor_ok = [&]() { return it_tile < total; };
  // (it_tile, it_c) -> the next chunk this warp really has: column chunks past the width of a ragged last tile do not exist
  auto cursor_normalize = [&]() {
    while (it_tile < total) {
      if (it_c < c_first + CPW && (it_tile % n_tiles) * BN + it_c * 32 < d.N) break;
      it_c = c_first;
      it_tile += tw.tstep;
      if (it_tile < total && (it_tile % n_tiles) * BN + c_first * 32 >= d.N) it_c = c_first + CPW;   // not even the first chunk: skip the tile
    }
  };
  auto cursor_advance = [&]() {
    ++it_c;
    cursor_normalize();
  };
  auto issue_load = [&]() {                     // lane 0 only
    const int m0 = tile_m0(tw, it_tile), n0 = (it_tile % n_tiles) * BN;
    const int b = q_issue % NBUF;
    mbar_arrive_expect_tx(&rfull[b], 4096);
    tma_load_2d(rbuf + b * 4096, &args.tmR, &rfull[b], n0 + it_c * 32, m0 + ew * 32);
  };
  cursor_normalize();
  if (cursor_ok()) {
    if (lane == 0) issue_load();
    ++q_issue;
    cursor_advance();
  }

  int it = 0;
  for (int tile = tw.t0; tile < total; tile += tw.tstep, ++it) {
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    const int m0 = tile_m0(tw, tile), n0 = (tile % n_tiles) * BN;
    const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * BN;
    const int ncols = min(BN, d.N - n0);
    const int row = m0 + ew * 32 + lane;
    bool valid = row < d.M;
    const float* grow = nullptr;
    if (valid && d.rows_per_batch > 0) {
      const int b = row / d.rows_per_batch, pos = row - b * d.rows_per_batch;
      if (d.lens) valid = pos < __ldg(d.lens + b);
      if (per_batch_gate) grow = d.gate + (size_t)b * d.gate_bstride + n0;
    }
    // gate / bias of this tile's columns -> shared memory (all epilogue warps; the barriers keep tiles apart)
    named_bar_sync(1, 32 * EW);
    for (int i = warp * 32 + lane; i < BN; i += 32 * EW) {
      const bool in = i < ncols;
      gvec[i] = (in && d.gate && !per_batch_gate) ? __ldg(d.gate + n0 + i) : 1.0f;
      bvec[i] = (in && d.bias) ? __ldg(d.bias + n0 + i) : 0.0f;
    }
    named_bar_sync(1, 32 * EW);
    mbar_wait(&tfull[as], aphase);
    tc_fence_after();
    float ss = 0.f;                            // sum of squares of this row over the current 128-column group
#pragma unroll 1
    for (int c = c_first; c < c_first + CPW; ++c) {
      if (c * 32 >= ncols) break;
      const int b = q_use % NBUF;
      const uint32_t ph = (q_use / NBUF) & 1;
      // the stores of the previous chunk must have finished READING their tiles before the bf16 tile / the other fp32 tile is reused
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      if constexpr (NBUF == 2) {                // request the next chunk's residual now: a whole chunk of latency cover
        if (cursor_ok()) {
          if (lane == 0) issue_load();
          ++q_issue;
          cursor_advance();
        }
      }
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      mbar_wait(&rfull[b], ph);
      tmem_ld_wait();
      uint8_t* rrow = rbuf + b * 4096 + lane * 128;
      uint8_t* hrow = hbuf + lane * 64;
      const float4* g4 = reinterpret_cast<const float4*>(gvec + c * 32);
      const float4* b4 = reinterpret_cast<const float4*>(bvec + c * 32);
      float sc = 0.f;                           // this chunk's 32 columns, in column order
      const float4* s4 = d.b16_scale ? reinterpret_cast<const float4*>((row >= d.b16_split_row && d.b16_scale2 ? d.b16_scale2 : d.b16_scale) + n0 + c * 32) : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4* rp = reinterpret_cast<float4*>(rrow + ((j ^ sw7) * 16));
        float4 r = *rp;
        float4 gg = g4[j];
        if (grow) gg = __ldg(reinterpret_cast<const float4*>(grow + c * 32) + j);
        const float4 bb = b4[j];
        if (valid) {
          r.x = fmaf(__uint_as_float(v[4 * j + 0]) + bb.x, gg.x, r.x);
          r.y = fmaf(__uint_as_float(v[4 * j + 1]) + bb.y, gg.y, r.y);
          r.z = fmaf(__uint_as_float(v[4 * j + 2]) + bb.z, gg.z, r.z);
          r.w = fmaf(__uint_as_float(v[4 * j + 3]) + bb.w, gg.w, r.w);
        }
        *rp = r;
        // (columns past N in a ragged last chunk hold stale accumulator columns -- the narrow MMA never wrote them; the TMA store
        // clips them, the row sums must skip them: N is a multiple of 4, so a float4 is live or dead as a whole)
        if (c * 32 + 4 * j < ncols) sc = fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, fmaf(r.w, r.w, sc))));
        if (d.out_b16) {
          float4 o = r;
          if (s4) { const float4 sc = __ldg(s4 + j); o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w; }
          // 64-byte swizzle: 16-byte column (j / 2) of the row XOR ((row / 2) % 4); 8 bytes per j
          *reinterpret_cast<uint2*>(hrow + (((j >> 1) ^ sw3) * 16) + (j & 1) * 8) = pack4_bf16(o);
        }
      }
      fence_proxy_async_smem();                 // this lane's tile writes are visible to the TMA unit
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&args.tmO, rbuf + b * 4096, n0 + c * 32, m0 + ew * 32);
        if (d.out_b16) tma_store_2d(&args.tmO16, hbuf, n0 + c * 32, m0 + ew * 32);
        bulk_commit();
      }
      // canonical partials: ((chunk 0 + chunk 1) + chunk 2) + chunk 3 of every 128-column group
      ss += sc;
      if ((c & 3) == 3 || (c + 1) * 32 >= ncols) {
        if (d.row_ss && row < d.M) d.row_ss[(size_t)((n0 >> 7) + (c >> 2)) * d.row_ss_ld + row] = ss;
        ss = 0.f;
      }
      ++q_use;
      if constexpr (NBUF == 1) {                // single tile: the next residual can only follow the store's read of this one
        if (cursor_ok()) {
          if (lane == 0) { bulk_wait_read0(); issue_load(); }
          ++q_issue;
          cursor_advance();
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive_tempty(tw, &tempty[as]);
  }
  if (lane == 0) bulk_wait0();                  // every store of this lane has landed before the CTA may exit
  __syncwarp();
}

// One interleaved-RoPE pair (x-transformers rotate_half on adjacent columns) with the q pre-scale / norm row scale folded in.  Spelled
// with explicit roundings so that the classic and the TMA-store QKV epilogues give the same bits (a clip's result must not depend
// on which tile configuration its batch size selects).
__device__ __forceinline__ void rope_pair(float x0, float x1, float c, float s, float sc, float& o0, float& o1) {
  o0 = __fmul_rn(fmaf(x0, c, -__fmul_rn(x1, s)), sc);
  o1 = __fmul_rn(fmaf(x1, c, __fmul_rn(x0, s)), sc);
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// EPI_QKV_TMA: the QKV epilogue without the shared-memory transpose and without global stores from the lanes.  The classic QKV
// epilogue issues per 32 x 32 chunk and lane 8 LDG.128 (rope), 8 STS.128 + 8 LDS.128 (transpose) and 8 STG.64 that cover only 64
// contiguous bytes per row; with K <= 1280 that is longer than the tile's MMAs (QKV audio 1150 TF/s against 1490 for the GEGLU GEMM
// of the same K).  Here a lane keeps the accumulator ROW tcgen05.ld gives it:
//   * the rope (cos, sin) values of a row depend on the column only through (column % 64), and a warp that takes every second
//     32-column chunk always sees the same half of the head: 32 values per lane, loaded ONCE per tile (before the accumulator is
//     waited for) and kept in registers;
//   * rotation, q pre-scale and the norm row scale are lane-local; the bf16 row goes into a 64-byte-swizzled 32 x 32 tile (two tiles
//     per warp, used alternately) and leaves by a TMA store into the q | k buffer or the V rows; rows past M are clipped by the unit;
//   * the head-gate columns (the ragged last tile) are stored directly: 16 floats per row.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int EW, int CG>
__device__ __forceinline__ void qt_epilogue(const GemmArgs& args, uint8_t* sEpi, uint64_t* tfull, uint64_t* tempty, uint32_t tmem_base,
                                            int warp, int lane, const TileWalk tw) {
  const e2b_gemm_desc& d = args.d;
  const int ew = warp & 3;      // TMEM lane quarter
  constexpr int HSTEP = EW / 4; // 8 warps: the two warps of a quarter take the even / the odd chunks; 4 warps: even chunks, then odd ones
  uint8_t* hbuf = sEpi + warp * 4096;
  const int sw3 = (lane >> 1) & 3;
  int q = 0;                    // running chunk count of this warp: tile buffer q & 1
  int it = 0;
  for (int tile = tw.t0; tile < tw.total; tile += tw.tstep, ++it) {
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    const int m0 = tile_m0(tw, tile), n0 = (tile % tw.n_tiles) * BN;
    const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * BN;
    const int row = m0 + ew * 32 + lane;
    const bool live = row < d.M;
    const float* rope_row = nullptr;
    float rs = 1.0f;
    if (live) {
      const int b = row / d.rows_per_batch, pos = row - b * d.rows_per_batch;
      rope_row = d.rope + (size_t)(pos + d.pos_off) * 64;
      if (d.in_row_ss) {      // RMSNorm of the A operand's rows as a scale of the accumulator rows (partials summed in a fixed order)
        float s2 = 0.f;
        for (int p = 0; p < d.in_row_parts; ++p) s2 += __ldg(d.in_row_ss + (size_t)p * d.in_row_ss_ld + row);
        rs = d.in_row_mult / fmaxf(sqrtf(s2), 1e-12f);
      }
    }
    bool waited = false;
#pragma unroll 1
    for (int h = (HSTEP == 2 ? (warp >> 2) : 0); h < 2; h += HSTEP) {
      float4 rr[8];             // (cos, sin) of the 16 column pairs of this half of the head, for this lane's position
      if (n0 + h * 32 < d.k_end) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rr[j] = live ? __ldg(reinterpret_cast<const float4*>(rope_row + h * 32) + j) : f4_zero();
      }
      if (!waited) {
        mbar_wait(&tfull[as], aphase);
        tc_fence_after();
        waited = true;
      }
#pragma unroll 1
      for (int c = h; c < BN / 32; c += 2) {
        const int col0 = n0 + c * 32;
        if (col0 >= d.N) break;
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        if (col0 >= d.v_end) {                     // head-gate columns: sigmoid(acc * rs + bias), fp32, straight from the lane
          tmem_ld_wait();
          if (live) {
            float* gp = d.hgate + (size_t)row * d.hgate_ld + (col0 - d.v_end);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col0 + 4 * j < d.N) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(d.hgate_bias + (col0 - d.v_end)) + j);
                float4 o;
                o.x = sigmoidf_(__fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 0]), rs), bb.x));
                o.y = sigmoidf_(__fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 1]), rs), bb.y));
                o.z = sigmoidf_(__fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 2]), rs), bb.z));
                o.w = sigmoidf_(__fadd_rn(__fmul_rn(__uint_as_float(v[4 * j + 3]), rs), bb.w));
                reinterpret_cast<float4*>(gp)[j] = o;
              }
            }
          }
          continue;
        }
        // the store issued two chunks ago has finished reading the tile this chunk is written into
        if (lane == 0) bulk_wait_read1();
        __syncwarp();
        uint8_t* hrow = hbuf + (q & 1) * 2048 + lane * 64;
        tmem_ld_wait();
        if (col0 < d.k_end) {
          const float sc = (col0 < d.q_end) ? __fmul_rn(d.q_scale, rs) : rs;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            rope_pair(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]), rr[j].x, rr[j].y, sc, o.x, o.y);
            rope_pair(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]), rr[j].z, rr[j].w, sc, o.z, o.w);
            *reinterpret_cast<uint2*>(hrow + (((j >> 1) ^ sw3) * 16) + (j & 1) * 8) = pack4_bf16(o);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 o = make_float4(__uint_as_float(v[4 * j + 0]) * rs, __uint_as_float(v[4 * j + 1]) * rs,
                                         __uint_as_float(v[4 * j + 2]) * rs, __uint_as_float(v[4 * j + 3]) * rs);
            *reinterpret_cast<uint2*>(hrow + (((j >> 1) ^ sw3) * 16) + (j & 1) * 8) = pack4_bf16(o);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (col0 < d.k_end) tma_store_2d(&args.tmO, hbuf + (q & 1) * 2048, col0, m0 + ew * 32);
          else tma_store_2d(&args.tmO16, hbuf + (q & 1) * 2048, col0 - d.k_end, m0 + ew * 32);
          bulk_commit();
        }
        ++q;
      }
    }
    if (!waited) {              // (a warp without chunks in a ragged tile still takes part in the accumulator hand-over)
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive_tempty(tw, &tempty[as]);
  }
  if (lane == 0) bulk_wait0();
  __syncwarp();
}

template <int BN, int EPI, int EW, int CG = 1>
__global__ void __launch_bounds__(128 + 32 * EW, 1) gemm_kernel(const __grid_constant__ GemmArgs args) {
  using Cfg = GemmCfg<BN, EW, EPI, CG>;
  // Used directly (no integer round-trip) so the compiler keeps the shared address space: the earlier manual 1024-byte
  // round-up through uintptr_t turned every access into generic LD.E/ST.E.  SWIZZLE_128B needs a 1024-byte aligned base;
  // with no static shared memory the dynamic window starts at offset 0 -- checked once below.
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
    printf("e2b: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint8_t* sEpi = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + Cfg::EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = bars + 2 * Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* rfull = bars + 32;                  // EPI_RESID_TMA: [EW][2] residual-tile barriers (byte offset 256)

  const e2b_gemm_desc& d = args.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (d.M + BM - 1) / BM;
  const int n_tiles = (d.N + BN - 1) / BN;
  const int KB = d.K / BK;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;     // CTA pair: 0 = leader (issues the MMAs), 1 = follower
  TileWalk tw;
  tw.n_tiles = n_tiles;
  tw.t0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  tw.tstep = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  tw.total = CG == 2 ? ((m_tiles + 1) / 2) * n_tiles : m_tiles * n_tiles;
  tw.mrows = BM * CG;
  tw.moff = (int)rank * BM;
  tw.remote_tempty = CG == 2 && rank != 0;
  const int total = tw.total;

  // Warp roles: 0..EW-1 epilogue, EW = TMA producer, EW+1 = MMA issuer, EW+2 = TMEM allocator.  The single-thread issuers
  // take the highest warp ids because the scheduler arbitrates highest-warp-id-first: as warps 0/1 they were starved by the
  // busy epilogue warps sharing their schedulers.
  constexpr int W_TMA = EW, W_MMA = EW + 1, W_ALLOC = EW + 2;
  if (warp == W_TMA && lane == 0) {
    for (int s = 0; s < d.num_src; ++s) tma_prefetch_desc(&args.tmA[s]);
    tma_prefetch_desc(&args.tmB);
    if (args.tail_rows > 0) tma_prefetch_desc(&args.tmBt);
    if constexpr (EPI == EPI_RESID_TMA) {
      tma_prefetch_desc(&args.tmR);
      tma_prefetch_desc(&args.tmO);
      if (d.out_b16) tma_prefetch_desc(&args.tmO16);
    }
    if constexpr (EPI == EPI_QKV_TMA) {
      tma_prefetch_desc(&args.tmO);
      if (d.v_end > d.k_end) tma_prefetch_desc(&args.tmO16);
    }
  }
  if (warp == W_MMA && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], EW * CG);       // CTA pair: the epilogue warps of BOTH CTAs release the leader's accumulator buffer
    }
    if constexpr (EPI == EPI_RESID_TMA)
      for (int s = 0; s < 2 * EW; ++s) mbar_init(&rfull[s], 1);
    fence_mbar_init();
  }
  if (warp == W_ALLOC) {
    if constexpr (CG == 2) { tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync();          // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The producer and MMA roles are executed by their WHOLE warp with warp-uniform control flow; only the TMA / MMA / commit
  // instructions sit under elect_one().  (With `lane == 0` role branches every UTCHMMA / UTMALDG was wrapped by the compiler in
  // a divergence loop -- R2UR moves, ELECT, predicate shuffles and a backward BRA.U.ANY -- costing ~100 cycles per MMA issue.)
  if (warp == W_TMA) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    // Optional A-operand prefetch cursor (OFF by default, E2B_GEMM_PF=<K blocks> to try it): the CTA that owns column tile 0 of a
    // row block asks the TMA unit to pull the A block into L2 pf_kb K blocks ahead of its own loads.  The idea was that A (0.2-1 GB,
    // written by the previous kernel) is cold in L2 with only STAGES - 1 K blocks in flight.  MEASURED: a clear loss on every shape
    // (tools/bench_gemm_pf.py, profiles/r02_gemm_prefetch_ab.txt: GEGLU text 1385 -> 1197 TF/s, FF2 audio 1125 -> 950, out-projections
    // unchanged) and 1.6-1.9x the algorithmic DRAM reads under ncu -- the operand ring already covers the latency, the extra requests
    // only compete with the real loads.  Kept as a switch so the result can be reproduced.
    int pf_tile = tw.t0, pf_k = 0, pf_src = 0, pf_k0 = 0;
    auto pf_step = [&]() {
      if (pf_tile >= total) return;
      if (pf_tile % n_tiles == 0 && elect_one())
        tma_prefetch_l2_2d(&args.tmA[pf_src], (pf_k - pf_k0) * BK, tile_m0(tw, pf_tile));
      if (++pf_k == KB) { pf_k = 0; pf_src = 0; pf_k0 = 0; pf_tile += tw.tstep; }
      else while (pf_k >= args.kb_end[pf_src]) { pf_k0 = args.kb_end[pf_src]; ++pf_src; }
    };
    if (args.pf_kb > 0)
      for (int i = 0; i < args.pf_kb; ++i) pf_step();
    for (int tile = tw.t0; tile < total; tile += tw.tstep) {
      const int n_tile = tile % n_tiles;
      const int m0 = tile_m0(tw, tile), n0 = n_tile * BN;
      const bool tail = args.tail_rows > 0 && n_tile == n_tiles - 1;
      const CUtensorMap* tmb = tail ? &args.tmBt : &args.tmB;
      // bytes this CTA stages per K block; rows of W it owns in the tile (a CTA pair splits the tile's columns in halves)
      const int brows = tail ? args.tail_rows / CG : BN / CG;
      const uint32_t bytes = (uint32_t)(Cfg::A_BYTES + brows * BK * 2);
      int src = 0, kb0 = 0;
      for (int kb = 0; kb < KB; ++kb) {
        while (kb >= args.kb_end[src]) { kb0 = args.kb_end[src]; ++src; }
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          if constexpr (CG == 2) {
            // both CTAs load their halves and signal the LEADER's barrier, which expects the bytes of the pair
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * bytes);
            tma_load_2d_cg2(sA + stage * Cfg::A_BYTES, &args.tmA[src], &full[stage], (kb - kb0) * BK, m0);
            tma_load_2d_cg2(sB + stage * Cfg::B_BYTES, tmb, &full[stage], kb * BK, n0 + (int)rank * brows);
          } else {
            mbar_arrive_expect_tx(&full[stage], bytes);
            tma_load_2d(sA + stage * Cfg::A_BYTES, &args.tmA[src], &full[stage], (kb - kb0) * BK, m0);
            tma_load_2d(sB + stage * Cfg::B_BYTES, tmb, &full[stage], kb * BK, n0);
          }
        }
        __syncwarp();
        if (args.pf_kb > 0) pf_step();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_MMA && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (CTA pair: the leader issues for both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tw.t0; tile < total; tile += tw.tstep, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      // A ragged last column tile (QKV + head-gate: N = 3088 = 12 x 256 + 16) only multiplies the columns it has, rounded up to the
      // MMA's N granularity of 16: a 16-column tile costs 1/16 of the tensor time instead of a full tile (8 % of the QKV GEMMs)
      const int nv = min(BN, d.N - (tile % n_tiles) * BN);
      const uint32_t idesc = umma_idesc_bf16(BM * CG, (uint32_t)((nv + 16 * CG - 1) & ~(16 * CG - 1)));   // 16 columns per CTA
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
          if constexpr (CG == 2) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss_cg2(tmem_d, da + k * UMMA_K_STEP_ENC, db + k * UMMA_K_STEP_ENC, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_cg2(&empty[stage]);         // frees the slot in both CTAs
            if (kb == KB - 1) umma_commit_cg2(&tfull[as]);
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss(tmem_d, da + k * UMMA_K_STEP_ENC, db + k * UMMA_K_STEP_ENC, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (kb == KB - 1) umma_commit(&tfull[as]);
          }
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < EW) {
   if constexpr (EPI == EPI_RESID_TMA) {
    rt_epilogue<BN, EW, CG>(args, sEpi, rfull, tfull, tempty, tmem_base, warp, lane, tw);
   } else if constexpr (EPI == EPI_QKV_TMA) {
    qt_epilogue<BN, EW, CG>(args, sEpi, tfull, tempty, tmem_base, warp, lane, tw);
   } else {
    // ------------------------------------------------------------ epilogue (TMEM -> regs -> smem transpose -> global)
    const int ewi = warp;
    const int ew = ewi & 3;    // == warp % 4: the TMEM lane quarter this warp may read
    const int cw = ewi >> 2;   // with 8 epilogue warps: which alternate 32-column chunks this warp takes
    constexpr int CSTEP = EW / 4;
    float4* buf = reinterpret_cast<float4*>(sEpi + ewi * EPI_BUF_BYTES);
    float* rsv = reinterpret_cast<float*>(sEpi + EW * EPI_BUF_BYTES) + ewi * 32;    // row scale of this warp's 32 rows (norm as a row scale)
    const bool rowscale = (EPI == E2B_EPI_GEGLU || EPI == E2B_EPI_QKV) && d.in_row_ss != nullptr;
    const int rsub = lane >> 3, cg = lane & 7;
    const float4* bufr = buf + rsub * EPI_PITCH4 + cg;          // + 4k * EPI_PITCH4 selects row 4k + rsub
    const bool per_batch_gate = (EPI == E2B_EPI_RESID) && d.gate && d.gate_bstride != 0;
    int it = 0;
    for (int tile = tw.t0; tile < total; tile += tw.tstep, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = tile_m0(tw, tile), n0 = (tile % n_tiles) * BN;
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * BN;
      const int ncols = min(BN, d.N - n0);                      // valid packed columns of this tile (multiple of 4)
      EpiRows R;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = m0 + ew * 32 + 4 * k + rsub;
        const bool ok = row < d.M;
        R.out[k] = nullptr; R.out2[k] = nullptr; R.aux[k] = nullptr; R.gate[k] = nullptr; R.valid[k] = true;
        if (ok) {
          if constexpr (EPI == E2B_EPI_BF16) {
            R.out[k] = reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)row * d.ldo + n0);
          } else if constexpr (EPI == E2B_EPI_GEGLU) {
            R.out[k] = reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)row * d.ldo + n0 / 2);
          } else if constexpr (EPI == E2B_EPI_F32) {
            int orow = row, pos = row;
            if (d.rpb_in > 0) { const int b = row / d.rpb_in; pos = row - b * d.rpb_in; orow = b * d.rpb_out + d.row_off + pos; }
            R.out[k] = reinterpret_cast<char*>(reinterpret_cast<float*>(d.out) + (size_t)orow * d.ldo + n0);
            if (d.out_b16) R.out2[k] = reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(d.out_b16) + (size_t)orow * d.ldo_b16 + n0);
            if (d.add_table) R.aux[k] = reinterpret_cast<const char*>(d.add_table + (size_t)pos * d.ld_add + n0);
          } else if constexpr (EPI == E2B_EPI_RESID) {
            R.out[k] = reinterpret_cast<char*>(reinterpret_cast<float*>(d.out) + (size_t)row * d.ldo + n0);
            R.aux[k] = reinterpret_cast<const char*>(d.resid + (size_t)row * d.ldr + n0);
            if (d.out_b16) R.out2[k] = reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(d.out_b16) + (size_t)row * d.ldo_b16 + n0);
            if (d.rows_per_batch > 0) {
              const int b = row / d.rows_per_batch, pos = row - b * d.rows_per_batch;
              if (d.lens) R.valid[k] = pos < __ldg(d.lens + b);
              if (per_batch_gate) R.gate[k] = reinterpret_cast<const char*>(d.gate + (size_t)b * d.gate_bstride + n0);
            }
          } else if constexpr (EPI == E2B_EPI_QKV) {
            const int b = row / d.rows_per_batch, pos = row - b * d.rows_per_batch;
            R.out[k] = reinterpret_cast<char*>(reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)row * d.ldo);
            R.aux[k] = reinterpret_cast<const char*>(d.rope + (size_t)(pos + d.pos_off) * 64);
            R.out2[k] = reinterpret_cast<char*>(d.hgate + (size_t)row * d.hgate_ld);
          }
        }
      }
      if constexpr (EPI == E2B_EPI_RESID) {
        // Pull the residual block of this CTA's NEXT tile towards L2 now, a whole epilogue ahead: the epilogue's residual loads
        // were DRAM misses with only ~32 KB in flight per SM (tools/bench_gemm3.py: out-projections 6-10 % faster with it).
        const int nt = tile + tw.tstep;
        if (nt < total) {
          const int pm = tile_m0(tw, nt) + ew * 32 + lane, pn = (nt % n_tiles) * BN;
          if (pm < d.M) {
            const float* src = d.resid + (size_t)pm * d.ldr + pn;
#pragma unroll
            for (int c = cw; c < BN / 32; c += CSTEP)
              if (pn + c * 32 < d.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + c * 32));
          }
        }
      }
      if (rowscale) {
        // RMSNorm of the A operand's rows as a scale of the accumulator rows: sqrt(C) / max(||x||, 1e-12), ||x||^2 from the partial
        // sums the producing epilogue left (summed in a fixed order: deterministic)
        const int row = m0 + ew * 32 + lane;
        float s2 = 0.f;
        if (row < d.M)
          for (int p = 0; p < d.in_row_parts; ++p) s2 += __ldg(d.in_row_ss + (size_t)p * d.in_row_ss_ld + row);
        __syncwarp();
        rsv[lane] = d.in_row_mult / fmaxf(sqrtf(s2), 1e-12f);
        __syncwarp();
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();

      if constexpr (EPI == E2B_EPI_GEGLU) {
        static_assert(BN == 256 || EPI != E2B_EPI_GEGLU, "GEGLU packs 128 value + 128 gate columns per tile");
#pragma unroll 1
        for (int c = cw; c < BN / 64; c += CSTEP) {
          uint32_t v[32];
          float4 gt[8];
          tmem_ld32(taddr + BN / 2 + c * 32, v);     // gate half first
          tmem_ld_wait();
          stage_rows(buf, lane, v);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) gt[k] = bufr[4 * k * EPI_PITCH4];
          __syncwarp();
          tmem_ld32(taddr + c * 32, v);              // value half
          tmem_ld_wait();
          stage_rows(buf, lane, v);
          __syncwarp();
          const int pc = n0 + c * 32 + cg * 4;       // packed column of the value element; its gate is at +BN/2
          float4 bv = f4_zero(), bg = f4_zero();
          if (d.bias) {
            bv = __ldg(reinterpret_cast<const float4*>(d.bias + pc));
            bg = __ldg(reinterpret_cast<const float4*>(d.bias + pc + BN / 2));
          }
          const int ob = (c * 32 + cg * 4) * 2;      // byte offset inside the tile's bf16 output row
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 a = bufr[4 * k * EPI_PITCH4];
            const float rs = rowscale ? rsv[4 * k + rsub] : 1.0f;
            const uint64_t one2 = f32x2_pack(rs, rs);            // the row scale rides on the bias fma
            const uint64_t g01 = gelu_erf2(f32x2_fma(f32x2_pack(gt[k].x, gt[k].y), one2, f32x2_pack(bg.x, bg.y)));
            const uint64_t g23 = gelu_erf2(f32x2_fma(f32x2_pack(gt[k].z, gt[k].w), one2, f32x2_pack(bg.z, bg.w)));
            const uint64_t o01 = f32x2_mul(f32x2_fma(f32x2_pack(a.x, a.y), one2, f32x2_pack(bv.x, bv.y)), g01);
            const uint64_t o23 = f32x2_mul(f32x2_fma(f32x2_pack(a.z, a.w), one2, f32x2_pack(bv.z, bv.w)), g23);
            float4 o;
            f32x2_unpack(o01, o.x, o.y);
            f32x2_unpack(o23, o.z, o.w);
            if (R.out[k]) st_bf16x4(R.out[k] + ob, o, d.split);
          }
          __syncwarp();
        }
      } else if constexpr (EPI == E2B_EPI_QKV) {
#pragma unroll 1
        for (int c = cw; c < BN / 32; c += CSTEP) {
          const int col0 = n0 + c * 32;
          if (col0 >= d.N) break;
          const int col = col0 + cg * 4;
          float4 cs[8];
          if (col0 < d.k_end) {                      // rope (cos, sin) pairs of this lane's two column pairs, all 8 rows
#pragma unroll
            for (int k = 0; k < 8; ++k)
              cs[k] = R.aux[k] ? __ldg(reinterpret_cast<const float4*>(R.aux[k] + ((col & 63) >> 1) * 8)) : f4_zero();
          }
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (col0 >= d.k_end && col0 < d.v_end && d.v_f32 == nullptr && !d.v_rowmajor) {
            // V^T store: thread = key position, so the 32 lanes of a store are 32 consecutive keys (64 contiguous bytes)
            const int row = m0 + ew * 32 + lane;
            if (row < d.M) {
              const int b = row / d.rows_per_batch, pos = row - b * d.rows_per_batch;
              const int cc = col0 - d.k_end;
              __nv_bfloat16* vp = reinterpret_cast<__nv_bfloat16*>(d.vt) + ((size_t)(b * d.heads_v + (cc >> 6)) * 64 + (cc & 63)) * d.vt_ld + pos;
              const float rs = rowscale ? rsv[lane] : 1.0f;
#pragma unroll
              for (int i = 0; i < 32; ++i) vp[(size_t)i * d.vt_ld] = __float2bfloat16_rn(__uint_as_float(v[i]) * rs);
            }
            continue;
          }
          stage_rows(buf, lane, v);
          __syncwarp();
          if (col0 < d.k_end) {
            // interleaved RoPE (x-transformers rotate_half on adjacent pairs): (x0,x1) -> (x0 c - x1 s, x1 c + x0 s)
            const float sc0 = (col0 < d.q_end) ? d.q_scale : 1.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 a = bufr[4 * k * EPI_PITCH4];
              const float sc = rowscale ? __fmul_rn(sc0, rsv[4 * k + rsub]) : sc0;
              float4 o;
              rope_pair(a.x, a.y, cs[k].x, cs[k].y, sc, o.x, o.y);
              rope_pair(a.z, a.w, cs[k].z, cs[k].w, sc, o.z, o.w);
              if (R.out[k]) {
                if (d.qk_f32) *reinterpret_cast<float4*>(d.qk_f32 + (size_t)(m0 + ew * 32 + 4 * k + rsub) * d.ldo + col) = o;
                else *reinterpret_cast<uint2*>(R.out[k] + col * 2) = pack4_bf16(o);
              }
            }
          } else if (col0 < d.v_end && d.v_f32 == nullptr) {   // v as plain bf16 rows [M, vt_ld] (MN-major operand of the attention kernel)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (R.out[k]) {
                float4 a = bufr[4 * k * EPI_PITCH4];
                if (rowscale) { const float rs = rsv[4 * k + rsub]; a.x *= rs; a.y *= rs; a.z *= rs; a.w *= rs; }
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.vt) + (size_t)(m0 + ew * 32 + 4 * k + rsub) * d.vt_ld + (col - d.k_end)) =
                    pack4_bf16(a);
              }
          } else if (col0 < d.v_end) {               // fp32 mode: v as plain fp32 [M, v_f32_ld]
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (R.out[k]) *reinterpret_cast<float4*>(d.v_f32 + (size_t)(m0 + ew * 32 + 4 * k + rsub) * d.v_f32_ld + (col - d.k_end)) = bufr[4 * k * EPI_PITCH4];
          } else if (col < d.N) {                    // head-gate columns [v_end, N): few, scalar stores
            const int gc = col - d.v_end;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 a0 = bufr[4 * k * EPI_PITCH4];
              const float rs = rowscale ? rsv[4 * k + rsub] : 1.0f;
              const float e[4] = {__fmul_rn(a0.x, rs), __fmul_rn(a0.y, rs), __fmul_rn(a0.z, rs), __fmul_rn(a0.w, rs)};
              if (R.out2[k]) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (col + i < d.N) reinterpret_cast<float*>(R.out2[k])[gc + i] = sigmoidf_(__fadd_rn(e[i], __ldg(d.hgate_bias + gc + i)));
              }
            }
          }
          __syncwarp();
        }
      } else {
        // BF16 / F32 / RESID: per-row operand sets A and B are loaded one chunk ahead and used alternately
        auto load_ops = [&](int c, float4 (&pre)[8], float4& bias4, float4& gate4) {
          const int lc = c * 32 + cg * 4;                         // column inside the tile
          bias4 = f4_zero();
          gate4 = f4_one();
#pragma unroll
          for (int k = 0; k < 8; ++k) pre[k] = f4_zero();
          if (lc >= ncols) return;
          if constexpr (EPI == E2B_EPI_RESID) {
            // plain (coherent) loads: the residual aliases the output buffer
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (R.aux[k]) pre[k] = *reinterpret_cast<const float4*>(R.aux[k] + lc * 4);
            if (d.gate && !per_batch_gate) gate4 = __ldg(reinterpret_cast<const float4*>(d.gate + n0 + lc));
          }
          if constexpr (EPI == E2B_EPI_F32) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (R.aux[k]) pre[k] = __ldg(reinterpret_cast<const float4*>(R.aux[k] + lc * 4));
          }
          if (d.bias) bias4 = __ldg(reinterpret_cast<const float4*>(d.bias + n0 + lc));
        };
        auto process = [&](int c, const float4 (&pre)[8], const float4& bias4, const float4& gate4) {
          const int lc = c * 32 + cg * 4;
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          stage_rows(buf, lane, v);
          __syncwarp();
          if (lc < ncols) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 a = bufr[4 * k * EPI_PITCH4];
              if constexpr (EPI == E2B_EPI_BF16) {
                const float4 o = make_float4(a.x + bias4.x, a.y + bias4.y, a.z + bias4.z, a.w + bias4.w);
                if (R.out[k]) st_bf16x4(R.out[k] + lc * 2, o, d.split);
              } else if constexpr (EPI == E2B_EPI_F32) {
                const float4 o = make_float4(a.x + bias4.x + pre[k].x, a.y + bias4.y + pre[k].y, a.z + bias4.z + pre[k].z, a.w + bias4.w + pre[k].w);
                if (R.out[k]) {
                  *reinterpret_cast<float4*>(R.out[k] + lc * 4) = o;
                  if (R.out2[k]) st_bf16x4(R.out2[k] + lc * 2, o, d.split);
                }
              } else {   // RESID
                float4 gg = gate4;
                if (per_batch_gate && R.gate[k]) gg = __ldg(reinterpret_cast<const float4*>(R.gate[k] + lc * 4));
                float4 r = pre[k];
                if (R.valid[k]) {
                  r.x = fmaf(a.x + bias4.x, gg.x, r.x);
                  r.y = fmaf(a.y + bias4.y, gg.y, r.y);
                  r.z = fmaf(a.z + bias4.z, gg.z, r.z);
                  r.w = fmaf(a.w + bias4.w, gg.w, r.w);
                }
                if (R.out[k]) {
                  *reinterpret_cast<float4*>(R.out[k] + lc * 4) = r;
                  if (R.out2[k]) st_bf16x4(R.out2[k] + lc * 2, r, d.split);
                }
              }
            }
          }
          __syncwarp();
        };
        if constexpr (EW == 4) {
          // two operand sets, used alternately (no register copies that would wait on the in-flight loads)
          float4 preA[8], preB[8];
          float4 biasA, gateA, biasB, gateB;
          load_ops(0, preA, biasA, gateA);
#pragma unroll 1
          for (int c = 0; c < BN / 32; c += 2) {
            if (c * 32 >= ncols) break;
            load_ops(c + 1, preB, biasB, gateB);
            process(c, preA, biasA, gateA);
            if ((c + 1) * 32 >= ncols) break;
            load_ops(c + 2, preA, biasA, gateA);
            process(c + 1, preB, biasB, gateB);
          }
        } else {
          // 8 epilogue warps: latency is covered by the second warp on each scheduler instead of a second operand set
          float4 preA[8];
          float4 biasA, gateA;
#pragma unroll 1
          for (int c = cw; c < BN / 32; c += CSTEP) {
            if (c * 32 >= ncols) break;
            load_ops(c, preA, biasA, gateA);
            process(c, preA, biasA, gateA);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_tempty(tw, &tempty[as]);     // one arrival per epilogue warp
    }
   }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync();          // neither CTA leaves while the pair's MMAs / remote arrivals may still touch it
  if (warp == W_ALLOC) {
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace e2b
// internal epilogue id, column tile width, epilogue warps and CTAs per tile of the most recent launch (tests assert which variant ran)
extern "C" { int e2b_gemm_last_variant[4] = {-1, 0, 0, 0}; }
namespace e2b {

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

// 2-D tiled map with a swizzle mode: dims / box innermost first, row stride in bytes
int make_tmap_swizzled(CUtensorMap* m, CUtensorMapDataType dt, const void* base, const uint64_t* dims, const uint64_t* stride1, const uint32_t* box,
                       CUtensorMapSwizzle sw) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { e2b_set_kernel_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint64_t gd[2] = {dims[0], dims[1]}, gs[1] = {stride1[0]};
  cuuint32_t bx[2] = {box[0], box[1]}, es[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { e2b_set_kernel_error("cuTensorMapEncodeTiled (swizzled 2-D) failed: %d", (int)r); return -1; }
  return 0;
}

// generic tiled map (no swizzle): dims/box innermost first, strides in bytes for dims 1..rank-1
int make_tmap_generic(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { e2b_set_kernel_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; if (i) gs[i - 1] = strides[i - 1]; }
  CUresult r = enc(m, dt, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { e2b_set_kernel_error("cuTensorMapEncodeTiled (generic) failed: %d", (int)r); return -1; }
  return 0;
}

static int num_sms() { return e2b_num_sms(); }

template <int BN, int EPI, int EW, int CG = 1>
static int launch_t(const GemmArgs& a, cudaStream_t st) {
  using Cfg = GemmCfg<BN, EW, EPI, CG>;
  static bool configured[E2B_MAX_DEVICES] = {false};
  bool& conf = configured[e2b_device_slot()];
  if (!conf) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, EPI, EW, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { e2b_set_kernel_error("gemm smem attribute: %s", cudaGetErrorString(e)); return -1; }
    conf = true;
  }
  const int m_tiles = (a.d.M + BM - 1) / BM, n_tiles = (a.d.N + BN - 1) / BN;
  int grid;
  if (CG == 2) {                                   // one cluster of two CTAs per 256-row tile
    const int pairs = ((m_tiles + 1) / 2) * n_tiles;
    grid = 2 * (pairs < num_sms() / 2 ? pairs : num_sms() / 2);
  } else {
    const int tiles = m_tiles * n_tiles;
    grid = tiles < num_sms() ? tiles : num_sms();
  }
  static const char* kinds[] = {"gemm_bf16", "gemm_f32", "gemm_geglu", "gemm_resid", "gemm_qkv", "gemm_resid", "gemm_qkv"};
  const double out_cols = (EPI == E2B_EPI_GEGLU) ? a.d.N / 2.0 : a.d.N;
  const double out_bytes = (EPI == E2B_EPI_F32) ? 4.0 : ((EPI == E2B_EPI_RESID || EPI == EPI_RESID_TMA) ? 8.0 : 2.0);
  ProfScope ps(st, kinds[EPI], a.d.M, a.d.N, a.d.K, 2.0 * a.d.M * a.d.N * a.d.K,
               2.0 * ((double)a.d.M * a.d.K + (double)a.d.N * a.d.K) + out_bytes * a.d.M * out_cols + (a.d.out_b16 ? 2.0 * a.d.M * out_cols : 0.0));
  e2b_gemm_last_variant[0] = EPI; e2b_gemm_last_variant[1] = BN; e2b_gemm_last_variant[2] = EW; e2b_gemm_last_variant[3] = CG;
  if (CG == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<BN, EPI, EW, CG>, a);
    if (e != cudaSuccess) { e2b_set_kernel_error("gemm cluster launch: %s", cudaGetErrorString(e)); return -1; }
    return 0;
  }
  gemm_kernel<BN, EPI, EW, CG><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(a);
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

// K threshold (inclusive) below which the 8-epilogue-warp configuration is used; settable for tuning / A-B tests.
// EPI_RESID through the TMA-based epilogue (1, default) or the classic load/store one (0; E2B_RESID_TMA=0), and the largest K that
// takes its 8-epilogue-warp / two-residual-tile configuration
extern "C" int e2b_gemm_resid_tma = -1;

extern "C" int e2b_gemm_resid_tma_ew8_max_k = 3072;
// EPI_QKV (bf16 mode, V as rows) through the row-per-lane / TMA-store epilogue (1, default) or the classic one (0; E2B_QKV_TMA=0)
extern "C" int e2b_gemm_qkv_tma = -1;
// CTA-pair (tcgen05 cta_group::2, 256-row tiles) variants for launches with 256-wide column tiles whose tile count fills the 74
// clusters (E2B_GEMM_CG2=0 switches them off).  tools/bench_gemm_pf.py (AB=pair), profiles/r02_gemm_cta_pair_ab.txt: +8-12 % on the
// large GEMMs standalone (GEGLU text 1365 -> 1516 TF/s), 122.9 -> 128.4 audio-s/s for the whole C2 step.
extern "C" int e2b_gemm_cta_pair = -1;
// A-operand L2 prefetch distance in K blocks of 64 (0 = off); settable for A/B tests (E2B_GEMM_PF)
extern "C" int e2b_gemm_prefetch_kb = -1;
extern "C" int e2b_gemm_ew8_max_k = 1536;   // tools/bench_gemm3.py: 8 warps win up to K=1280, tie at 2048, lose at 5120

static bool tile256(const e2b_gemm_desc* d) {
  // Narrow outputs use 128-wide tiles; GEGLU needs the 128+128 packed 256 tile.
  bool bn256 = (d->epi == E2B_EPI_GEGLU) || (d->N % 256 == 0) || (d->N > 1024);
  if (bn256 && d->epi != E2B_EPI_GEGLU) {
    // Few rows (one clip per call): when even the 128x128 tiling fits in one wave, the 128x256 tiling leaves most SMs idle
    // (single 10 s clip: 52 tiles for 148 SMs at N = 1024) -- take the narrower tile (sample() latency 190 -> 166 ms).
    const long long mt = (d->M + BM - 1) / BM;
    if (mt * ((d->N + 127) / 128) <= e2b_num_sms()) bn256 = false;
  }
  return bn256;
}

static void read_env_knobs() {
  if (e2b_gemm_resid_tma < 0) {
    const char* e = getenv("E2B_RESID_TMA");
    e2b_gemm_resid_tma = e ? atoi(e) : 1;
  }
  if (e2b_gemm_prefetch_kb < 0) {
    const char* e = getenv("E2B_GEMM_PF");
    e2b_gemm_prefetch_kb = e ? atoi(e) : 0;
  }
  if (e2b_gemm_qkv_tma < 0) {
    const char* e = getenv("E2B_QKV_TMA");
    e2b_gemm_qkv_tma = e ? atoi(e) : 1;
  }
  if (e2b_gemm_cta_pair < 0) {
    const char* e = getenv("E2B_GEMM_CG2");
    e2b_gemm_cta_pair = e ? atoi(e) : 1;
  }
  static bool once = false;
  if (!once) {
    once = true;
    const char* e = getenv("E2B_RT_EW8_MAXK");
    if (e && atoi(e) > 0) e2b_gemm_resid_tma_ew8_max_k = atoi(e);
  }
}

// does an EPI_RESID launch of this description take the TMA-based residual epilogue (the one that can emit row sums / a scaled copy)?
extern "C" int e2b_gemm_resid_uses_tma(const e2b_gemm_desc* d) {
  read_env_knobs();
  return d->epi == E2B_EPI_RESID && e2b_gemm_resid_tma && d->resid && d->ldo % 4 == 0 && d->ldr % 4 == 0 &&
         !(reinterpret_cast<uintptr_t>(d->out) & 15) && !(reinterpret_cast<uintptr_t>(d->resid) & 15) &&
         (!d->out_b16 || (d->split == 0 && d->ldo_b16 % 8 == 0 && !(reinterpret_cast<uintptr_t>(d->out_b16) & 15)));
}

// does an EPI_QKV launch of this description take the TMA-store epilogue?  (bf16 outputs, V as plain rows, 16-byte aligned rows)
static bool qkv_uses_tma(const e2b_gemm_desc* d) {
  read_env_knobs();
  const bool has_v = d->v_end > d->k_end, has_gate = d->N > d->v_end;
  return d->epi == E2B_EPI_QKV && e2b_gemm_qkv_tma && !d->qk_f32 && !d->v_f32 && d->rope && d->ldo % 8 == 0 && !(reinterpret_cast<uintptr_t>(d->out) & 15) &&
         (!has_v || (d->v_rowmajor && d->vt && d->vt_ld % 8 == 0 && !(reinterpret_cast<uintptr_t>(d->vt) & 15))) &&
         (!has_gate || (d->hgate && d->hgate_bias && d->hgate_ld % 4 == 0 && (d->N - d->v_end) % 4 == 0 && !(reinterpret_cast<uintptr_t>(d->hgate) & 15) &&
                        !(reinterpret_cast<uintptr_t>(d->hgate_bias) & 15)));
}

extern "C" int e2b_gemm_row_parts(const e2b_gemm_desc* d) {
  return (d->N + 127) / 128;      // one partial per 128 columns, independent of the tile configuration the launch takes
}

extern "C" int e2b_gemm_launch(const e2b_gemm_desc* d, cudaStream_t stream) {
  if (d->M <= 0 || d->N <= 0) return 0;
  if (d->num_src < 1 || d->num_src > E2B_MAX_SRC) { e2b_set_kernel_error("gemm: num_src %d", d->num_src); return -1; }
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.d = *d;
  int k = 0;
  for (int s = 0; s < E2B_MAX_SRC; ++s) {
    if (s < d->num_src) {
      if (d->ka[s] <= 0 || d->ka[s] % BK) { e2b_set_kernel_error("gemm: ka[%d]=%d must be a positive multiple of 64", s, d->ka[s]); return -1; }
      if (make_tmap_bf16(&a.tmA[s], d->a[s], d->M, d->ka[s], d->lda[s], BM)) return -1;
      k += d->ka[s];
    }
    a.kb_end[s] = (s < d->num_src) ? k / BK : (1 << 30);
  }
  if (k != d->K) { e2b_set_kernel_error("gemm: sum(ka)=%d != K=%d", k, d->K); return -1; }
  const bool bn256 = tile256(d);
  if (d->epi == E2B_EPI_GEGLU && d->N % 256) { e2b_set_kernel_error("gemm: GEGLU needs N %% 256 == 0 (N=%d)", d->N); return -1; }
  if (d->epi != E2B_EPI_QKV && (d->N % 4 || d->ldo % 4 || (d->out_b16 && d->ldo_b16 % 4) || (d->resid && d->ldr % 4) || (d->add_table && d->ld_add % 4) ||
                                (d->gate && d->gate_bstride % 4))) {
    e2b_set_kernel_error("gemm: N and all leading dimensions must be multiples of 4");
    return -1;
  }
  if (d->epi == E2B_EPI_QKV && ((d->q_end | d->k_end | d->v_end) % 64 || d->rows_per_batch <= 0)) {
    e2b_set_kernel_error("gemm: QKV segment ends must be multiples of 64 and rows_per_batch > 0");
    return -1;
  }
  read_env_knobs();
  a.pf_kb = e2b_gemm_prefetch_kb;
  // CTA pair (256-row tiles, tcgen05 cta_group::2): 256-wide column tiles and enough tiles to fill the 74 clusters; every epilogue
  // except the classic residual one (only reached with unaligned buffers or the fp32 mode's hi/lo copies)
  const bool pair = e2b_gemm_cta_pair && bn256 && (long long)((d->M + 255) / 256) * ((d->N + 255) / 256) >= e2b_num_sms() / 2 &&
                    (d->epi != E2B_EPI_RESID || e2b_gemm_resid_uses_tma(d));
  if (make_tmap_bf16(&a.tmB, d->w, d->N, d->K, d->ldw, pair ? 128 : (bn256 ? 256 : 128))) return -1;
  {
    // ragged last column tile: its own B box; the MMA takes N in steps of 16 per CTA, a pair splits the rounded-up tile in halves
    const int bn = bn256 ? 256 : 128, rem = d->N % bn, gran = pair ? 32 : 16, rows = (rem + gran - 1) / gran * gran;
    a.tail_rows = (rem > 0 && rows < bn) ? rows : 0;
    if (a.tail_rows && make_tmap_bf16(&a.tmBt, d->w, d->N, d->K, d->ldw, (uint32_t)(pair ? a.tail_rows / 2 : a.tail_rows))) return -1;
  }
#define E2B_DISPATCH(BN_, EW_)                                                        \
  switch (d->epi) {                                                                   \
    case E2B_EPI_BF16: return launch_t<BN_, E2B_EPI_BF16, EW_>(a, stream);            \
    case E2B_EPI_F32: return launch_t<BN_, E2B_EPI_F32, EW_>(a, stream);              \
    case E2B_EPI_RESID: return launch_t<BN_, E2B_EPI_RESID, EW_>(a, stream);          \
    case E2B_EPI_QKV: return launch_t<BN_, E2B_EPI_QKV, EW_>(a, stream);              \
    default: break;                                                                   \
  }
#define E2B_DISPATCH_PAIR(EW_)                                                        \
  switch (d->epi) {                                                                   \
    case E2B_EPI_BF16: return launch_t<256, E2B_EPI_BF16, EW_, 2>(a, stream);         \
    case E2B_EPI_F32: return launch_t<256, E2B_EPI_F32, EW_, 2>(a, stream);           \
    case E2B_EPI_QKV: return launch_t<256, E2B_EPI_QKV, EW_, 2>(a, stream);           \
    default: break;                                                                   \
  }
  if (e2b_gemm_resid_uses_tma(d)) {
    // residual stream through the TMA unit (row-per-lane epilogue); 8 epilogue warps with two residual tiles each up to K = 3072
    const uint64_t dims[2] = {(uint64_t)d->N, (uint64_t)d->M};
    const uint32_t box[2] = {32, 32};
    const uint64_t sr[1] = {(uint64_t)d->ldr * 4}, so[1] = {(uint64_t)d->ldo * 4}, sh[1] = {(uint64_t)d->ldo_b16 * 2};
    if (make_tmap_swizzled(&a.tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d->resid, dims, sr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (make_tmap_swizzled(&a.tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d->out, dims, so, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (d->out_b16 && make_tmap_swizzled(&a.tmO16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->out_b16, dims, sh, box, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
    if (!bn256) return launch_t<128, EPI_RESID_TMA, 4>(a, stream);
    if (pair) return d->K <= e2b_gemm_resid_tma_ew8_max_k ? launch_t<256, EPI_RESID_TMA, 8, 2>(a, stream) : launch_t<256, EPI_RESID_TMA, 4, 2>(a, stream);
    if (d->K <= e2b_gemm_resid_tma_ew8_max_k) return launch_t<256, EPI_RESID_TMA, 8>(a, stream);
    return launch_t<256, EPI_RESID_TMA, 4>(a, stream);
  }
  if (d->in_row_ss && ((d->epi != E2B_EPI_GEGLU && d->epi != E2B_EPI_QKV) || d->qk_f32 || d->in_row_parts <= 0)) {
    e2b_set_kernel_error("gemm: in_row_ss (norm as a row scale) is for the bf16 GEGLU / QKV epilogues");
    return -1;
  }
  if (qkv_uses_tma(d)) {
    // q | k and v rows leave through the TMA unit: 32 x 32 bf16 boxes, 64-byte swizzle
    const uint32_t box[2] = {32, 32};
    const uint64_t dqk[2] = {(uint64_t)d->k_end, (uint64_t)d->M}, sqk[1] = {(uint64_t)d->ldo * 2};
    if (make_tmap_swizzled(&a.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->out, dqk, sqk, box, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
    if (d->v_end > d->k_end) {
      const uint64_t dv[2] = {(uint64_t)(d->v_end - d->k_end), (uint64_t)d->M}, sv[1] = {(uint64_t)d->vt_ld * 2};
      if (make_tmap_swizzled(&a.tmO16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->vt, dv, sv, box, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
    }
    const bool ew8q = d->K <= e2b_gemm_ew8_max_k;
    if (!bn256) return launch_t<128, EPI_QKV_TMA, 4>(a, stream);
    if (pair) return ew8q ? launch_t<256, EPI_QKV_TMA, 8, 2>(a, stream) : launch_t<256, EPI_QKV_TMA, 4, 2>(a, stream);
    return ew8q ? launch_t<256, EPI_QKV_TMA, 8>(a, stream) : launch_t<256, EPI_QKV_TMA, 4>(a, stream);
  }
  if (d->row_ss || d->b16_scale) { e2b_set_kernel_error("gemm: row sums / scaled bf16 copy need the TMA residual epilogue (EPI_RESID, 16-byte aligned buffers, no hi/lo split)"); return -1; }
  const bool ew8 = d->K <= e2b_gemm_ew8_max_k;   // small K: epilogue-bound, use 8 epilogue warps + 3 stages
  if (bn256) {
    if (d->epi == E2B_EPI_GEGLU && pair) return ew8 ? launch_t<256, E2B_EPI_GEGLU, 8, 2>(a, stream) : launch_t<256, E2B_EPI_GEGLU, 4, 2>(a, stream);
    if (d->epi == E2B_EPI_GEGLU) return ew8 ? launch_t<256, E2B_EPI_GEGLU, 8>(a, stream) : launch_t<256, E2B_EPI_GEGLU, 4>(a, stream);
    if (pair && d->epi != E2B_EPI_RESID) { if (ew8) { E2B_DISPATCH_PAIR(8) } else { E2B_DISPATCH_PAIR(4) } }
    if (ew8) { E2B_DISPATCH(256, 8) } else { E2B_DISPATCH(256, 4) }
  } else {
    E2B_DISPATCH(128, 4)
  }
#undef E2B_DISPATCH
#undef E2B_DISPATCH_PAIR
  e2b_set_kernel_error("gemm: unknown epilogue %d", d->epi);
  return -1;
}
