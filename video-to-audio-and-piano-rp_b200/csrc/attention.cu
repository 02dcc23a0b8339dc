// Flash-style attention on tcgen05 for the x-transformers `Attention` the reference instantiates with
// softclamp_logits=True, gate_value_heads=True (e2_tts_crossatt3.py:808,813,881,914):
//
//     sim = 50 * tanh(q k^T * 64^-0.5 / 50) ; key-length mask ; softmax (fp32) ; out = attn v ;
//     out *= sigmoid(head_gate)[token, head]
//
// One persistent CTA per SM walks work items (128-query tile, head, batch item).  S = Q K^T lives in TMEM (two 128-column buffers), the softmax
// warps read it with tcgen05.ld and write P (bf16 pairs) back into TMEM with tcgen05.st -- every warp into the first 16 of the
// 32 S columns it owns -- and O += P V (A operand from TMEM; V from shared memory as plain rows = MN-major B operand, or as a
// transposed copy V^T = K-major, template VMN) accumulates in TMEM over all key tiles; the row sums come from one more 16-wide MMA
// of P against a tile of ones.
// P never touches shared memory: no st.shared + fence.proxy.async (a MEMBAR) per tile, and the 64 KB the P buffers took
// now deepen the K/V ring.  The key tile width `bk` is a per-call value (multiple of 16, <= 128): 128 for self attention, 16 for
// the T5 cross-attention's 8 keys.  Because the soft-clamp bounds every logit to [-50, 50],
// exp(sim) cannot overflow or underflow in fp32/bf16, so no running maximum and no O rescaling is needed:
// out = (sum_j exp(sim_j) v_j) / (sum_j exp(sim_j)) is evaluated directly.  q arrives pre-scaled by 64^-0.5 and
// RoPE-rotated, k RoPE-rotated, V as rows [keys, d] (or transposed, [d, keys]) -- all produced by the QKV GEMM epilogue (gemm.cu).
#include <cstdlib>

#include "kernels.h"
#include "prof.h"
#include "ptx.cuh"

namespace e2b {

int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

constexpr int ATT_SOFTMAX_WARPS = 16;      // 4 TMEM lane quarters x 4 column quarters of every S tile
constexpr int ATT_THREADS = 128 + 32 * ATT_SOFTMAX_WARPS;
// Warp roles: 0..15 softmax + epilogue, then TMA producer, MMA issuer, TMEM allocator, spare.  The single-thread issuers get
// the HIGHEST warp ids on purpose: the scheduler arbitrates highest-warp-id-first (B300_MICROARCH.md), and with ids 0/1 they
// were starved by the four busy softmax warps sharing their schedulers, delaying every TMA / MMA issue.
constexpr int ATT_W_TMA = ATT_SOFTMAX_WARPS, ATT_W_MMA = ATT_SOFTMAX_WARPS + 1, ATT_W_ALLOC = ATT_SOFTMAX_WARPS + 2;
constexpr int ATT_BQ = 128, ATT_BK = 128 /* widest key tile; the per-call width is AttnArgs::bk */, ATT_D = 64;
constexpr int ATT_KV = 5;                       // K/V ring depth
constexpr int ATT_ON = 80;                      // PV MMA N: 64 value channels + a ones row (col 64 = row sum of P) + 15 zero rows
constexpr int ATT_VATOM = ATT_ON * 128;         // one V^T swizzle atom: 80 rows x 128 B (rows 64..79 are constants written once)
constexpr int ATT_VSTAGE = 2 * ATT_VATOM;       // two 64-key atoms per 128-key tile
constexpr int ATT_SQ = 0;                       // 2 x 16 KB  Q   [128 q, 64 d]        (double-buffered across work items)
constexpr int ATT_SK = 2 * 16384;               // ATT_KV x 16 KB  K   [128 keys, 64 d]
constexpr int ATT_SV = ATT_SK + ATT_KV * 16384; // ATT_KV x 20 KB  V^T 2 x [80 rows, 64 keys]
constexpr int ATT_SONES = ATT_SV + ATT_KV * 16384;   // row-major V takes 16 KB per slot: the ones tile (2 KB) sits in the tail of the V region
constexpr int ATT_BAR = ATT_SV + ATT_KV * ATT_VSTAGE;
constexpr int ATT_LENS = ATT_BAR + 256;          // clamped kv length of every kv sequence of the call
constexpr int ATT_MAXB = 1024;                  // more kv sequences than this: lengths are read from global memory instead
constexpr int ATT_SMEM = ATT_LENS + ATT_MAXB * 4;
static_assert(ATT_SMEM <= 232448, "shared memory budget");
constexpr uint32_t ATT_TMEM_COLS = 512;         // S0 @0, S1 @128 (P aliased: key columns 32c..32c+31 of a tile -> TMEM columns 32c..32c+15),
                                                // O0 @256, O1 @336 (80 columns each)
constexpr uint32_t ATT_TMEM_O = 256;

// Soft-clamp + exponent in one polynomial.  With w = z^2 and |z| / clamp < 0.5,
//   log2(e) * clamp * tanh(z / clamp) = z * (c0 + w (c1 + w (c2 + w (c3 + w c4))))     (abs. error < 3e-4 in the exponent
// at the edge, < 1e-5 for |z/clamp| < 0.3), i.e. 6 FMA-pipe ops and ONE MUFU (ex2) per logit instead of tanh.approx (only
// 2^-11 accurate, amplified 50x by the exp) + ex2.  Larger logits (rare) take the exact exp-based tanh.
struct ClampPoly { float c0, c1, c2, c3, c4, wmax, clamp, wlo, ex_a, ex_b; };

// log2(e) * clamp * tanh(z / clamp) from the exp-based identity; ex_a = 2 log2(e) / clamp, ex_b = log2(e) * clamp
__device__ __forceinline__ float softclamp_exp2_arg_exact(float z, float ex_a, float ex_b) {
  const float e = ex2_approx(z * ex_a);                    // exp(2 z / clamp)
  return fmaf(-2.0f * ex_b, __frcp_rn(1.0f + e), ex_b);
}

// p = 2^(z * poly(z^2)) for the 32 logits of one warp-tile, packed to bf16 pairs, in packed-pair arithmetic (FFMA2 / FMUL2:
// half the issue slots of the scalar forms).  HI selects the degree-9 series.
// 2^a for a pair on the FMA / ALU pipes only (no MUFU): Cody-Waite split a = n + f with the magic-number rounding, a cubic for
// 2^f on [-0.5, 0.5] (minimax fit, relative error < 7.5e-5, far below the bf16 rounding of P) and the integer n added into the exponent field.
// |a| <= log2(e) * clamp = 72.2, so neither the split nor the exponent add can overflow.
__device__ __forceinline__ void exp2_pair_fma(uint64_t a, float& r0, float& r1) {
  const uint64_t magic = f32x2_pack(12582912.0f, 12582912.0f);                    // 1.5 * 2^23
  const uint64_t t = f32x2_add(a, magic);                                         // n sits in the low mantissa bits
  const uint64_t f = f32x2_sub(a, f32x2_sub(t, magic));                           // a - n
  uint64_t p = f32x2_fma(f, f32x2_pack(0.05517167f, 0.05517167f), f32x2_pack(0.24261113f, 0.24261113f));
  p = f32x2_fma(f, p, f32x2_pack(0.69326097f, 0.69326097f));
  p = f32x2_fma(f, p, f32x2_pack(0.99992806f, 0.99992806f));
  float p0, p1, t0, t1;
  f32x2_unpack(p, p0, p1);
  f32x2_unpack(t, t0, t1);
  r0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  r1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// p = 2^(z * poly(z^2)) for the logits of one warp-tile, packed to bf16 pairs, in packed-pair arithmetic (FFMA2 / FMUL2:
// half the issue slots of the scalar forms).  HI selects the degree-9 series.  POLY: every fourth pair takes its exponential
// from exp2_pair_fma instead of the MUFU unit -- ex2 (16 / clk / SM) is the binding pipe of this kernel while the FMA pipe has
// slack (tools/pipe_bench2: 277 clk per block, 256 of them MUFU; 196 without any ex2).
template <bool HI, int NP, bool POLY>
__device__ __forceinline__ void exp_block(const uint32_t (&v)[32], const ClampPoly& cp, uint32_t (&pk)[16]) {
  const uint64_t c0 = f32x2_pack(cp.c0, cp.c0), c1 = f32x2_pack(cp.c1, cp.c1), c2 = f32x2_pack(cp.c2, cp.c2);
  const uint64_t c3 = f32x2_pack(cp.c3, cp.c3), c4 = f32x2_pack(cp.c4, cp.c4);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const uint64_t z = f32x2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    const uint64_t w = f32x2_mul(z, z);
    uint64_t q;
    if (HI) {
      q = f32x2_fma(w, c4, c3);
      q = f32x2_fma(w, q, c2);
      q = f32x2_fma(w, q, c1);
    } else {
      q = f32x2_fma(w, c2, c1);
    }
    q = f32x2_fma(w, q, c0);
    const uint64_t a = f32x2_mul(z, q);
    if (POLY && (i & 3) == 3) {
      float r0, r1;
      exp2_pair_fma(a, r0, r1);
      pk[i] = pack_bf16(r0, r1);
    } else {
      float a0, a1;
      f32x2_unpack(a, a0, a1);
      pk[i] = pack_bf16(ex2_approx(a0), ex2_approx(a1));
    }
  }
}

struct AttnArgs {
  CUtensorMap tmQ, tmK, tmV;
  e2b_attn_desc d;
  ClampPoly cp;
  int bk;                      // key tile width of this call (multiple of 16, <= ATT_BK); the K tensor map's box has bk rows
  long long* dbg;
};

// Persistent kernel: one CTA per SM loops over work items (q-tile, head, sequence).  The TMA producer and the MMA issuer
// run ahead across item boundaries (Q and O are double-buffered, K/V tiles stream through one ring for all items), so the
// softmax warps never see a per-item prologue: ncu on the one-CTA-per-item version showed ~14 % of their samples waiting
// for the first S tile of each CTA plus the launch/alloc/teardown of 14 336 CTAs per call.
// Optional timeline for tools/attn_timeline.py: CTA 0 stamps clock64 per KV tile into dbg[g * 8 + slot] (null in production).
static long long* g_att_dbg_host = nullptr;      // set by e2b_attention_set_debug; passed to the kernel as an argument
constexpr int ATT_DBG_TILES = 96;
#define dbg_stamp(g, slot)                                                                                      \
  do {                                                                                                          \
    if (args.dbg != nullptr && blockIdx.x == 0 && (g) < ATT_DBG_TILES) args.dbg[(g) * 8 + (slot)] = clock64();   \
  } while (0)

struct ItemCursor {          // iterates the (item, kv-tile) sequence of this CTA; every role walks the same sequence
  int it, item, j, nt, b, h, qt, kvb, kv_len;
  bool valid;
};
// item = (b * heads + h) * qtiles + qt advances by gridDim.x per step.  The step is decomposed once into mixed-radix digits so
// that an item switch is a few adds and compares, and the clamped kv lengths of the call sit in shared memory: a timeline
// showed the direct decode (five integer divisions, each with a MUFU.RCP queued behind the softmax warps' ex2 stream, and a
// kv_lens global load whose destination register was spilled, i.e. waited for at once) costing ~2K cycles per item.
struct ItemWalk {
  ItemCursor c;
  int dq, dh, db, dkvb, bk;
  const int* lens;           // shared memory, [kv batches] (null: more than ATT_MAXB sequences, read global memory)
};
__device__ __forceinline__ int clamped_kv_len(const e2b_attn_desc& d, int kvb) {
  const int kv_len = d.kv_lens ? (__ldg(d.kv_lens + kvb) + d.kv_lens_add) : d.kv_rows_per_batch;
  return max(0, min(kv_len, d.kv_rows_per_batch));
}
__device__ __forceinline__ void cursor_fetch(ItemWalk& w, const e2b_attn_desc& d, int total) {
  ItemCursor& c = w.c;
  c.valid = c.item < total;
  c.j = 0;
  c.kv_len = !c.valid ? 0 : w.lens ? w.lens[c.kvb] : clamped_kv_len(d, c.kvb);
  c.nt = max(1, (c.kv_len + w.bk - 1) / w.bk);            // at least one (fully masked) tile so O is defined
}
__device__ __forceinline__ void walk_init(ItemWalk& w, const e2b_attn_desc& d, int qtiles, int total, const int* lens, int bk) {
  const int g = gridDim.x;
  w.bk = bk;
  w.dq = g % qtiles;
  const int gh = g / qtiles;
  w.dh = gh % d.heads;
  w.db = gh / d.heads;
  w.dkvb = d.kv_batch_mod > 0 ? w.db % d.kv_batch_mod : w.db;
  w.lens = lens;
  ItemCursor& c = w.c;
  c.it = 0;
  c.item = blockIdx.x;
  c.qt = c.item % qtiles;
  const int bh = c.item / qtiles;
  c.h = bh % d.heads;
  c.b = bh / d.heads;
  c.kvb = d.kv_batch_mod > 0 ? c.b % d.kv_batch_mod : c.b;
  cursor_fetch(w, d, total);
}
__device__ __forceinline__ void walk_next_item(ItemWalk& w, const e2b_attn_desc& d, int qtiles, int total) {
  ItemCursor& c = w.c;
  ++c.it;
  c.item += gridDim.x;
  int carry = 0;
  c.qt += w.dq;
  if (c.qt >= qtiles) { c.qt -= qtiles; carry = 1; }
  c.h += w.dh + carry;
  carry = 0;
  if (c.h >= d.heads) { c.h -= d.heads; carry = 1; }
  c.b += w.db + carry;
  if (d.kv_batch_mod > 0) {
    c.kvb += w.dkvb + carry;
    if (c.kvb >= d.kv_batch_mod) c.kvb -= d.kv_batch_mod;
    if (c.kvb >= d.kv_batch_mod) c.kvb -= d.kv_batch_mod;
  } else {
    c.kvb = c.b;
  }
  cursor_fetch(w, d, total);
}
__device__ __forceinline__ void walk_next_tile(ItemWalk& w, const e2b_attn_desc& d, int qtiles, int total) {
  if (++w.c.j == w.c.nt) walk_next_item(w, d, qtiles, total);
}

// VMN: V arrives as plain rows [keys, 64 d] (like K) and is the MN-major B operand of the P V product; the row sums then come
// from a second, 16-wide MMA against a constant tile of ones.  !VMN: V^T [64 d, keys] (K-major B) whose atoms carry a ones row.
template <bool VMN, bool POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1) attention_kernel(const __grid_constant__ AttnArgs args) {
  constexpr int VSTAGE = VMN ? 16384 : ATT_VSTAGE;          // bytes of V per ring slot
  // Used directly (no integer round-trip) so the compiler keeps the shared address space; SWIZZLE_128B needs a 1024-byte
  // aligned base: with no static shared memory the dynamic window starts at offset 0 -- checked once below.
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
    printf("e2b: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_BAR);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* o_full = bars + 4;        // [2]
  uint64_t* o_empty = bars + 6;       // [2] (one arrival per softmax warp)
  uint64_t* s_full = bars + 8;        // [2] S(g) complete in TMEM buffer g & 1
  uint64_t* p_full = bars + 10;       // [2] (one arrival per softmax warp) P(g) written over S(g)
  uint64_t* p_empty = bars + 12;      // [2] PV(g) complete: TMEM buffer g & 1 may take S(g + 2)
  uint64_t* kv_full = bars + 14;      // [ATT_KV]
  uint64_t* kv_empty = bars + 14 + ATT_KV;   // [ATT_KV]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14 + 2 * ATT_KV);

  const e2b_attn_desc& d = args.d;
  const int bk = args.bk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtiles = (d.q_rows_per_batch + ATT_BQ - 1) / ATT_BQ;
  const int total = qtiles * d.heads * d.batch;
  // One elected arrival per softmax warp (after __syncwarp): mbarrier arrivals are lane-serialised shared-memory atomics, and
  // 512 per-thread arrivals per S tile sat on the critical path (ncu: warps spinning on s_full, no pipe busy).
  constexpr uint32_t NSOFT = ATT_SOFTMAX_WARPS;

  if (warp == ATT_W_TMA && lane == 0) {
    tma_prefetch_desc(&args.tmQ);
    tma_prefetch_desc(&args.tmK);
    tma_prefetch_desc(&args.tmV);
  }
  if (warp == ATT_W_MMA && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_empty[s], NSOFT);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], NSOFT);
      mbar_init(&p_empty[s], 1);
    }
    for (int s = 0; s < ATT_KV; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == ATT_W_ALLOC) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  int* lens_s = reinterpret_cast<int*>(smem + ATT_LENS);
  {
    const int kvb_n = d.kv_batch_mod > 0 ? d.kv_batch_mod : d.batch;
    if (kvb_n <= ATT_MAXB) {
      for (int i = threadIdx.x; i < kvb_n; i += ATT_THREADS) lens_s[i] = clamped_kv_len(d, i);
    } else {
      lens_s = nullptr;
    }
  }
  if constexpr (VMN) {
    // a K-major [16 x 16] tile of ones (16 rows of 128 B, all 1.0): P times it puts the row sums of the bf16 P the MMA actually
    // used into 16 extra accumulator columns (64..79 of O)
    for (int i = threadIdx.x; i < 2048 / 16; i += ATT_THREADS) {
      const uint32_t one2 = 0x3F803F80u;   // two bf16 1.0
      *reinterpret_cast<uint4*>(smem + ATT_SONES + i * 16) = make_uint4(one2, one2, one2, one2);
    }
  } else {
    // rows 64..79 of every V^T atom: row 64 = 1.0 (so column 64 of O accumulates the row sums of the bf16 P the MMA actually
    // used), rows 65..79 = 0.  A whole 128-byte row is constant, so the 128-byte swizzle does not matter.  TMA only ever
    // rewrites rows 0..63.
    for (int i = threadIdx.x; i < ATT_KV * 2 * 16 * 8; i += ATT_THREADS) {
      const int atom = i / 128, rem = i % 128, row = rem / 8, chunk = rem % 8;
      const uint32_t one2 = 0x3F803F80u;   // two bf16 1.0
      const uint32_t val = row == 0 ? one2 : 0u;
      *reinterpret_cast<uint4*>(smem + ATT_SV + atom * ATT_VATOM + (64 + row) * 128 + chunk * 16) = make_uint4(val, val, val, val);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool two_atoms = bk > 64;                     // keys 64.. of a tile live in the second V^T atom

  // Producer / issuer roles run on their whole warp with uniform control flow; only the TMA / MMA / commit instructions are
  // under elect_one() (a `lane == 0` role branch makes the compiler wrap every UTCHMMA / UTMALDG in a divergence loop).
  if (warp == ATT_W_TMA) {
    // ------------------------------------------------------------ TMA producer
    // The K/V working set of a call (hundreds of MB) does not live in L2, and a tile can only be requested into shared
    // memory once its ring slot is free, i.e. at most ATT_KV tiles ahead.  A second cursor therefore runs ATT_PF tiles further
    // ahead and asks the TMA unit to pull those tiles into L2 (no shared-memory cost), so the real loads are L2 hits.
    constexpr int ATT_PF = 6;
    ItemWalk wc, wpf;
    walk_init(wc, d, qtiles, total, lens_s, bk);
    walk_init(wpf, d, qtiles, total, lens_s, bk);
    ItemCursor& c = wc.c;
    ItemCursor& pf = wpf.c;
    const uint32_t kv_bytes = VMN ? (uint32_t)bk * 256u : (uint32_t)bk * 128u + (two_atoms ? 16384u : 8192u);
    auto prefetch_tile = [&](const ItemCursor& t) {
      if (t.j == 0) tma_prefetch_l2_2d(&args.tmQ, d.q_col0 + t.h * ATT_D, t.b * d.q_rows_per_batch + t.qt * ATT_BQ);
      tma_prefetch_l2_2d(&args.tmK, d.k_col0 + t.h * ATT_D, t.kvb * d.kv_rows_per_batch + t.j * bk);
      if constexpr (VMN) {
        tma_prefetch_l2_2d(&args.tmV, d.v_col0 + t.h * ATT_D, t.kvb * d.kv_rows_per_batch + t.j * bk);
      } else {
        const int vr = (t.kvb * d.heads + t.h) * ATT_D;
        tma_prefetch_l2_2d(&args.tmV, t.j * bk, vr);
        if (two_atoms) tma_prefetch_l2_2d(&args.tmV, t.j * bk + 64, vr);
      }
    };
    for (int i = 0; i < ATT_PF + ATT_KV && pf.valid; ++i) {
      if (elect_one()) prefetch_tile(pf);
      __syncwarp();
      walk_next_tile(wpf, d, qtiles, total);
    }
    int g = 0;
    while (c.valid) {
      const int qb = c.it & 1;
      mbar_wait(&q_empty[qb], ((c.it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[qb], 16384);
        tma_load_2d(smem + ATT_SQ + qb * 16384, &args.tmQ, &q_full[qb], d.q_col0 + c.h * ATT_D, c.b * d.q_rows_per_batch + c.qt * ATT_BQ);
      }
      __syncwarp();
      const int vrow = (c.kvb * d.heads + c.h) * ATT_D;
      for (int j = 0; j < c.nt; ++j, ++g) {
        const int s = g % ATT_KV;
        mbar_wait(&kv_empty[s], ((g / ATT_KV) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[s], kv_bytes);
          tma_load_2d(smem + ATT_SK + s * 16384, &args.tmK, &kv_full[s], d.k_col0 + c.h * ATT_D, c.kvb * d.kv_rows_per_batch + j * bk);
          if constexpr (VMN) {
            tma_load_2d(smem + ATT_SV + s * VSTAGE, &args.tmV, &kv_full[s], d.v_col0 + c.h * ATT_D, c.kvb * d.kv_rows_per_batch + j * bk);
          } else {
            tma_load_2d(smem + ATT_SV + s * VSTAGE, &args.tmV, &kv_full[s], j * bk, vrow);
            if (two_atoms) tma_load_2d(smem + ATT_SV + s * VSTAGE + ATT_VATOM, &args.tmV, &kv_full[s], j * bk + 64, vrow);
          }
          if (pf.valid) prefetch_tile(pf);
        }
        __syncwarp();
        if (pf.valid) walk_next_tile(wpf, d, qtiles, total);
      }
      walk_next_item(wc, d, qtiles, total);
    }
  } else if (warp == ATT_W_MMA) {
    // ------------------------------------------------------------ S = Q K^T issuer
    // Two issuing threads (this one and the PV issuer below) in different warps / schedulers: a clock64 timeline showed a single
    // thread needs ~1000 cycles to issue one PV group (8 tcgen05.mma + commits), ~350 for one S group and ~90 per mbarrier
    // probe -- ~3K cycles of serial work per tile, which (not the softmax, ~1.9K) set the kernel's pace.
    const uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, (uint32_t)bk);
    ItemWalk ws;
    walk_init(ws, d, qtiles, total, lens_s, bk);
    ItemCursor& cs = ws.c;
    for (int gs = 0; cs.valid; ++gs) {
      const int sb = gs & 1, ks = gs % ATT_KV, qb = cs.it & 1;
      if (cs.j == 0) mbar_wait(&q_full[qb], (cs.it >> 1) & 1);
      mbar_wait(&kv_full[ks], (gs / ATT_KV) & 1);
      // TMEM buffer sb holds S(g-2) and then P(g-2) in the same columns: it is free once PV(g-2) has read P(g-2)
      mbar_wait(&p_empty[sb], ((gs >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        dbg_stamp(gs, 1);
        const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(smem + ATT_SQ + qb * 16384));
        const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(smem + ATT_SK + ks * 16384));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_bf16_ss(tmem_base + sb * ATT_BK, dq + k * UMMA_K_STEP_ENC, dk + k * UMMA_K_STEP_ENC, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[sb]);
        if (cs.j == cs.nt - 1) umma_commit(&q_empty[qb]);          // Q buffer is free once the item's last S has completed
        dbg_stamp(gs, 4);
      }
      __syncwarp();
      walk_next_tile(ws, d, qtiles, total);
    }
  } else if (warp == ATT_W_ALLOC) {
    // ------------------------------------------------------------ O += P V issuer (A = P from TMEM, B = V^T from shared memory)
    constexpr uint32_t idesc_o = VMN ? umma_idesc_bf16_bmn(ATT_BQ, ATT_D) : umma_idesc_bf16(ATT_BQ, ATT_ON);
    constexpr uint32_t idesc_1 = umma_idesc_bf16(ATT_BQ, 16);          // VMN: P x ones -> row sums
    const int ksteps = bk / 16;
    ItemWalk wp;
    walk_init(wp, d, qtiles, total, lens_s, bk);
    ItemCursor& cp = wp.c;
    for (int gp = 0; cp.valid; ++gp) {
      const int pb = gp & 1, ks = gp % ATT_KV, ob = cp.it & 1;
      mbar_wait(&p_full[pb], (gp >> 1) & 1);
      if (cp.j == 0) mbar_wait(&o_empty[ob], ((cp.it >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        dbg_stamp(gp, 2);
        const uint32_t tmem_o = tmem_base + ATT_TMEM_O + ob * ATT_ON;
        const uint32_t tmem_p = tmem_base + pb * ATT_BK;
        const uint64_t dv0 = umma_desc_kmajor_sw128(smem_u32(smem + ATT_SV + ks * VSTAGE));   // (same encoding for the MN-major tile)
        const uint64_t d1 = umma_desc_kmajor_sw128(smem_u32(smem + ATT_SONES));
        for (int kk = 0; kk < ksteps; ++kk) {
          // P: keys 16kk..16kk+15 are 8 TMEM columns inside the 16 the owning softmax warp wrote: 32 (kk / 2) + 8 (kk % 2)
          const uint32_t ta = tmem_p + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8);
          const uint32_t acc = (cp.j | kk) != 0 ? 1u : 0u;
          if constexpr (VMN) {
            // V rows [keys, 64 d], 128 B per key, SWIZZLE_128B: 16 keys of a K-step are two 8-row groups, +2 KB per step
            umma_bf16_ts(tmem_o, ta, dv0 + (uint64_t)(kk * (2048 >> 4)), idesc_o, acc);
            umma_bf16_ts(tmem_o + ATT_D, ta, d1, idesc_1, acc);
          } else {
            // V^T: atom (kk >> 2) is +10 KB (encoded >> 4), then 32 B per 16-key step inside the swizzle atom
            const uint64_t dv = dv0 + (uint64_t)((kk >> 2) * (ATT_VATOM >> 4) + (kk & 3) * UMMA_K_STEP_ENC);
            umma_bf16_ts(tmem_o, ta, dv, idesc_o, acc);
          }
        }
        umma_commit(&kv_empty[ks]);      // S(g) finished long before P(g) existed, so this also covers K of the slot
        umma_commit(&p_empty[pb]);
        if (cp.j == cp.nt - 1) umma_commit(&o_full[ob]);
        dbg_stamp(gp, 3);
      }
      __syncwarp();
      walk_next_tile(wp, d, qtiles, total);
    }
  } else if (warp < ATT_SOFTMAX_WARPS) {
    // ------------------------------------------------------------ softmax + epilogue
    // 16 warps: warp (4c + q) owns TMEM lane quarter q (query rows 32q..32q+31) and key columns 32c..32c+31 of every
    // S tile.  (ncu on the 4- and 8-warp versions: IPC per scheduler ~0.3, nothing saturated -- latency bound.)
    const int sw = warp;
    const int quarter = sw & 3, cq = sw >> 2;
    const int r = quarter * 32 + lane;                  // row inside the tile == TMEM lane
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const ClampPoly cp = args.cp;
    const int c0 = cq * 32;                             // first key column of this warp inside the tile
    const bool col_live = c0 < bk;                      // with bk < 128 the last column quarters of a tile do not exist
    ItemWalk wk;
    walk_init(wk, d, qtiles, total, lens_s, bk);
    ItemCursor& c = wk.c;
    // Deferred item epilogue: O of item i is read out after the FIRST tile of item i+1, when its last PV has long completed,
    // so the o_full wait and the stores are off the critical path (O and Q are double-buffered across items).
    struct Pending {
      bool valid, q_valid;
      int ob;
      uint32_t parity;
      float gate;
      __nv_bfloat16* out;
    } pend;
    pend.valid = false;
    auto flush = [&](const Pending& pd) {
      mbar_wait(&o_full[pd.ob], pd.parity);
      tc_fence_after();
      uint32_t v[16], ls[1];
      tmem_ld16(tmem_base + ATT_TMEM_O + pd.ob * ATT_ON + lane_base + cq * 16, v);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(ls[0]) : "r"(tmem_base + ATT_TMEM_O + pd.ob * ATT_ON + lane_base + 64) : "memory");
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[pd.ob]);
      const float l = __uint_as_float(ls[0]);              // row sum of the bf16 probabilities, from the ones row of V^T
      if (pd.q_valid) {
        const float scale = l > 0.f ? pd.gate / l : 0.f;
        uint4* o4 = reinterpret_cast<uint4*>(pd.out);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[8 * i + 0]) * scale, __uint_as_float(v[8 * i + 1]) * scale);
          u.y = pack_bf16(__uint_as_float(v[8 * i + 2]) * scale, __uint_as_float(v[8 * i + 3]) * scale);
          u.z = pack_bf16(__uint_as_float(v[8 * i + 4]) * scale, __uint_as_float(v[8 * i + 5]) * scale);
          u.w = pack_bf16(__uint_as_float(v[8 * i + 6]) * scale, __uint_as_float(v[8 * i + 7]) * scale);
          o4[i] = u;
        }
      }
    };
    int g = 0;
    while (c.valid) {
      const int q_pos = c.qt * ATT_BQ + r;
      const bool q_valid = q_pos < d.q_rows_per_batch;
      const bool warp_valid = (c.qt * ATT_BQ + quarter * 32) < d.q_rows_per_batch;
      // value-head gate of this row: loaded now, used in the (deferred) epilogue
      const float gate_v = (d.hgate && q_valid) ? __ldg(d.hgate + (size_t)(c.b * d.q_rows_per_batch + q_pos) * d.hgate_ld + c.h) : 1.0f;
      for (int j = 0; j < c.nt; ++j, ++g) {
        const int s = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        const int nvalid = min(bk, c.kv_len - j * bk);  // live keys of this tile
        if (warp == 0 && lane == 0) dbg_stamp(g, 5);
        mbar_wait(&s_full[s], ph);
        tc_fence_after();
        if (warp == 0 && lane == 0) dbg_stamp(g, 6);
        const uint32_t t_blk = tmem_base + lane_base + s * ATT_BK + c0;     // this warp's 32 S columns; P goes to the first 16
        if (col_live) {
          const int ncol = nvalid - c0;                 // live logits of this warp's block (<= 0: none)
          uint32_t pk[16];
          if (warp_valid && ncol > 0) {
            // One 32-column block per warp and tile: the polynomial (FMA pipe) and ex2 (MUFU pipe, 16/clk/SM: the binding unit of
            // this kernel, tools/pipe_bench.cu) streams of the logits sit in ONE basic block per tier so that they interleave.
            // A block with at most 16 live logits (the last column quarter at bk = 112, the 8..16-key cross-attention tile) only
            // computes its first half: the MUFU pipe is shared by the four warps of a scheduler.
            uint32_t v[32];
            tmem_ld32(t_blk, v);
            tmem_ld_wait();
            if (ncol < 32) {                            // columns past the last key hold stale TMEM: take them out of the tier test
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i >= ncol) v[i] = 0u;
            }
            float wm = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) wm = fmaxf(wm, fabsf(__uint_as_float(v[i])));
            const bool lo = !__any_sync(0xffffffffu, wm > cp.wlo);
            const bool hi = !lo && !__any_sync(0xffffffffu, wm * wm >= cp.wmax);
            if (ncol > 16) {
              if (lo) {
                exp_block<false, 16, POLY>(v, cp, pk);        // |z/clamp| <= 0.16: degree-5 series exact to 1e-5 in the exponent
              } else if (hi) {
                exp_block<true, 16, POLY>(v, cp, pk);
              } else {                                  // rare: logits beyond the series' range -> exact tanh for the block
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  pk[i] = pack_bf16(ex2_approx(softclamp_exp2_arg_exact(__uint_as_float(v[2 * i]), cp.ex_a, cp.ex_b)),
                                    ex2_approx(softclamp_exp2_arg_exact(__uint_as_float(v[2 * i + 1]), cp.ex_a, cp.ex_b)));
              }
            } else {
#pragma unroll
              for (int i = 8; i < 16; ++i) pk[i] = 0u;
              if (lo) {
                exp_block<false, 8, POLY>(v, cp, pk);
              } else if (hi) {
                exp_block<true, 8, POLY>(v, cp, pk);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  pk[i] = pack_bf16(ex2_approx(softclamp_exp2_arg_exact(__uint_as_float(v[2 * i]), cp.ex_a, cp.ex_b)),
                                    ex2_approx(softclamp_exp2_arg_exact(__uint_as_float(v[2 * i + 1]), cp.ex_a, cp.ex_b)));
              }
            }
            if (ncol < 32) {                            // ragged last tile: keys beyond kv_len contribute exactly zero
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const uint32_t keep = (2 * i + 1 < ncol) ? 0xffffffffu : (2 * i < ncol) ? 0x0000ffffu : 0u;
                pk[i] &= keep;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          }
          tmem_st16(t_blk, pk);                         // P over the first 16 of the 32 columns this warp has just read
          tmem_st_wait();
        }
        if (warp == 0 && lane == 0) dbg_stamp(g, 0);
        tc_fence_before();             // this lane's tcgen05.ld / tcgen05.st are complete: order them before the arrival
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (j == 0 && pend.valid) {    // previous item's O: its last PV completed during this tile
          flush(pend);
          pend.valid = false;
        }
      }
      pend.valid = true;
      pend.q_valid = q_valid;
      pend.ob = c.it & 1;
      pend.parity = (c.it >> 1) & 1;
      pend.gate = gate_v;
      pend.out = reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)(c.b * d.q_rows_per_batch + q_pos) * d.ldo + c.h * ATT_D + cq * 16;
      walk_next_item(wk, d, qtiles, total);
    }
    if (pend.valid) flush(pend);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATT_W_ALLOC) tmem_dealloc(tmem_base, ATT_TMEM_COLS);
}

}  // namespace e2b

using namespace e2b;

extern "C" int e2b_attention_set_debug(long long* dev_buf) {
  e2b::g_att_dbg_host = dev_buf;
  return 0;
}

// > 0: use this key tile width instead of the per-call choice (A/B tests: 128 = the fixed tiling of round 1)
extern "C" { int e2b_attention_force_bk = 0; }
// 0: this kernel; 1: the round-1 kernel (attention_v1.cu).  Initialised from the environment (E2B_ATTN=v1) on first use.
extern "C" { int e2b_attention_impl = -1; }
// != 0: a quarter of the exponentials of every softmax block are evaluated on the FMA pipe (E2B_ATTN_POLY=1; measured slower, off)
extern "C" { int e2b_attention_poly = 0; }
extern "C" int e2b_attention_v1_launch(const e2b_attn_desc* d, cudaStream_t stream);

extern "C" int e2b_attention_launch(const e2b_attn_desc* d, cudaStream_t stream) {
  if (e2b_attention_impl < 0) {
    const char* e = getenv("E2B_ATTN");
    e2b_attention_impl = (e && (e[0] == 'v' ? e[1] == '1' : e[0] == '1')) ? 1 : 0;
    const char* f = getenv("E2B_ATTN_BK");
    if (f && atoi(f) > 0) e2b_attention_force_bk = atoi(f);
    const char* p = getenv("E2B_ATTN_POLY");
    if (p) e2b_attention_poly = atoi(p);
  }
  if (e2b_attention_impl == 1) {
    if (d->v_rowmajor) { e2b_set_kernel_error("attention: the round-1 kernel takes V^T only"); return -1; }
    return e2b_attention_v1_launch(d, stream);
  }
  if (d->batch <= 0 || d->heads <= 0 || d->q_rows_per_batch <= 0) return 0;
  if (d->kv_rows_per_batch <= 0) { e2b_set_kernel_error("attention: kv_rows_per_batch must be positive"); return -1; }
  if ((d->ldo % 8) || (reinterpret_cast<uintptr_t>(d->out) & 15)) { e2b_set_kernel_error("attention: out must be 16-byte aligned"); return -1; }
  AttnArgs a;
  a.d = *d;
  a.dbg = g_att_dbg_host;
  {
    const double L2E = 1.4426950408889634, c = d->softclamp > 0 ? d->softclamp : 50.0, c2 = c * c;
    a.cp.c0 = (float)L2E;
    a.cp.c1 = (float)(-L2E / (3.0 * c2));
    a.cp.c2 = (float)(2.0 * L2E / (15.0 * c2 * c2));
    a.cp.c3 = (float)(-17.0 * L2E / (315.0 * c2 * c2 * c2));
    a.cp.c4 = (float)(62.0 * L2E / (2835.0 * c2 * c2 * c2 * c2));
    a.cp.wmax = (float)(0.25 * c2);
    a.cp.clamp = (float)c;
    a.cp.wlo = (float)(0.16 * c);
    a.cp.ex_a = (float)(2.0 * L2E / c);
    a.cp.ex_b = (float)(L2E * c);
  }
  const uint64_t q_rows = (uint64_t)d->batch * d->q_rows_per_batch;
  const int kv_batches = d->kv_batch_mod > 0 ? d->kv_batch_mod : d->batch;
  const uint64_t k_rows = (uint64_t)kv_batches * d->kv_rows_per_batch;
  if (make_tmap_bf16(&a.tmQ, d->q, q_rows, (uint64_t)d->q_col0 + d->heads * 64, d->ldq, ATT_BQ)) return -1;
  {
    // Key tile width.  Sequences longer than one tile use the full 128: an even split (782 keys as 7 x 112 instead of 6 x 128 + 14)
    // measured 8 % SLOWER with both V layouts (tools/bench_attention.py) -- the softmax warps already skip the dead columns of a
    // ragged last tile, so an even split saves no MUFU work, and it turns one nearly free tile into a seventh full-latency one.
    // Short key sets (the T5 cross-attention: 8 keys) take the smallest multiple of 16 that holds them.
    a.bk = d->kv_rows_per_batch >= ATT_BK ? ATT_BK : (d->kv_rows_per_batch + 15) / 16 * 16;
    if (e2b_attention_force_bk > 0) a.bk = e2b_attention_force_bk;
    if (a.bk < 16 || a.bk > ATT_BK || a.bk % 16) { e2b_set_kernel_error("attention: key tile width %d must be a multiple of 16 in [16,128]", a.bk); return -1; }
  }
  if (make_tmap_bf16(&a.tmK, d->k, k_rows, (uint64_t)d->k_col0 + d->heads * 64, d->ldk, (uint32_t)a.bk)) return -1;
  const bool vmn = d->v_rowmajor != 0;
  if (vmn) {
    if (make_tmap_bf16(&a.tmV, d->vt, k_rows, (uint64_t)d->v_col0 + d->heads * 64, d->vt_ld, (uint32_t)a.bk)) return -1;
  } else {
    if (make_tmap_bf16(&a.tmV, d->vt, (uint64_t)kv_batches * d->heads * 64, d->kv_rows_per_batch, d->vt_ld, 64)) return -1;
  }
  static bool configured[E2B_MAX_DEVICES] = {false};
  bool& conf = configured[e2b_device_slot()];
  if (!conf) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) { e2b_set_kernel_error("attention smem attribute: %s", cudaGetErrorString(e)); return -1; }
    conf = true;
  }
  const long long items = (long long)((d->q_rows_per_batch + ATT_BQ - 1) / ATT_BQ) * d->heads * d->batch;
  const int sms = e2b_num_sms();
  dim3 grid((unsigned)(items < sms ? items : sms));
  ProfScope ps(stream, "attention", (long long)d->batch * d->q_rows_per_batch, d->kv_rows_per_batch, d->heads,
               4.0 * d->batch * d->heads * (double)d->q_rows_per_batch * d->kv_rows_per_batch * 64.0,
               2.0 * d->batch * d->heads * 64.0 * (2.0 * d->q_rows_per_batch + 2.0 * d->kv_rows_per_batch));
  const bool poly = e2b_attention_poly != 0;
  if (vmn && poly) attention_kernel<true, true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(a);
  else if (vmn) attention_kernel<true, false><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(a);
  else if (poly) attention_kernel<false, true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(a);
  else attention_kernel<false, false><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("attention launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
