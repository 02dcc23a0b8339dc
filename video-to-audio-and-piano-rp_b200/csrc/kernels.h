// Internal kernel-launcher interface of libe2b (host side).  Plain C structs so the same descriptors are exported
// through the kernel-level C-ABI in include/e2b_kernels.h (used by the -m gpu unit tests).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// ---------------------------------------------------------------------------------------------------------------
// GEMM  C[M,N] = A[M,K] * W[N,K]^T  (bf16 x bf16 -> fp32 in TMEM), A given as up to 3 K-concatenated sources.
// ---------------------------------------------------------------------------------------------------------------
enum { E2B_EPI_BF16 = 0, E2B_EPI_F32 = 1, E2B_EPI_GEGLU = 2, E2B_EPI_RESID = 3, E2B_EPI_QKV = 4 };
enum { E2B_MAX_SRC = 9 };

typedef struct e2b_gemm_desc {
  int M, N, K;               // K = sum of ka[]; every ka[] a multiple of 64; N = rows of W
  int num_src;               // up to E2B_MAX_SRC (3 logical sources x (hi, lo, hi) in the error-compensated fp32 mode)
  const void* a[9];          // bf16 row-major sources [M, ka[s]] with leading dimension lda[s] (elements)
  int lda[9];
  int ka[9];
  const void* w;             // bf16 [N, K] row-major (torch nn.Linear layout), leading dimension ldw
  int ldw;
  int epi;                   // E2B_EPI_*
  const float* bias;         // [N] fp32 or NULL (EPI_BF16 / F32 / GEGLU (packed order) / RESID)
  void* out;                 // EPI_BF16/GEGLU/QKV: bf16 ; EPI_F32/RESID: fp32
  int ldo;
  void* out_b16;             // optional bf16 copy of the fp32 result (EPI_F32 / EPI_RESID), same row mapping
  int ldo_b16;
  // EPI_RESID: out = resid + valid(row) * gate[b, col] * (acc + bias)
  const float* resid;
  int ldr;
  const float* gate;         // [*, N] fp32 (already sigmoid-ed) or NULL
  int gate_bstride;          // elements between batches of `gate` (0 = shared by all batches)
  const int* lens;           // per-batch number of valid rows (registers included) or NULL
  int rows_per_batch;        // sequence rows per batch item (RESID mask/gate index, QKV positions)
  // EPI_F32 row remap: out_row = (r / rpb_in) * rpb_out + row_off + r % rpb_in ; add_table[r % rpb_in, :] is added
  int rpb_in, rpb_out, row_off;
  const float* add_table;
  int ld_add;
  // EPI_QKV: packed output columns [0,q_end) = q, [q_end,k_end) = k, [k_end,v_end) = v, [v_end,N) = head gate
  int q_end, k_end, v_end;
  float q_scale;
  const float* rope;         // [positions, 32, 2] (cos, sin) fp32
  int pos_off;               // position = row % rows_per_batch + pos_off
  void* vt;                  // bf16 V^T: [(b*heads_v + h)*64 + d, vt_ld]
  int vt_ld;
  int heads_v;
  float* hgate;              // fp32 [M, hgate_ld] = sigmoid(acc + hgate_bias)
  int hgate_ld;
  const float* hgate_bias;
  // Error-compensated ("fp32") mode.  split > 0: every bf16 output v is stored as the pair hi = bf16(v) at column c and
  // lo = bf16(v - hi) at column c + split (EPI_BF16 / EPI_GEGLU `out`; EPI_F32 / EPI_RESID `out_b16`).
  int split;
  // EPI_QKV in fp32 mode: q,k (RoPE'd, scaled) go to qk_f32 [M, ldo] and v to v_f32 [M, v_f32_ld] instead of bf16 out / vt.
  float* qk_f32;
  float* v_f32;
  int v_f32_ld;
  // EPI_QKV, bf16 mode: v_rowmajor != 0 stores v as plain rows, `vt` = bf16 [M, vt_ld] with column (packed col - k_end), instead of V^T
  int v_rowmajor;
  // EPI_RESID, norm-as-row-scale (TMA epilogue only): row_ss != NULL receives the partial sums of squares of the fp32 result,
  // row_ss[p * row_ss_ld + row] for p < e2b_gemm_row_parts(desc); b16_scale != NULL turns the bf16 copy into bf16(out * scale[col])
  // (rows >= b16_split_row use b16_scale2 when it is set) -- together the A operand and the row statistics of a following RMSNorm
  float* row_ss;
  int row_ss_ld;
  const float* b16_scale;
  const float* b16_scale2;
  int b16_split_row;
  // EPI_GEGLU / EPI_QKV (bf16 mode), the consuming side: in_row_ss != NULL multiplies every accumulator row by
  // in_row_mult / max(sqrt(sum_p in_row_ss[p * in_row_ss_ld + row]), 1e-12) before bias / RoPE / activation -- the RMSNorm of the A
  // operand's rows (in_row_mult = sqrt(C)), whose per-column gain is already inside A (or folded into W)
  const float* in_row_ss;
  int in_row_parts;
  int in_row_ss_ld;
  float in_row_mult;
} e2b_gemm_desc;

int e2b_gemm_launch(const e2b_gemm_desc* d, cudaStream_t stream);
// number of row_ss partials per row an EPI_RESID launch of this description writes (column tiles x epilogue-warp groups)
int e2b_gemm_row_parts(const e2b_gemm_desc* d);
// does an EPI_RESID launch of this description take the TMA-based residual epilogue (which row_ss / b16_scale require)?
int e2b_gemm_resid_uses_tma(const e2b_gemm_desc* d);
// dst[r, c] = src[r, c] * gain[c]
int e2b_scale_cols_launch(const float* src, const float* gain, float* dst, int rows, int cols, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Flash-style attention with tanh soft-clamp, key-length mask and fused per-head gate.
//   q, k: bf16 row-major, head h at columns [col0 + h*64, +64); rows of batch b start at b*rows_per_batch
//   vt  : bf16 V^T [(b*heads + h)*64 + d, vt_ld]   (keys contiguous)
//   out : bf16 [rows, heads*64]
// ---------------------------------------------------------------------------------------------------------------
typedef struct e2b_attn_desc {
  int batch, heads;
  int q_rows_per_batch;      // query rows per batch item (N)
  int kv_rows_per_batch;     // key rows per batch item (N for self, padded nc for cross)
  const void* q; int ldq; int q_col0;
  const void* k; int ldk; int k_col0;
  const void* vt; int vt_ld;
  int kv_batch_mod;          // key/value batch index = b % kv_batch_mod (0 => b); CFG passes share the T5 context
  const int* kv_lens;        // per kv batch: valid keys (prefix), or NULL => kv_rows_per_batch
  int kv_lens_add;           // added to kv_lens[b] (e.g. +32 registers)
  const float* hgate; int hgate_ld;   // [rows, heads] sigmoid gate or NULL
  void* out; int ldo;
  float softclamp;           // 50.0
  // v_rowmajor != 0: `vt` points to V as plain rows instead -- bf16 [kv rows, vt_ld] with head h at columns [v_col0 + h*64, +64),
  // the same layout as k (no transposed copy; the P V product takes it as an MN-major operand)
  int v_rowmajor;
  int v_col0;
} e2b_attn_desc;

int e2b_attention_launch(const e2b_attn_desc* d, cudaStream_t stream);

// Exact fp32 attention for the error-compensated mode (SIMT, one warp per query row): q,k,v fp32 row-major with head h at
// columns [col0 + h*64, +64); same masking / gating semantics; out is a bf16 (hi, lo) pair when out_split > 0.
typedef struct e2b_attn_f32_desc {
  int batch, heads, q_rows_per_batch, kv_rows_per_batch;
  const float* q; int ldq; int q_col0;
  const float* k; int ldk; int k_col0;
  const float* v; int ldv; int v_col0;
  int kv_batch_mod;
  const int* kv_lens; int kv_lens_add;
  const float* hgate; int hgate_ld;
  void* out; int ldo; int out_split;
  float softclamp;
} e2b_attn_f32_desc;
int e2b_attention_f32_launch(const e2b_attn_f32_desc* d, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// element-wise / warp-level kernels (elementwise.cu)
// ---------------------------------------------------------------------------------------------------------------
// y[r_out,:] = x[r,:] / max(||x[r]||, 1e-12) * sqrt(C) * scale[b(r), :]     (scale = g, or gamma+1 per batch)
// rows_per_batch/skip_rows: r = b*rows_per_batch + skip_rows + i  ->  r_out = b*(rows_per_batch-skip_rows) + i
// y is bf16 unless out_f32 != 0
// out_mode: 0 = bf16, 1 = fp32, 2 = bf16 (hi, lo) pair with lo at column + C (ldy must be >= 2C)
int e2b_rmsnorm_launch(const float* x, int ldx, void* y, int ldy, const float* scale, int scale_bstride, int batch,
                       int rows_per_batch, int skip_rows, int C, int out_mode, cudaStream_t stream);

// y = x + mask * silu(dwconv31(mask * x) + bias) over the sequence axis, channels-last [batch, N, C]; w is [K, C]
int e2b_dwconv_launch(const float* x, float* y, const float* w /*[K,C] (transposed conv weight)*/, const float* bias, const int* lens,
                      int batch, int N, int C, int ksize, cudaStream_t stream);

// The same, additionally emitting what the RMSNorm + GEMM that follow need (norm as a row scale): y_b16 [batch*N, C] = bf16(y * gain[c])
// (gain NULL = 1) and row_ss[p * row_ss_ld + row] = sum of y^2 over channels [128 p, 128 p + 128), p < ceil(C / 128).  C % 32 == 0.
int e2b_dwconv_norm_launch(const float* x, float* y, const float* w, const float* bias, const int* lens, int batch, int N, int C, int ksize,
                           void* y_b16, const float* gain, float* row_ss, int row_ss_ld, cudaStream_t stream);

// time conditioning: tcond[s,:] = silu(W1 * [t, sin(2pi t w), cos(2pi t w)] + b1) for nt times;
// then for each of `nmat` matrices  out[s, m, :] = act_m(Wm * tcond[s] + bm)  (act: 0 => x+1 (AdaRMSNorm), 1 => sigmoid)
int e2b_time_mlp_launch(const float* times, int nt, const float* fourier_w, const float* w1, const float* b1, int dim,
                        float* tcond, cudaStream_t stream);
int e2b_time_gemv_launch(const float* tcond, int nt, int dim, const float* const* w, const float* const* b,
                         const int* act, int nmat, float* out, cudaStream_t stream);

// stream init: dst[b, 0:R, :] = registers ; dst[b, R+i, :] = src[b % src_batches, i, :] + add_table[i, :]
// (0 when src == NULL or drop[b]); src_batches < 0 => write the register rows only.  Optional bf16 copy.
// dst_b16 has leading dimension ldb (0 => C); b16_split > 0 stores the (hi, lo) pair with lo at column + b16_split.
int e2b_init_stream_launch(float* dst, void* dst_b16, const float* registers, const float* src, int src_batches,
                           const unsigned char* drop, const float* add_table, int batch, int n, int R, int C,
                           cudaStream_t stream);
int e2b_init_stream_split_launch(float* dst, void* dst_b16, int ldb, int b16_split, const float* registers, const float* src,
                                 int src_batches, const unsigned char* drop, const float* add_table, int batch, int n, int R, int C,
                                 cudaStream_t stream);
int e2b_transpose_launch(const float* src, float* dst, int rows, int cols, cudaStream_t stream);

// fp32 -> bf16 cast with column padding: dst[r, 0:C] = src[r, 0:C], dst[r, C:ldd] = 0
int e2b_cast_pad_launch(const float* src, int lds, void* dst, int ldd, int rows, int C, cudaStream_t stream);
// dst[r, 0:W] for r < rows: part 0 => bf16(src) ("hi"), part 1 => bf16(src - hi) ("lo"); columns C..W are zero.  No padding
// beyond W is written (dst may be a column block of a wider matrix with leading dimension ldd).
int e2b_cast_part_launch(const float* src, int lds, void* dst, int ldd, int rows, int C, int W, int part, cudaStream_t stream);

// guided Euler step.  pred: [P, B, n, d] fp32 (pass 0 = full conditioning, passes 1..P-1 the dropped ones)
//   v = pred0 + sum_k w[k] * (pred0 - pred_k) ; y += dt * v ; optional bf16 copies of the new y for P passes
//   apg != 0 (P == 2 only): cfg update projected orthogonal to pred0 per sample (fp64 dot products)
//   w is a HOST array of P-1 weights
int e2b_guided_euler_launch(float* y, const float* pred, int P, int B, long long per_sample, const float* w, float dt,
                            int apg, float keep_parallel, double* scratch, void* y_b16, int n_copies,
                            cudaStream_t stream);

// the same with the final in-painting select folded in: rows below inpaint_lens[b] of sample b take inpaint[b, row, :]
// (row_elems values per row) instead of the updated state
int e2b_guided_euler_inpaint_launch(float* y, const float* pred, int P, int B, long long per_sample, const float* w, float dt,
                                    int apg, float keep_parallel, double* scratch, void* y_b16, int n_copies,
                                    const float* inpaint, const int* inpaint_lens, int row_elems, cudaStream_t stream);
// dst[b, i, :] = i < lens_dev[b] ? src[b, i, :] : 0
int e2b_mask_rows_launch(const float* src, float* dst, const int* lens_dev, int B, int n, int C, cudaStream_t stream);

// mel front end: wav [B, nw] fp32 -> log-mel [B, n_mels, T], T = nw / hop + 1 (center, reflect pad)
int e2b_melspec_launch(const float* wav, int B, int nw, int n_fft, int hop, int n_mels, const float* window,
                       const float* fb /*[n_fft/2+1, n_mels]*/, const float* twiddle /*[n_fft/2][2]*/, float* out,
                       float log_eps, cudaStream_t stream);

const char* e2b_kernel_last_error(void);
void e2b_set_kernel_error(const char* fmt, ...);

#ifdef __cplusplus
}

// Function attributes (dynamic shared-memory opt-in) and the SM count belong to a DEVICE, not to the process: one slot per
// device ordinal, so a second GPU driven from the same process is configured too.
constexpr int E2B_MAX_DEVICES = 64;
inline int e2b_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < E2B_MAX_DEVICES) ? dev : 0;
}
inline int e2b_num_sms() {
  static int sms[E2B_MAX_DEVICES] = {0};
  const int dev = e2b_device_slot();
  if (!sms[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : 148;
  }
  return sms[dev];
}
#endif
