// libe2b engine: weight repacking, workspace, the 3-stream E2 transformer forward as a sequence of sm_100a kernels,
// and the CFM sampler loop.  Host-side C++ only (no torch); see include/e2b.h for the C-ABI and the reference
// lines each entry replaces.  Layer dataflow follows Transformer.forward, e2_tts_crossatt3.py:941-1143:
//   text -> frames -> cross-condition -> U-Net skip -> conv -> self-attn -> cross-attn(T5) -> GEGLU FF.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/e2b.h"
#include "kernels.h"
#include "prof.h"

typedef __nv_bfloat16 bf16;

namespace {

thread_local char g_api_err[768] = "";

struct Stream3 {   // per-stream (text / frames) layer weights
  float *conv_w = nullptr, *conv_b = nullptr, *g1 = nullptr, *g2 = nullptr, *hg_b = nullptr, *ff1_b = nullptr, *ff2_b = nullptr;
  bf16 *qkv_w = nullptr, *out_w = nullptr, *ff1_w = nullptr, *ff2_w = nullptr;
};
struct LayerW {
  bf16 *skip_w = nullptr, *qkv_w = nullptr, *out_w = nullptr, *q2_w = nullptr, *kv2_w = nullptr, *out2_w = nullptr, *ff1_w = nullptr,
       *ff2_w = nullptr;
  float *conv_w = nullptr, *conv_b = nullptr, *hg_b = nullptr, *hg2_b = nullptr, *ff1_b = nullptr, *ff2_b = nullptr;
  Stream3 t, f;
  bf16 *tfa_w = nullptr, *at_w = nullptr, *af_w = nullptr;
};

}  // namespace

struct e2b_handle {
  e2b_config cfg{};
  int HD = 0, HDt = 0, HDf = 0, inner = 0, inner_t = 0, inner_f = 0, nmat = 0;
  bool weights_loaded = false;
  std::vector<void*> wallocs, sallocs;
  std::string err;
  long long launches = 0;

  // CUDA graph of the whole e2b_sample step loop (all Euler updates: ~380 launches each).  The eager path stays the default
  // for a first call; a second call with the same signature (shapes, grid, guidance, pass flags, state pointer) is captured on
  // the library's own stream and replayed from then on.  E2B_GRAPH=0 disables it.
  cudaStream_t gstream = nullptr;
  cudaEvent_t gev0 = nullptr, gev1 = nullptr;
  cudaGraphExec_t gexec = nullptr;
  std::vector<unsigned char> gkey, gcand;
  long long glaunches = 0;
  int epoch = 0;                   // bumped whenever buffers are (re)allocated: invalidates the graph

  // global weights
  float *registers = nullptr, *t_registers = nullptr, *f_registers = nullptr, *abs_pos = nullptr, *final_g = nullptr;
  float* ones = nullptr;           // [max stream width] of 1.0: the gain of a norm whose g has been folded into the next weight
  float *fourier_w = nullptr, *time_w1 = nullptr, *time_b1 = nullptr;
  float *proj_in_b = nullptr, *to_pred_b = nullptr, *pf_b = nullptr;
  bf16 *proj_in_w = nullptr, *to_pred_w = nullptr, *pf_w = nullptr;
  // audio-conditioned mode (E2TTS(if_cond_proj_in=True), X3:2029-2035): proj_in and cond_proj_in as ONE GEMM over the two
  // K-concatenated sources (state | condition), weights [proj_in.weight | cond_proj_in.weight], summed biases
  bf16* proj_in2_w = nullptr;
  float* proj_in2_b = nullptr;
  float* rope = nullptr;            // [max_pos, 32, 2]
  int rope_rows = 0;
  std::vector<LayerW> L;
  const float** tm_w = nullptr;     // device arrays for the time GEMVs
  const float** tm_b = nullptr;
  int* tm_act = nullptr;

  // prepared shape
  bool f32 = false;                // error-compensated mode: bf16 A operands are (hi, lo) pairs, weights [W_hi|W_hi|W_lo]
  bool v_rows = true;              // bf16 mode: V kept as plain rows [M, H*64] (MN-major operand of P V); false: transposed copy V^T
                                   // (the round-1 layout: E2B_ATTN=v1 or E2B_VT=1)
  int B = 0, n = 0, nc = 0, P = 0, Pctx = 0, N = 0, Bt = 0, Npad = 0, ncpad = 0;
  size_t M = 0;
  bool conditions_set = false;
  float *x[2] = {nullptr, nullptr}, *text[2] = {nullptr, nullptr}, *frames[2] = {nullptr, nullptr};
  bf16 *xb = nullptr, *textb = nullptr, *framesb = nullptr, *xtmpb = nullptr, *nb = nullptr, *qk = nullptr, *vt = nullptr, *ob = nullptr,
       *hb = nullptr, *ybf = nullptr, *finalb = nullptr, *fin = nullptr, *ctxb = nullptr;
  std::vector<bf16*> skipb, k2, vt2;
  float *qkf = nullptr, *vf = nullptr;      // fp32 q/k and v (fp32 mode)
  std::vector<float*> k2f, v2f;
  float* ystate = nullptr;         // the ODE state the captured graph works on (the caller's y is copied in and out)
  float *hg = nullptr, *fr0 = nullptr, *clip = nullptr, *pred = nullptr, *gam = nullptr, *tcond = nullptr, *times_dev = nullptr;
  double* apg_scratch = nullptr;
  int *lens_dev = nullptr, *ctx_lens_dev = nullptr, *cond_lens_dev = nullptr;
  unsigned char* drop_clip_dev = nullptr;
  // norm as a row scale (bf16 mode): the residual GEMM that produces a stream value also emits its row sums of squares (rss) and the
  // bf16 A operand of the next GEMM, which scales its accumulator rows -- no rmsnorm launch between them
  bool fuse_conv = false;          // conv -> norm -> QKV fused as well (the conv kernel emits the normed operand; E2B_FUSE_CONV)
  bool fuse_side = false;          // text / frames FF norm fused (decided at create: bf16 mode and the TMA residual epilogue enabled)
  bool fuse_norm = false;          // audio-stream norms fused too (per call: time conditioning shared by the batch)
  float* rss = nullptr;            // [RSS_PARTS, M]
  bool audio_cond = false;         // in-painting call: condbf / condm / cond_lens_dev are live
  bf16* condbf = nullptr;          // [P, B, n, num_channels] bf16 A operand of cond_proj_in: where(cond_mask, cond, 0), 0 for dropped passes
  float* condm = nullptr;          // [B, n, num_channels] the condition itself, for the final where(cond_mask, cond, out)
  int gam_capacity = 0;
  int pass_flags[8] = {0};

  // The text and frames branches of layer l+1 run on their own streams beside the audio stream of layer l (at small batches their
  // kernels do not fill the GPU -- one 10 s clip is 13 row tiles -- and at large ones partial waves and tails overlap).  Each branch
  // then needs its own scratch set.
  struct Scratch { bf16 *nb = nullptr, *qk = nullptr, *vt = nullptr, *ob = nullptr, *hb = nullptr; float *hg = nullptr, *rss = nullptr; };
  Scratch scr_t, scr_f;
  bool overlap = false;            // decided per prepared shape (allocate_workspace)
  cudaStream_t s_text = nullptr, s_frames = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_tside = nullptr, ev_fside = nullptr, ev_tfa = nullptr, ev_at = nullptr, ev_af = nullptr, ev_audio = nullptr;
};

namespace {

int fail(e2b_handle* h, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_api_err, sizeof(g_api_err), fmt, ap);
  va_end(ap);
  if (h) h->err = g_api_err;
  return -1;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    if ((call) != 0) return fail(h, "%s failed: %s", #call, e2b_kernel_last_error());             \
    ++h->launches;                                                                                \
  } while (0)
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) return fail(h, "%s: %s", #call, cudaGetErrorString(e__));            \
  } while (0)

template <typename T>
int dalloc(e2b_handle* h, std::vector<void*>& pool, T** p, size_t count) {
  void* q = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return fail(h, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(q, 0, bytes);
  if (e != cudaSuccess) return fail(h, "cudaMemset: %s", cudaGetErrorString(e));
  pool.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}
#define DA(pool, ptr, count)                             \
  do {                                                   \
    if (dalloc(h, pool, &(ptr), (size_t)(count))) return -1; \
  } while (0)

void free_pool(std::vector<void*>& pool) {
  for (void* p : pool) cudaFree(p);
  pool.clear();
}

inline int rup(int v, int m) { return (v + m - 1) / m * m; }
constexpr int RSS_PARTS = 16;      // upper bound of e2b_gemm_row_parts for this model family (N <= 2048)

struct WMap {
  std::map<std::string, const e2b_tensor*> m;
  const e2b_tensor* get(const std::string& k) const {
    auto it = m.find(k);
    return it == m.end() ? nullptr : it->second;
  }
};

// ------------------------------------------------------------------------------------------ weight packing helpers
int need(e2b_handle* h, const WMap& w, const std::string& name, const e2b_tensor** out, long long s0, long long s1 = -1, long long s2 = -1) {
  const e2b_tensor* t = w.get(name);
  if (!t) return fail(h, "load_weights: missing tensor '%s'", name.c_str());
  long long want[3] = {s0, s1, s2};
  int nd = s1 < 0 ? 1 : (s2 < 0 ? 2 : 3);
  if (t->ndim != nd) return fail(h, "load_weights: '%s' has ndim %d, expected %d", name.c_str(), t->ndim, nd);
  for (int i = 0; i < nd; ++i)
    if (t->shape[i] != want[i])
      return fail(h, "load_weights: '%s' dim %d is %lld, expected %lld", name.c_str(), i, t->shape[i], want[i]);
  *out = t;
  return 0;
}

int copy_f32(e2b_handle* h, const WMap& w, const std::string& name, float** dst, long long s0, long long s1, cudaStream_t st) {
  const e2b_tensor* t;
  if (need(h, w, name, &t, s0, s1)) return -1;
  const size_t cnt = (size_t)s0 * (s1 < 0 ? 1 : s1);
  DA(h->wallocs, *dst, cnt);
  CU(cudaMemcpyAsync(*dst, t->dev, cnt * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// K-dimension expansion of the error-compensated mode: a block of C input columns (padded to W) becomes [hi | hi | lo].
inline int kx(const e2b_handle* h, int K) { return h->f32 ? 3 * K : K; }

// dst rows [drow0, drow0+rows), columns from dcol0 <- src fp32 rows [srow0, ...), columns [scol0, scol0+C) (zero-padded to W).
// Returns the next free destination column.
int cast_block(e2b_handle* h, const float* src, int lds, int srow0, int scol0, int C, int W, bf16* dst, int ldd, int drow0, int dcol0, int rows,
               cudaStream_t st, int* next_col) {
  const float* sp = src + (size_t)srow0 * lds + scol0;
  bf16* dp = dst + (size_t)drow0 * ldd + dcol0;
  CK(e2b_cast_part_launch(sp, lds, dp, ldd, rows, C, W, 0, st));
  if (h->f32) {
    CK(e2b_cast_part_launch(sp, lds, dp + W, ldd, rows, C, W, 0, st));
    CK(e2b_cast_part_launch(sp, lds, dp + 2 * W, ldd, rows, C, W, 1, st));
  }
  *next_col = dcol0 + kx(h, W);
  return 0;
}

struct KBlock { int C, W; };   // source columns, padded width

// rows [srow0, srow0+rows) of a [*, in_f] fp32 weight -> rows [drow0, ...) of the packed bf16 matrix, K split into blocks
int cast_weight_rows(e2b_handle* h, const float* src, int in_f, int srow0, int rows, bf16* dst, int ldd, int drow0,
                     const std::vector<KBlock>& blocks, cudaStream_t st) {
  int scol = 0, dcol = 0;
  for (const KBlock& kb : blocks) {
    if (cast_block(h, src, in_f, srow0, scol, kb.C, kb.W, dst, ldd, drow0, dcol, rows, st, &dcol)) return -1;
    scol += kb.C;
  }
  return 0;
}

int packed_k(const e2b_handle* h, const std::vector<KBlock>& blocks) {
  int k = 0;
  for (const KBlock& b : blocks) k += kx(h, b.W);
  return k;
}

int pack_linear(e2b_handle* h, const WMap& w, const std::string& name, bf16** dst, int out_f, int in_f, cudaStream_t st,
                std::vector<KBlock> blocks = {}) {
  const e2b_tensor* t;
  if (need(h, w, name, &t, out_f, in_f)) return -1;
  if (blocks.empty()) blocks = {{in_f, in_f}};
  const int ldd = packed_k(h, blocks);
  DA(h->wallocs, *dst, (size_t)out_f * ldd);
  return cast_weight_rows(h, t->dev, in_f, 0, out_f, *dst, ldd, 0, blocks, st);
}

// [Wq; Wk; Wv; Wgate] (or a subset) -> one bf16 matrix
int pack_concat(e2b_handle* h, const WMap& w, const std::vector<std::pair<std::string, int>>& parts, int in_f, bf16** dst, cudaStream_t st) {
  int rows = 0;
  for (auto& p : parts) rows += p.second;
  const std::vector<KBlock> blocks = {{in_f, in_f}};
  const int ldd = packed_k(h, blocks);
  DA(h->wallocs, *dst, (size_t)rows * ldd);
  int r = 0;
  for (auto& p : parts) {
    const e2b_tensor* t;
    if (need(h, w, p.first, &t, p.second, in_f)) return -1;
    if (cast_weight_rows(h, t->dev, in_f, 0, p.second, *dst, ldd, r, blocks, st)) return -1;
    r += p.second;
  }
  return 0;
}

// GEGLU interleave: packed tile of 256 rows = 128 value rows then the matching 128 gate rows.
// col_gain != nullptr: W[:, k] *= gain[k] before the bf16 rounding -- the static RMSNorm gain g of the preceding norm folded into
// the weight, so that the A operand of this GEMM can be the plain bf16 copy of the stream (norm as a row scale)
int pack_geglu(e2b_handle* h, const WMap& w, const std::string& base, int dim, int inner, bf16** wd, float** bd, cudaStream_t st,
               const float* col_gain = nullptr) {
  const e2b_tensor *tw, *tb;
  if (need(h, w, base + ".weight", &tw, 2 * inner, dim) || need(h, w, base + ".bias", &tb, 2 * inner)) return -1;
  if (inner % 128) return fail(h, "GEGLU inner dim %d must be a multiple of 128", inner);
  const float* wsrc = tw->dev;
  float* scaled = nullptr;
  if (col_gain) {
    if (cudaMalloc(&scaled, (size_t)2 * inner * dim * sizeof(float)) != cudaSuccess) return fail(h, "load_weights: temporary for the gain-folded GEGLU weight");
    if (e2b_scale_cols_launch(tw->dev, col_gain, scaled, 2 * inner, dim, st) != 0) { cudaFree(scaled); return fail(h, "scale_cols: %s", e2b_kernel_last_error()); }
    wsrc = scaled;
  }
  struct Free { float* p; cudaStream_t st; ~Free() { if (p) { cudaStreamSynchronize(st); cudaFree(p); } } } free_scaled{scaled, st};
  const std::vector<KBlock> blocks = {{dim, dim}};
  const int ldd = packed_k(h, blocks);
  DA(h->wallocs, *wd, (size_t)2 * inner * ldd);
  DA(h->wallocs, *bd, (size_t)2 * inner);
  for (int t = 0; t < inner / 128; ++t) {
    if (cast_weight_rows(h, wsrc, dim, t * 128, 128, *wd, ldd, t * 256, blocks, st)) return -1;
    if (cast_weight_rows(h, wsrc, dim, inner + t * 128, 128, *wd, ldd, t * 256 + 128, blocks, st)) return -1;
    CU(cudaMemcpyAsync(*bd + t * 256, tb->dev + t * 128, 128 * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(*bd + t * 256 + 128, tb->dev + inner + t * 128, 128 * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int pack_conv(e2b_handle* h, const WMap& w, const std::string& base, int C, float** wd, float** bd, cudaStream_t st) {
  const e2b_tensor* tw;
  const int ks = h->cfg.kernel_size;
  if (need(h, w, base + ".dw_conv1d.0.weight", &tw, C, 1, ks)) return -1;
  DA(h->wallocs, *wd, (size_t)ks * C);
  CK(e2b_transpose_launch(tw->dev, *wd, C, ks, st));
  return copy_f32(h, w, base + ".dw_conv1d.0.bias", bd, C, -1, st);
}

int pack_stream(e2b_handle* h, const WMap& w, const std::string& p, int C, int heads, int inner, Stream3& s, cudaStream_t st) {
  const int HDs = heads * 64;
  if (pack_conv(h, w, p + "0", C, &s.conv_w, &s.conv_b, st)) return -1;
  if (copy_f32(h, w, p + "1.g", &s.g1, C, -1, st)) return -1;
  if (pack_concat(h, w, {{p + "2.to_q.weight", HDs}, {p + "2.to_k.weight", HDs}, {p + "2.to_v.weight", HDs}, {p + "2.to_v_head_gate.weight", heads}},
                  C, &s.qkv_w, st)) return -1;
  if (copy_f32(h, w, p + "2.to_v_head_gate.bias", &s.hg_b, heads, -1, st)) return -1;
  if (pack_linear(h, w, p + "2.to_out.weight", &s.out_w, C, HDs, st)) return -1;
  if (copy_f32(h, w, p + "3.g", &s.g2, C, -1, st)) return -1;
  // bf16 mode: the FF norm of the text / frames streams runs as a row scale, its static gain g is folded into the FF1 weight
  if (pack_geglu(h, w, p + "4.ff.0.proj", C, inner, &s.ff1_w, &s.ff1_b, st, h->fuse_side ? s.g2 : nullptr)) return -1;
  if (pack_linear(h, w, p + "4.ff.2.weight", &s.ff2_w, C, inner, st)) return -1;
  return copy_f32(h, w, p + "4.ff.2.bias", &s.ff2_b, C, -1, st);
}

void free_workspace(e2b_handle* h) {
  free_pool(h->sallocs);
  h->skipb.clear();
  h->k2.clear();
  h->vt2.clear();
  h->k2f.clear();
  h->v2f.clear();
  h->B = h->n = h->nc = h->P = 0;
  h->gam_capacity = 0;
  h->conditions_set = false;
  h->audio_cond = false;
  h->overlap = false;
  h->scr_t = e2b_handle::Scratch();
  h->scr_f = e2b_handle::Scratch();
}

// ------------------------------------------------------------------------------------------ gemm descriptor helpers
// A logical bf16 A operand [rows, C].  In the error-compensated mode it is stored as [rows, 2C] = (hi | lo) and consumed as
// the three K-concatenated sources (hi, lo, hi) against weights packed (W_hi | W_hi | W_lo).
struct Src { const bf16* p; int C; };
inline int ldb(const e2b_handle* h, int C) { return h->f32 ? 2 * C : C; }     // leading dimension of a bf16 A-operand buffer
inline int spl(const e2b_handle* h, int C) { return h->f32 ? C : 0; }         // `split` offset for its producers

e2b_gemm_desc gdesc(const e2b_handle* h, size_t M, int N, std::initializer_list<Src> srcs, const void* w) {
  e2b_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = (int)M; d.N = N;
  int n = 0, K = 0;
  for (const Src& s : srcs) {
    if (!h->f32) {
      d.a[n] = s.p; d.lda[n] = s.C; d.ka[n] = s.C; ++n;
    } else {
      d.a[n] = s.p; d.a[n + 1] = s.p + s.C; d.a[n + 2] = s.p;
      for (int i = 0; i < 3; ++i) { d.lda[n + i] = 2 * s.C; d.ka[n + i] = s.C; }
      n += 3;
    }
    K += kx(h, s.C);
  }
  d.num_src = n; d.K = K; d.w = w; d.ldw = K;
  return d;
}

// fp32 activations -> bf16 A operand (columns zero-padded C -> W)
int cast_act(e2b_handle* h, const float* src, int lds, bf16* dst, int C, int W, size_t rows, cudaStream_t st) {
  if (!h->f32) {
    CK(e2b_cast_pad_launch(src, lds, dst, W, (int)rows, C, st));
  } else {
    CK(e2b_cast_part_launch(src, lds, dst, 2 * W, (int)rows, C, W, 0, st));
    CK(e2b_cast_part_launch(src, lds, dst + W, 2 * W, (int)rows, C, W, 1, st));
  }
  return 0;
}

int norm(e2b_handle* h, const float* x, int C, bf16* y, const float* scale, int bstride, int batch, int skip, cudaStream_t st) {
  CK(e2b_rmsnorm_launch(x, C, y, ldb(h, C), scale, bstride, batch, h->N, skip, C, h->f32 ? 2 : 0, st));
  return 0;
}

struct GamRef { const float* p; int bstride; };

// time tables: gam[s, m, :] for nt times given on the host
int compute_time_tables(e2b_handle* h, const float* times_host, int nt, cudaStream_t st) {
  const int dim = h->cfg.dim;
  if (nt > h->gam_capacity) return fail(h, "time tables: %d times exceed capacity %d", nt, h->gam_capacity);
  CU(cudaMemcpyAsync(h->times_dev, times_host, nt * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(e2b_time_mlp_launch(h->times_dev, nt, h->fourier_w, h->time_w1, h->time_b1, dim, h->tcond, st));
  CK(e2b_time_gemv_launch(h->tcond, nt, dim, h->tm_w, h->tm_b, h->tm_act, h->nmat, h->gam, st));
  return 0;
}

// q/k/v(+gate) projection of `nb` followed by attention; result in h->ob.  kv = nullptr: self attention over `batch`
// sequences; otherwise cross attention to the precomputed context K/V of layer `l`.
int attention_block(e2b_handle* h, int C, int heads, const bf16* w_q, const float* hg_b, int batch, int layer_ctx, cudaStream_t st,
                    int in_parts = 0) {
  const int HDs = heads * 64;
  const bool cross = layer_ctx >= 0;
  const size_t rows = (size_t)batch * h->N;
  const int qcols = cross ? HDs : 2 * HDs;                       // q | k side by side for self attention
  e2b_gemm_desc d = gdesc(h, rows, (cross ? HDs : 3 * HDs) + heads, {{h->nb, C}}, w_q);
  d.epi = E2B_EPI_QKV;
  d.ldo = qcols;
  d.q_end = HDs; d.k_end = cross ? HDs : 2 * HDs; d.v_end = cross ? HDs : 3 * HDs;
  d.q_scale = 0.125f;
  d.rope = h->rope; d.pos_off = 0; d.rows_per_batch = h->N;
  d.hgate = h->hg; d.hgate_ld = heads; d.hgate_bias = hg_b;
  if (in_parts > 0) { d.in_row_ss = h->rss; d.in_row_parts = in_parts; d.in_row_ss_ld = (int)h->M; d.in_row_mult = sqrtf((float)C); }
  if (!h->f32) {
    d.out = h->qk;
    d.vt = h->vt; d.vt_ld = h->v_rows ? HDs : h->Npad; d.heads_v = heads; d.v_rowmajor = h->v_rows;
  } else {
    d.out = h->qkf; d.qk_f32 = h->qkf; d.v_f32 = h->vf; d.v_f32_ld = HDs; d.heads_v = heads;
  }
  CK(e2b_gemm_launch(&d, st));
  if (!h->f32) {
    e2b_attn_desc a;
    memset(&a, 0, sizeof(a));
    a.batch = batch; a.heads = heads; a.q_rows_per_batch = h->N;
    a.q = h->qk; a.ldq = qcols; a.q_col0 = 0;
    if (cross) {
      a.kv_rows_per_batch = h->nc;
      a.k = h->k2[layer_ctx]; a.ldk = HDs; a.k_col0 = 0;
      a.vt = h->vt2[layer_ctx]; a.vt_ld = h->v_rows ? HDs : h->ncpad;
      a.kv_batch_mod = h->B; a.kv_lens = h->ctx_lens_dev;
    } else {
      a.kv_rows_per_batch = h->N;
      a.k = h->qk; a.ldk = qcols; a.k_col0 = HDs;
      a.vt = h->vt; a.vt_ld = h->v_rows ? HDs : h->Npad;
      a.kv_lens = h->lens_dev;
    }
    a.v_rowmajor = h->v_rows; a.v_col0 = 0;
    a.hgate = h->hg; a.hgate_ld = heads;
    a.out = h->ob; a.ldo = HDs;
    a.softclamp = 50.0f;
    CK(e2b_attention_launch(&a, st));
  } else {
    e2b_attn_f32_desc a;
    memset(&a, 0, sizeof(a));
    a.batch = batch; a.heads = heads; a.q_rows_per_batch = h->N;
    a.q = h->qkf; a.ldq = qcols; a.q_col0 = 0;
    if (cross) {
      a.kv_rows_per_batch = h->nc;
      a.k = h->k2f[layer_ctx]; a.ldk = HDs; a.k_col0 = 0;
      a.v = h->v2f[layer_ctx]; a.ldv = HDs; a.v_col0 = 0;
      a.kv_batch_mod = h->B; a.kv_lens = h->ctx_lens_dev;
    } else {
      a.kv_rows_per_batch = h->N;
      a.k = h->qkf; a.ldk = qcols; a.k_col0 = HDs;
      a.v = h->vf; a.ldv = HDs; a.v_col0 = 0;
      a.kv_lens = h->lens_dev;
    }
    a.hgate = h->hg; a.hgate_ld = heads;
    a.out = h->ob; a.ldo = 2 * HDs; a.out_split = HDs;
    a.softclamp = 50.0f;
    CK(e2b_attention_f32_launch(&a, st));
  }
  return 0;
}

int side_stream(e2b_handle* h, float* (&s)[2], bf16* sb, int C, int heads, int inner, const Stream3& w, cudaStream_t st, const char* name) {
  e2b::NvtxRange range(name);
  const int HDs = heads * 64;
  const int conv_parts = (C + 127) / 128;
  if (h->fuse_conv && C % 32 == 0 && conv_parts <= RSS_PARTS) {
    // conv -> RMSNorm -> QKV with the norm as a row scale: the conv kernel also leaves bf16(y * g1) and the row sums of squares
    CK(e2b_dwconv_norm_launch(s[0], s[1], w.conv_w, w.conv_b, h->lens_dev, h->Bt, h->N, C, h->cfg.kernel_size, h->nb, w.g1, h->rss, (int)h->M, st));
    std::swap(s[0], s[1]);
    if (attention_block(h, C, heads, w.qkv_w, w.hg_b, h->Bt, -1, st, conv_parts)) return -1;
  } else {
    CK(e2b_dwconv_launch(s[0], s[1], w.conv_w, w.conv_b, h->lens_dev, h->Bt, h->N, C, h->cfg.kernel_size, st));
    std::swap(s[0], s[1]);
    if (norm(h, s[0], C, h->nb, w.g1, 0, h->Bt, 0, st)) return -1;
    if (attention_block(h, C, heads, w.qkv_w, w.hg_b, h->Bt, -1, st)) return -1;
  }
  int parts = 0;                  // > 0: the out-projection left the FF norm's row statistics and A operand behind
  {
    e2b_gemm_desc d = gdesc(h, h->M, C, {{h->ob, HDs}}, w.out_w);
    d.epi = E2B_EPI_RESID;
    d.out = s[0]; d.ldo = C; d.resid = s[0]; d.ldr = C;
    d.lens = h->lens_dev; d.rows_per_batch = h->N;
    if (h->fuse_side && e2b_gemm_resid_uses_tma(&d) && e2b_gemm_row_parts(&d) <= RSS_PARTS) {
      // norm as a row scale: bf16(x) (the gain g2 lives in the FF1 weight) + row sums of squares, no rmsnorm launch
      d.out_b16 = h->nb; d.ldo_b16 = C;
      d.row_ss = h->rss; d.row_ss_ld = (int)h->M;
      parts = e2b_gemm_row_parts(&d);
    }
    CK(e2b_gemm_launch(&d, st));
  }
  // the gain of this norm is folded into ff1_w whenever fuse_side is on, also when this launch shape took the classic epilogue
  if (!parts && norm(h, s[0], C, h->nb, h->fuse_side ? h->ones : w.g2, 0, h->Bt, 0, st)) return -1;
  {
    e2b_gemm_desc d = gdesc(h, h->M, 2 * inner, {{h->nb, C}}, w.ff1_w);
    d.epi = E2B_EPI_GEGLU; d.bias = w.ff1_b; d.out = h->hb; d.ldo = ldb(h, inner); d.split = spl(h, inner);
    if (parts) { d.in_row_ss = h->rss; d.in_row_parts = parts; d.in_row_ss_ld = (int)h->M; d.in_row_mult = sqrtf((float)C); }
    CK(e2b_gemm_launch(&d, st));
  }
  {
    e2b_gemm_desc d = gdesc(h, h->M, C, {{h->hb, inner}}, w.ff2_w);
    d.epi = E2B_EPI_RESID; d.bias = w.ff2_b;
    d.out = s[0]; d.ldo = C; d.resid = s[0]; d.ldr = C;
    d.out_b16 = sb; d.ldo_b16 = ldb(h, C); d.split = spl(h, C);
    CK(e2b_gemm_launch(&d, st));
  }
  return 0;
}

// The scratch buffers a branch works in: while a ScratchScope with a set is alive, the handle's nb / qk / vt / ob / hb / hg / rss
// point into that set (launches capture the pointers when they are queued; the host queues one branch at a time).
struct ScratchScope {
  e2b_handle* h;
  e2b_handle::Scratch saved;
  bool on;
  ScratchScope(e2b_handle* h_, const e2b_handle::Scratch* set) : h(h_), on(set != nullptr) {
    if (!on) return;
    saved.nb = h->nb; saved.qk = h->qk; saved.vt = h->vt; saved.ob = h->ob; saved.hb = h->hb; saved.hg = h->hg; saved.rss = h->rss;
    h->nb = set->nb; h->qk = set->qk; h->vt = set->vt; h->ob = set->ob; h->hb = set->hb; h->hg = set->hg; h->rss = set->rss;
  }
  ~ScratchScope() {
    if (!on) return;
    h->nb = saved.nb; h->qk = saved.qk; h->vt = saved.vt; h->ob = saved.ob; h->hb = saved.hb; h->hg = saved.hg; h->rss = saved.rss;
  }
};

// Streams x/text/frames (fp32) and xb (bf16 of x) are initialised; gam = time tables for this call.
int forward_core(e2b_handle* h, GamRef gam, cudaStream_t st) {
  const e2b_config& c = h->cfg;
  const int dim = c.dim, dt = c.dim_text, df = c.dim_frames, H = c.heads;
  const int HD = h->HD;
  const size_t M = h->M;
  const int ctx_batch = h->Pctx * h->B;
  const size_t Mc = (size_t)ctx_batch * h->N;
  auto G = [&](int layer, int which) { return gam.p + (size_t)(layer * 6 + which) * dim; };
  // audio-stream norms as row scales: bf16 mode, TMA residual epilogue on, and one time-conditioning vector for the whole batch
  // (e2b_sample / e2b_forward; Transformer.forward with per-item times keeps the rmsnorm kernel)
  const bool fuse = h->fuse_side && gam.bstride == 0;

  // (the per-kernel event profiler wants unperturbed kernel times: one stream while it is on)
  const bool overlap_now = h->overlap && !e2b_prof_is_on_();
  if (overlap_now) {                 // fork: the branch streams start after everything already queued on st
    CU(cudaEventRecord(h->ev_fork, st));
    CU(cudaStreamWaitEvent(h->s_text, h->ev_fork, 0));
    CU(cudaStreamWaitEvent(h->s_frames, h->ev_fork, 0));
  }
  for (int l = 0; l < c.depth; ++l) {
    const LayerW& w = h->L[l];
    e2b::NvtxRange layer_range("e2b.layer");
    // One stream, or (h->overlap, the default in bf16 mode) text / frames branches on their own streams: layer l+1's branches run beside
    // the audio stream of layer l.  Ordering: the cross-condition GEMMs read the PRE-update bf16 copies xb / textb / framesb, so
    //   tfa(l) waits for both branches of layer l;  at(l) / af(l) wait for the audio stream of layer l-1 (xb) and run beside tfa(l);
    //   the branches of layer l+1 overwrite textb / framesb only after tfa(l) has read them;  the audio stream of layer l
    //   overwrites xb only after at(l) / af(l) have read it.
    const bool ov = overlap_now;
    cudaStream_t st_t = ov ? h->s_text : st, st_f = ov ? h->s_frames : st;
    {
      ScratchScope sc(h, ov ? &h->scr_t : nullptr);
      if (side_stream(h, h->text, h->textb, dt, H, h->inner_t, w.t, st_t, "e2b.text_stream")) return -1;
    }
    {
      ScratchScope sc(h, ov ? &h->scr_f : nullptr);
      if (side_stream(h, h->frames, h->framesb, df, c.frames_heads, h->inner_f, w.f, st_f, "e2b.frames_stream")) return -1;
    }
    if (ov) {
      CU(cudaEventRecord(h->ev_tside, st_t));
      CU(cudaEventRecord(h->ev_fside, st_f));
      CU(cudaStreamWaitEvent(st, h->ev_tside, 0));
      CU(cudaStreamWaitEvent(st, h->ev_fside, 0));
    }

    // cross condition (all three read the pre-update bf16 copies)
    bf16* xnew_b = (l < c.depth / 2) ? h->skipb[l] : h->xtmpb;
    {
      e2b::NvtxRange r("e2b.cross_condition");
      {
        e2b_gemm_desc d = gdesc(h, M, dim, {{h->xb, dim}, {h->textb, dt}, {h->framesb, df}}, w.tfa_w);
        d.epi = E2B_EPI_RESID;
        d.out = h->x[0]; d.ldo = dim; d.resid = h->x[0]; d.ldr = dim;
        d.out_b16 = xnew_b; d.ldo_b16 = ldb(h, dim); d.split = spl(h, dim);
        CK(e2b_gemm_launch(&d, st));
      }
      if (ov) CU(cudaEventRecord(h->ev_tfa, st));
      if (w.at_w) {
        if (ov && l > 0) {             // xb of this layer = the audio stream's output of the previous one
          CU(cudaStreamWaitEvent(st_t, h->ev_audio, 0));
          CU(cudaStreamWaitEvent(st_f, h->ev_audio, 0));
        }
        e2b_gemm_desc d = gdesc(h, M, dt, {{h->xb, dim}, {h->textb, dt}}, w.at_w);
        d.epi = E2B_EPI_RESID;
        d.out = h->text[0]; d.ldo = dt; d.resid = h->text[0]; d.ldr = dt;
        CK(e2b_gemm_launch(&d, st_t));
        e2b_gemm_desc e = gdesc(h, M, df, {{h->xb, dim}, {h->framesb, df}}, w.af_w);
        e.epi = E2B_EPI_RESID;
        e.out = h->frames[0]; e.ldo = df; e.resid = h->frames[0]; e.ldr = df;
        CK(e2b_gemm_launch(&e, st_f));
        if (ov) {
          CU(cudaEventRecord(h->ev_at, st_t));
          CU(cudaEventRecord(h->ev_af, st_f));
          CU(cudaStreamWaitEvent(st, h->ev_at, 0));
          CU(cudaStreamWaitEvent(st, h->ev_af, 0));
        }
      }
      if (ov) {
        CU(cudaStreamWaitEvent(st_t, h->ev_tfa, 0));
        CU(cudaStreamWaitEvent(st_f, h->ev_tfa, 0));
      }
    }
    // U-Net skip
    if (l >= c.depth / 2) {
      e2b::NvtxRange r("e2b.unet_skip");
      e2b_gemm_desc d = gdesc(h, M, dim, {{h->xtmpb, dim}, {h->skipb[c.depth - 1 - l], dim}}, w.skip_w);
      d.epi = E2B_EPI_F32;
      d.out = h->x[0]; d.ldo = dim;
      CK(e2b_gemm_launch(&d, st));
    }
    // audio stream
    int parts = 0;                 // > 0: the self-attention out-projection left the next norms' statistics and operands behind
    bool parts_ff_ok = true;
    {
      e2b::NvtxRange r("e2b.audio.self_attn");
      if (fuse && h->fuse_conv && dim % 32 == 0 && (dim + 127) / 128 <= RSS_PARTS) {
        CK(e2b_dwconv_norm_launch(h->x[0], h->x[1], w.conv_w, w.conv_b, h->lens_dev, h->Bt, h->N, dim, c.kernel_size, h->nb, G(l, 0), h->rss, (int)M, st));
        std::swap(h->x[0], h->x[1]);
        if (attention_block(h, dim, H, w.qkv_w, w.hg_b, h->Bt, -1, st, (dim + 127) / 128)) return -1;
      } else {
        CK(e2b_dwconv_launch(h->x[0], h->x[1], w.conv_w, w.conv_b, h->lens_dev, h->Bt, h->N, dim, c.kernel_size, st));
        std::swap(h->x[0], h->x[1]);
        if (norm(h, h->x[0], dim, h->nb, G(l, 0), gam.bstride, h->Bt, 0, st)) return -1;
        if (attention_block(h, dim, H, w.qkv_w, w.hg_b, h->Bt, -1, st)) return -1;
      }
      {
        e2b_gemm_desc d = gdesc(h, M, dim, {{h->ob, HD}}, w.out_w);
        d.epi = E2B_EPI_RESID;
        d.out = h->x[0]; d.ldo = dim; d.resid = h->x[0]; d.ldr = dim;
        d.gate = G(l, 1); d.gate_bstride = gam.bstride;
        d.lens = h->lens_dev; d.rows_per_batch = h->N;
        if (fuse && e2b_gemm_resid_uses_tma(&d) && e2b_gemm_row_parts(&d) <= RSS_PARTS) {
          // The out-projection produces the input of the NEXT norm: rows of context-live passes go on to the cross-attention
          // (AdaptiveRMSNorm gain gamma2 + 1), the others straight to the feed-forward (gamma3 + 1).  It leaves bf16(x * gain) and
          // the row sums of squares; the consuming GEMM scales its accumulator rows by sqrt(dim) / ||x||.
          d.out_b16 = h->nb; d.ldo_b16 = dim;
          d.b16_scale = Mc > 0 ? G(l, 2) : G(l, 4); d.b16_scale2 = G(l, 4); d.b16_split_row = (int)Mc;
          d.row_ss = h->rss; d.row_ss_ld = (int)M;
          parts = e2b_gemm_row_parts(&d);
        }
        CK(e2b_gemm_launch(&d, st));
      }
    }
    // cross attention to the T5 context: only for passes whose context is live (a zero context gives exactly 0)
    if (Mc > 0) {
      e2b::NvtxRange r("e2b.audio.cross_attn");
      if (!parts && norm(h, h->x[0], dim, h->nb, G(l, 2), gam.bstride, ctx_batch, 0, st)) return -1;
      if (attention_block(h, dim, H, w.q2_w, w.hg2_b, ctx_batch, l, st, parts)) return -1;
      e2b_gemm_desc o = gdesc(h, Mc, dim, {{h->ob, HD}}, w.out2_w);
      o.epi = E2B_EPI_RESID;
      o.out = h->x[0]; o.ldo = dim; o.resid = h->x[0]; o.ldr = dim;
      o.gate = G(l, 3); o.gate_bstride = gam.bstride;
      o.lens = h->lens_dev; o.rows_per_batch = h->N;
      if (parts) {
        // these rows now carry the cross-attention update: rewrite their FF-norm operand and statistics.  The launch must take the
        // same epilogue variant as the out-projection (same N and K: same number of partials); otherwise fall back below.
        if (e2b_gemm_resid_uses_tma(&o) && e2b_gemm_row_parts(&o) == parts) {
          o.out_b16 = h->nb; o.ldo_b16 = dim;
          o.b16_scale = G(l, 4);
          o.row_ss = h->rss; o.row_ss_ld = (int)M;
        } else {
          parts_ff_ok = false;
        }
      }
      CK(e2b_gemm_launch(&o, st));
    }
    e2b::NvtxRange ff_range("e2b.audio.ff");
    const bool ff_fused = parts > 0 && parts_ff_ok;
    if (!ff_fused && norm(h, h->x[0], dim, h->nb, G(l, 4), gam.bstride, h->Bt, 0, st)) return -1;
    {
      e2b_gemm_desc d = gdesc(h, M, 2 * h->inner, {{h->nb, dim}}, w.ff1_w);
      d.epi = E2B_EPI_GEGLU; d.bias = w.ff1_b; d.out = h->hb; d.ldo = ldb(h, h->inner); d.split = spl(h, h->inner);
      if (ff_fused) { d.in_row_ss = h->rss; d.in_row_parts = parts; d.in_row_ss_ld = (int)M; d.in_row_mult = sqrtf((float)dim); }
      CK(e2b_gemm_launch(&d, st));
    }
    {
      e2b_gemm_desc d = gdesc(h, M, dim, {{h->hb, h->inner}}, w.ff2_w);
      d.epi = E2B_EPI_RESID; d.bias = w.ff2_b;
      d.out = h->x[0]; d.ldo = dim; d.resid = h->x[0]; d.ldr = dim;
      d.gate = G(l, 5); d.gate_bstride = gam.bstride; d.rows_per_batch = h->N;
      d.out_b16 = h->xb; d.ldo_b16 = ldb(h, dim); d.split = spl(h, dim);
      CK(e2b_gemm_launch(&d, st));
    }
    if (overlap_now) CU(cudaEventRecord(h->ev_audio, st));
  }
  if (overlap_now) {                 // join: the last layer's at / af ran on the branch streams
    CU(cudaEventRecord(h->ev_tside, h->s_text));
    CU(cudaEventRecord(h->ev_fside, h->s_frames));
    CU(cudaStreamWaitEvent(st, h->ev_tside, 0));
    CU(cudaStreamWaitEvent(st, h->ev_fside, 0));
  }
  return 0;
}

// x/text/frames streams from the sampler state: proj_in(y)+abs_pos, CLIP (dropped per pass), precomputed roll stream.
int init_streams_from_state(e2b_handle* h, cudaStream_t st) {
  const e2b_config& c = h->cfg;
  const int R = c.num_registers;
  CK(e2b_init_stream_split_launch(h->x[0], h->xb, ldb(h, c.dim), spl(h, c.dim), h->registers, nullptr, -1, nullptr, nullptr, h->Bt, h->n, R,
                                  c.dim, st));
  e2b_gemm_desc d = h->audio_cond ? gdesc(h, (size_t)h->Bt * h->n, c.dim, {{h->ybf, c.num_channels}, {h->condbf, c.num_channels}}, h->proj_in2_w)
                                  : gdesc(h, (size_t)h->Bt * h->n, c.dim, {{h->ybf, c.num_channels}}, h->proj_in_w);
  d.epi = E2B_EPI_F32; d.bias = h->audio_cond ? h->proj_in2_b : h->proj_in_b;
  d.out = h->x[0]; d.ldo = c.dim; d.out_b16 = h->xb; d.ldo_b16 = ldb(h, c.dim); d.split = spl(h, c.dim);
  d.rpb_in = h->n; d.rpb_out = h->N; d.row_off = R;
  d.add_table = h->abs_pos; d.ld_add = c.dim;
  CK(e2b_gemm_launch(&d, st));
  CK(e2b_init_stream_launch(h->text[0], nullptr, h->t_registers, h->clip, h->B, h->drop_clip_dev, nullptr, h->Bt, h->n, R, c.dim_text, st));
  CU(cudaMemcpyAsync(h->frames[0], h->fr0, h->M * c.dim_frames * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int pred_head(e2b_handle* h, float* pred, cudaStream_t st) {
  const e2b_config& c = h->cfg;
  if (norm(h, h->x[0], c.dim, h->finalb, h->final_g, 0, h->Bt, c.num_registers, st)) return -1;
  e2b_gemm_desc d = gdesc(h, (size_t)h->Bt * h->n, c.num_channels, {{h->finalb, c.dim}}, h->to_pred_w);
  d.epi = E2B_EPI_F32; d.bias = h->to_pred_b; d.out = pred; d.ldo = c.num_channels;
  CK(e2b_gemm_launch(&d, st));
  return 0;
}

int set_ctx(e2b_handle* h, const float* ctx_dev, const int* ctx_lens_host, cudaStream_t st) {
  const e2b_config& c = h->cfg;
  std::vector<int> cl(h->B);
  for (int b = 0; b < h->B; ++b) {
    cl[b] = ctx_lens_host ? ctx_lens_host[b] : h->nc;
    if (cl[b] < 0 || cl[b] > h->nc) return fail(h, "ctx_lens[%d]=%d outside [0,%d]", b, cl[b], h->nc);
  }
  CU(cudaMemcpyAsync(h->ctx_lens_dev, cl.data(), h->B * sizeof(int), cudaMemcpyHostToDevice, st));
  if (cast_act(h, ctx_dev, c.dim, h->ctxb, c.dim, c.dim, (size_t)h->B * h->nc, st)) return -1;
  for (int l = 0; l < c.depth; ++l) {
    // k2 = rope(to_k(ctx)) at positions N-nc..N-1 (x-transformers uses the LAST nc rows of the table), v2 = to_v(ctx)
    e2b_gemm_desc d = gdesc(h, (size_t)h->B * h->nc, 2 * h->HD, {{h->ctxb, c.dim}}, h->L[l].kv2_w);
    d.epi = E2B_EPI_QKV;
    d.ldo = h->HD;
    d.q_end = 0; d.k_end = h->HD; d.v_end = 2 * h->HD;
    d.q_scale = 1.0f;
    d.rope = h->rope; d.pos_off = h->N - h->nc; d.rows_per_batch = h->nc;
    d.heads_v = c.heads;
    if (!h->f32) {
      d.out = h->k2[l];
      d.vt = h->vt2[l]; d.vt_ld = h->v_rows ? h->HD : h->ncpad; d.v_rowmajor = h->v_rows;
    } else {
      d.out = h->k2f[l]; d.qk_f32 = h->k2f[l]; d.v_f32 = h->v2f[l]; d.v_f32_ld = h->HD;
    }
    CK(e2b_gemm_launch(&d, st));
  }
  CU(cudaStreamSynchronize(st));   // cl (host vector) must outlive the async copy
  return 0;
}

int set_lens(e2b_handle* h, const int* lens_host, cudaStream_t st) {
  std::vector<int> lv(h->Bt);
  for (int b = 0; b < h->Bt; ++b) {
    const int l = lens_host ? lens_host[b % h->B] : h->n;
    if (l < 0 || l > h->n) return fail(h, "lens[%d]=%d outside [0,%d]", b % h->B, l, h->n);
    lv[b] = l + h->cfg.num_registers;
  }
  CU(cudaMemcpyAsync(h->lens_dev, lv.data(), h->Bt * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace

// =================================================================================================== C-ABI
extern "C" const char* e2b_last_error(e2b_handle* h) { return h ? h->err.c_str() : g_api_err; }

extern "C" int e2b_create(const e2b_config* cfg, e2b_handle** out) {
  e2b_handle* h = nullptr;
  if (!cfg || !out) return fail(h, "e2b_create: null argument");
  if (cfg->dim_head != 64) return fail(h, "e2b_create: dim_head must be 64 (got %d)", cfg->dim_head);
  if (cfg->depth < 2 || cfg->depth % 2) return fail(h, "e2b_create: depth must be even");
  if (cfg->dim % 64 || cfg->dim_text % 64 || cfg->dim_frames % 64 || cfg->num_channels % 64)
    return fail(h, "e2b_create: dim/dim_text/dim_frames/num_channels must be multiples of 64");
  if (cfg->notes <= 0 || cfg->notes > 64) return fail(h, "e2b_create: notes must be in (0,64]");
  if (cfg->precision != 0 && cfg->precision != 1) return fail(h, "e2b_create: precision must be 0 (bf16) or 1 (error-compensated fp32)");
  if (cfg->kernel_size != 31) return fail(h, "e2b_create: kernel_size must be 31");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(h, "e2b_create: no CUDA device (this library has no CPU path)");
  h = new e2b_handle();
  h->cfg = *cfg;
  h->HD = cfg->heads * 64; h->HDt = cfg->heads * 64; h->HDf = cfg->frames_heads * 64;
  h->inner = cfg->dim * cfg->ff_mult; h->inner_t = cfg->dim_text * cfg->ff_mult; h->inner_f = cfg->dim_frames * cfg->ff_mult;
  h->nmat = cfg->depth * 6;
  h->f32 = cfg->precision == 1;
  {
    // norm as a row scale needs the TMA-based residual epilogue (bf16 mode; E2B_RESID_TMA=0 or E2B_FUSE_NORM=0 switch it off)
    const char *r = getenv("E2B_RESID_TMA"), *f = getenv("E2B_FUSE_NORM");
    h->fuse_side = !h->f32 && !(r && r[0] == '0') && !(f && f[0] == '0');
    const char* cv = getenv("E2B_FUSE_CONV");
    h->fuse_conv = h->fuse_side && !(cv && cv[0] == '0');
  }
  {
    const char *a = getenv("E2B_ATTN"), *v = getenv("E2B_VT");
    const bool v1 = a && (a[0] == 'v' ? a[1] == '1' : a[0] == '1');
    h->v_rows = !(v1 || (v && v[0] == '1'));
  }
  *out = h;
  return 0;
}

static void drop_graph(e2b_handle* h) {
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  h->gexec = nullptr;
  h->gkey.clear();
  h->gcand.clear();
  ++h->epoch;
}

extern "C" void e2b_destroy(e2b_handle* h) {
  if (!h) return;
  drop_graph(h);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->gev0) cudaEventDestroy(h->gev0);
  if (h->gev1) cudaEventDestroy(h->gev1);
  if (h->s_text) cudaStreamDestroy(h->s_text);
  if (h->s_frames) cudaStreamDestroy(h->s_frames);
  for (cudaEvent_t e : {h->ev_fork, h->ev_tside, h->ev_fside, h->ev_tfa, h->ev_at, h->ev_af, h->ev_audio})
    if (e) cudaEventDestroy(e);
  free_workspace(h);
  free_pool(h->wallocs);
  delete h;
}

extern "C" long long e2b_launch_count(e2b_handle* h) { return h ? h->launches : 0; }
extern "C" int e2b_config_size(void) { return (int)sizeof(e2b_config); }

extern "C" int e2b_load_weights(e2b_handle* h, const e2b_tensor* tensors, int n, e2b_stream stream) {
  if (!h) return fail(h, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  const e2b_config& c = h->cfg;
  free_pool(h->wallocs);
  h->weights_loaded = false;
  drop_graph(h);
  WMap w;
  for (int i = 0; i < n; ++i) w.m[tensors[i].name] = &tensors[i];
  const std::string T = "transformer.";
  const int dim = c.dim, dt = c.dim_text, df = c.dim_frames, H = c.heads, HD = h->HD;

  {
    const int cmax = std::max(dim, std::max(dt, df));
    std::vector<float> one(cmax, 1.0f);
    DA(h->wallocs, h->ones, cmax);
    CU(cudaMemcpy(h->ones, one.data(), cmax * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (copy_f32(h, w, T + "registers", &h->registers, c.num_registers, dim, st)) return -1;
  if (copy_f32(h, w, T + "text_registers", &h->t_registers, c.num_registers, dt, st)) return -1;
  if (copy_f32(h, w, T + "frames_registers", &h->f_registers, c.num_registers, df, st)) return -1;
  if (copy_f32(h, w, T + "abs_pos_emb.weight", &h->abs_pos, c.max_seq_len, dim, st)) return -1;
  if (copy_f32(h, w, T + "final_norm.g", &h->final_g, dim, -1, st)) return -1;
  if (copy_f32(h, w, T + "time_cond_mlp.0.weights", &h->fourier_w, dim / 2, -1, st)) return -1;
  if (copy_f32(h, w, T + "time_cond_mlp.1.weight", &h->time_w1, dim, dim + 1, st)) return -1;
  if (copy_f32(h, w, T + "time_cond_mlp.1.bias", &h->time_b1, dim, -1, st)) return -1;
  if (pack_linear(h, w, "proj_in.weight", &h->proj_in_w, dim, c.num_channels, st)) return -1;
  if (copy_f32(h, w, "proj_in.bias", &h->proj_in_b, dim, -1, st)) return -1;
  h->proj_in2_w = nullptr; h->proj_in2_b = nullptr;
  if (w.get("cond_proj_in.weight")) {
    const e2b_tensor *tp, *tc, *tb;
    const int nch = c.num_channels;
    if (need(h, w, "proj_in.weight", &tp, dim, nch) || need(h, w, "cond_proj_in.weight", &tc, dim, nch) || need(h, w, "proj_in.bias", &tb, dim)) return -1;
    const int ldd = 2 * kx(h, nch);
    DA(h->wallocs, h->proj_in2_w, (size_t)dim * ldd);
    int col = 0;
    if (cast_block(h, tp->dev, nch, 0, 0, nch, nch, h->proj_in2_w, ldd, 0, col, dim, st, &col)) return -1;
    if (cast_block(h, tc->dev, nch, 0, 0, nch, nch, h->proj_in2_w, ldd, 0, col, dim, st, &col)) return -1;
    std::vector<float> b0(dim), b1(dim, 0.f);
    CU(cudaMemcpy(b0.data(), tb->dev, dim * sizeof(float), cudaMemcpyDeviceToHost));
    if (const e2b_tensor* tcb = w.get("cond_proj_in.bias")) {          // E2TTS(cond_proj_in_bias=False) has none
      if (need(h, w, "cond_proj_in.bias", &tcb, dim)) return -1;
      CU(cudaMemcpy(b1.data(), tcb->dev, dim * sizeof(float), cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < dim; ++i) b0[i] += b1[i];
    DA(h->wallocs, h->proj_in2_b, dim);
    CU(cudaMemcpy(h->proj_in2_b, b0.data(), dim * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (pack_linear(h, w, "to_pred.weight", &h->to_pred_w, c.num_channels, dim, st)) return -1;
  if (copy_f32(h, w, "to_pred.bias", &h->to_pred_b, c.num_channels, -1, st)) return -1;
  if (pack_linear(h, w, "proj_frames.weight", &h->pf_w, df, c.notes, st, {{c.notes, 64}})) return -1;   // K padded 51 -> 64 with zeros
  if (copy_f32(h, w, "proj_frames.bias", &h->pf_b, df, -1, st)) return -1;

  // RoPE table from the (shared) inv_freq buffer: cos/sin(pos * inv_freq[j]) in fp32 like the reference
  {
    const e2b_tensor* t;
    if (need(h, w, T + "rotary_emb.inv_freq", &t, 32)) return -1;
    float inv[32];
    CU(cudaMemcpy(inv, t->dev, sizeof(inv), cudaMemcpyDeviceToHost));
    h->rope_rows = c.max_seq_len + c.num_registers;
    std::vector<float> tab((size_t)h->rope_rows * 64);
    for (int p = 0; p < h->rope_rows; ++p)
      for (int j = 0; j < 32; ++j) {
        const float f = (float)p * inv[j];
        tab[((size_t)p * 32 + j) * 2] = cosf(f);
        tab[((size_t)p * 32 + j) * 2 + 1] = sinf(f);
      }
    DA(h->wallocs, h->rope, tab.size());
    CU(cudaMemcpy(h->rope, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  }

  h->L.assign(c.depth, LayerW());
  std::vector<const float*> tmw(h->nmat), tmb(h->nmat);
  std::vector<int> tma(h->nmat);
  for (int l = 0; l < c.depth; ++l) {
    LayerW& Lw = h->L[l];
    const std::string a = T + "layers." + std::to_string(l) + ".0.";
    const std::string x = T + "layers." + std::to_string(l) + ".1.";
    const std::string f = T + "layers." + std::to_string(l) + ".2.";
    if (l >= c.depth / 2 && pack_linear(h, w, a + "0.weight", &Lw.skip_w, dim, 2 * dim, st, {{dim, dim}, {dim, dim}})) return -1;
    if (pack_conv(h, w, a + "1", dim, &Lw.conv_w, &Lw.conv_b, st)) return -1;
    if (pack_concat(h, w, {{a + "3.to_q.weight", HD}, {a + "3.to_k.weight", HD}, {a + "3.to_v.weight", HD}, {a + "3.to_v_head_gate.weight", H}}, dim,
                    &Lw.qkv_w, st)) return -1;
    if (copy_f32(h, w, a + "3.to_v_head_gate.bias", &Lw.hg_b, H, -1, st)) return -1;
    if (pack_linear(h, w, a + "3.to_out.weight", &Lw.out_w, dim, HD, st)) return -1;
    if (pack_concat(h, w, {{a + "6.to_q.weight", HD}, {a + "6.to_v_head_gate.weight", H}}, dim, &Lw.q2_w, st)) return -1;
    if (copy_f32(h, w, a + "6.to_v_head_gate.bias", &Lw.hg2_b, H, -1, st)) return -1;
    if (pack_concat(h, w, {{a + "6.to_k.weight", HD}, {a + "6.to_v.weight", HD}}, dim, &Lw.kv2_w, st)) return -1;
    if (pack_linear(h, w, a + "6.to_out.weight", &Lw.out2_w, dim, HD, st)) return -1;
    if (pack_geglu(h, w, a + "9.ff.0.proj", dim, h->inner, &Lw.ff1_w, &Lw.ff1_b, st)) return -1;
    if (pack_linear(h, w, a + "9.ff.2.weight", &Lw.ff2_w, dim, h->inner, st)) return -1;
    if (copy_f32(h, w, a + "9.ff.2.bias", &Lw.ff2_b, dim, -1, st)) return -1;
    // time-conditioning matrices stay fp32: 0 norm1, 1 adaln1, 2 norm2, 3 adaln2, 4 norm3, 5 adaln3
    const int idx[6] = {2, 4, 5, 7, 8, 10};
    for (int k = 0; k < 6; ++k) {
      float *wp = nullptr, *bp = nullptr;
      const std::string base = a + std::to_string(idx[k]) + ".to_gamma";
      if (copy_f32(h, w, base + ".weight", &wp, dim, dim, st)) return -1;
      if (k % 2 == 1 && copy_f32(h, w, base + ".bias", &bp, dim, -1, st)) return -1;
      tmw[l * 6 + k] = wp; tmb[l * 6 + k] = bp; tma[l * 6 + k] = k % 2;
    }
    if (pack_stream(h, w, x, dt, H, h->inner_t, Lw.t, st)) return -1;
    if (pack_stream(h, w, f, df, c.frames_heads, h->inner_f, Lw.f, st)) return -1;
    if (pack_linear(h, w, x + "5.text_frames_to_audio.weight", &Lw.tfa_w, dim, dim + dt + df, st, {{dim, dim}, {dt, dt}, {df, df}})) return -1;
    if (l < c.depth - 1) {
      if (pack_linear(h, w, x + "5.audio_to_text.weight", &Lw.at_w, dt, dim + dt, st, {{dim, dim}, {dt, dt}})) return -1;
      if (pack_linear(h, w, x + "5.audio_to_frames.weight", &Lw.af_w, df, dim + df, st, {{dim, dim}, {df, df}})) return -1;
    }
  }
  DA(h->wallocs, h->tm_w, h->nmat);
  DA(h->wallocs, h->tm_b, h->nmat);
  DA(h->wallocs, h->tm_act, h->nmat);
  CU(cudaMemcpy(h->tm_w, tmw.data(), h->nmat * sizeof(float*), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->tm_b, tmb.data(), h->nmat * sizeof(float*), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->tm_act, tma.data(), h->nmat * sizeof(int), cudaMemcpyHostToDevice));
  CU(cudaStreamSynchronize(st));
  h->weights_loaded = true;
  return 0;
}

static int allocate_workspace(e2b_handle* h, int B, int n, int nc, int P);

extern "C" int e2b_prepare(e2b_handle* h, int B, int n, int nc, int P) {
  if (!h) return fail(h, "null handle");
  const e2b_config& c = h->cfg;
  if (B <= 0 || n <= 0 || nc <= 0 || P < 1 || P > 8) return fail(h, "e2b_prepare: bad shape B=%d n=%d nc=%d P=%d", B, n, nc, P);
  if (n > c.max_seq_len) return fail(h, "e2b_prepare: n=%d exceeds max_seq_len=%d", n, c.max_seq_len);   // X3:958
  if (nc > n + c.num_registers) return fail(h, "e2b_prepare: nc=%d exceeds sequence length %d", nc, n + c.num_registers);
  if (B == h->B && n == h->n && nc == h->nc && P == h->P) return 0;
  drop_graph(h);
  free_workspace(h);
  // The shape is recorded only once every buffer exists: a failed allocation leaves the handle unprepared (B = 0), so a
  // retry with the same shape allocates again instead of taking the early-out above with null / partial buffers.
  if (allocate_workspace(h, B, n, nc, P)) {
    const std::string msg = h->err;
    free_workspace(h);
    h->err = msg;
    return -1;
  }
  return 0;
}

static int allocate_workspace(e2b_handle* h, int B, int n, int nc, int P) {
  const e2b_config& c = h->cfg;
  h->n = n; h->nc = nc; h->P = P; h->Pctx = P;
  h->N = n + c.num_registers; h->Bt = B * P; h->M = (size_t)h->Bt * h->N;
  h->Npad = rup(h->N, 8); h->ncpad = rup(nc, 8);
  const size_t M = h->M;
  const int dim = c.dim, dt = c.dim_text, df = c.dim_frames;
  const int Cmax = std::max(dim, std::max(dt, df));
  const int HDmax = std::max(h->HD, h->HDf), Hmax = std::max(c.heads, c.frames_heads);
  const int inner_max = std::max(h->inner, std::max(h->inner_t, h->inner_f));
  for (int i = 0; i < 2; ++i) {
    DA(h->sallocs, h->x[i], M * dim);
    DA(h->sallocs, h->text[i], M * dt);
    DA(h->sallocs, h->frames[i], M * df);
  }
  const size_t w2 = h->f32 ? 2 : 1;      // bf16 A operands are (hi, lo) pairs in the error-compensated mode
  DA(h->sallocs, h->xb, M * dim * w2);
  DA(h->sallocs, h->textb, M * dt * w2);
  DA(h->sallocs, h->framesb, M * df * w2);
  DA(h->sallocs, h->xtmpb, M * dim * w2);
  h->skipb.resize(c.depth / 2);
  for (auto& p : h->skipb) DA(h->sallocs, p, M * dim * w2);
  DA(h->sallocs, h->nb, M * Cmax * w2);
  DA(h->sallocs, h->hg, M * Hmax);
  DA(h->sallocs, h->ob, M * HDmax * w2);
  DA(h->sallocs, h->hb, M * inner_max * w2);
  DA(h->sallocs, h->ybf, (size_t)h->Bt * n * c.num_channels * w2);
  DA(h->sallocs, h->finalb, (size_t)h->Bt * n * dim * w2);
  DA(h->sallocs, h->fin, (size_t)h->Bt * n * 64 * w2);
  DA(h->sallocs, h->ctxb, (size_t)B * nc * dim * w2);
  if (!h->f32) {
    DA(h->sallocs, h->qk, M * 2 * HDmax);
    DA(h->sallocs, h->vt, (size_t)h->Bt * Hmax * 64 * h->Npad);
    h->k2.resize(c.depth);
    h->vt2.resize(c.depth);
    for (int l = 0; l < c.depth; ++l) {
      DA(h->sallocs, h->k2[l], (size_t)B * nc * h->HD);
      DA(h->sallocs, h->vt2[l], (size_t)B * c.heads * 64 * h->ncpad);
    }
  } else {
    DA(h->sallocs, h->qkf, M * 2 * HDmax);
    DA(h->sallocs, h->vf, M * HDmax);
    h->k2f.resize(c.depth);
    h->v2f.resize(c.depth);
    for (int l = 0; l < c.depth; ++l) {
      DA(h->sallocs, h->k2f[l], (size_t)B * nc * h->HD);
      DA(h->sallocs, h->v2f[l], (size_t)B * nc * h->HD);
    }
  }
  DA(h->sallocs, h->fr0, M * df);
  DA(h->sallocs, h->clip, (size_t)B * n * dt);
  DA(h->sallocs, h->pred, (size_t)h->Bt * n * c.num_channels);
  DA(h->sallocs, h->ystate, (size_t)B * n * c.num_channels);
  h->gam_capacity = std::max(1024, h->Bt);
  DA(h->sallocs, h->gam, (size_t)h->gam_capacity * h->nmat * dim);
  DA(h->sallocs, h->tcond, (size_t)h->gam_capacity * dim);
  DA(h->sallocs, h->times_dev, h->gam_capacity);
  DA(h->sallocs, h->apg_scratch, 2 * B);
  DA(h->sallocs, h->lens_dev, h->Bt);
  DA(h->sallocs, h->ctx_lens_dev, B);
  DA(h->sallocs, h->drop_clip_dev, h->Bt);
  DA(h->sallocs, h->cond_lens_dev, B);
  DA(h->sallocs, h->rss, (size_t)RSS_PARTS * M);
  DA(h->sallocs, h->condbf, (size_t)h->Bt * n * c.num_channels * w2);
  DA(h->sallocs, h->condm, (size_t)B * n * c.num_channels);
  {
    // branch overlap (E2B_OVERLAP_ROWS: largest row count that takes it, 0 = never).  One clip per call: sample() 162 -> 127 ms; it
    // still pays at the C2 batch (partial waves and kernel tails of one branch are filled by another: 130.5 -> 132.3 audio-s/s on the
    // same box, 125.0 -> 128.2 at 16 clips), so the default is every shape; the two extra scratch sets cost 30 KB per row (3 GB at C2).
    const char* e = getenv("E2B_OVERLAP_ROWS");
    const long long max_rows = e ? atoll(e) : (1ll << 40);
    h->overlap = !h->f32 && h->v_rows && (long long)M <= max_rows;
    if (h->overlap) {
      struct { e2b_handle::Scratch* s; int C, HDs, heads, inner; } sets[2] = {{&h->scr_t, dt, h->HDt, c.heads, h->inner_t},
                                                                               {&h->scr_f, df, h->HDf, c.frames_heads, h->inner_f}};
      for (auto& t : sets) {
        DA(h->sallocs, t.s->nb, M * t.C);
        DA(h->sallocs, t.s->qk, M * 2 * t.HDs);
        DA(h->sallocs, t.s->vt, M * t.HDs);
        DA(h->sallocs, t.s->ob, M * t.HDs);
        DA(h->sallocs, t.s->hb, M * t.inner);
        DA(h->sallocs, t.s->hg, M * t.heads);
        DA(h->sallocs, t.s->rss, (size_t)RSS_PARTS * M);
      }
      if (!h->s_text) {
        CU(cudaStreamCreateWithFlags(&h->s_text, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->s_frames, cudaStreamNonBlocking));
        for (cudaEvent_t* ev : {&h->ev_fork, &h->ev_tside, &h->ev_fside, &h->ev_tfa, &h->ev_at, &h->ev_af, &h->ev_audio})
          CU(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
      }
    }
  }
  h->B = B;                        // last: marks the workspace as complete
  return 0;
}

extern "C" int e2b_set_conditions(e2b_handle* h, const float* clip_dev, const float* roll_dev, const float* ctx_dev, const int* lens_host,
                                  const int* ctx_lens_host, const int* pass_flags_host, e2b_stream stream) {
  if (!h) return fail(h, "null handle");
  if (!h->weights_loaded) return fail(h, "e2b_set_conditions: weights not loaded");
  if (!h->B) return fail(h, "e2b_set_conditions: call e2b_prepare first");
  if (!clip_dev || !ctx_dev) return fail(h, "e2b_set_conditions: clip and ctx are required");
  cudaStream_t st = (cudaStream_t)stream;
  const e2b_config& c = h->cfg;
  const int R = c.num_registers;
  // passes: pass 0 = full conditioning; context-live passes must come first (their cross-attention is batched)
  int pctx = 0;
  bool seen_dropped = false;
  std::vector<unsigned char> dc(h->Bt);
  for (int p = 0; p < h->P; ++p) {
    const int fl = pass_flags_host ? pass_flags_host[p] : (p == 0 ? 0 : (E2B_DROP_CLIP | E2B_DROP_CTX | E2B_DROP_AUDIO));
    if (p == 0 && fl != 0) return fail(h, "pass 0 must keep every condition");
    if (fl & E2B_DROP_CTX) seen_dropped = true;
    else {
      if (seen_dropped) return fail(h, "passes that keep the T5 context must precede passes that drop it");
      ++pctx;
    }
    h->pass_flags[p] = fl;
    for (int b = 0; b < h->B; ++b) dc[p * h->B + b] = (fl & E2B_DROP_CLIP) ? 1 : 0;
  }
  h->Pctx = pctx;
  CU(cudaMemcpyAsync(h->drop_clip_dev, dc.data(), h->Bt, cudaMemcpyHostToDevice, st));
  if (set_lens(h, lens_host, st)) return -1;
  CU(cudaMemcpyAsync(h->clip, clip_dev, (size_t)h->B * h->n * c.dim_text * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // piano-roll stream: proj_frames (K padded to 64) + registers, per pass (roll dropped => proj_frames(0) = bias)
  const size_t per_pass = (size_t)h->B * h->n;
  for (int p = 0; p < h->P; ++p) {
    bf16* dst = h->fin + p * per_pass * ldb(h, 64);
    if (roll_dev && !(h->pass_flags[p] & E2B_DROP_ROLL)) {
      if (cast_act(h, roll_dev, c.notes, dst, c.notes, 64, per_pass, st)) return -1;
    } else {
      CU(cudaMemsetAsync(dst, 0, per_pass * ldb(h, 64) * sizeof(bf16), st));
    }
  }
  CK(e2b_init_stream_launch(h->fr0, nullptr, h->f_registers, nullptr, -1, nullptr, nullptr, h->Bt, h->n, R, c.dim_frames, st));
  {
    e2b_gemm_desc d = gdesc(h, (size_t)h->Bt * h->n, c.dim_frames, {{h->fin, 64}}, h->pf_w);
    d.epi = E2B_EPI_F32; d.bias = h->pf_b;
    d.out = h->fr0; d.ldo = c.dim_frames;
    d.rpb_in = h->n; d.rpb_out = h->N; d.row_off = R;
    CK(e2b_gemm_launch(&d, st));
  }
  if (set_ctx(h, ctx_dev, ctx_lens_host, st)) return -1;
  h->conditions_set = true;
  h->audio_cond = false;           // in-painting is per call: e2b_set_audio_cond after this switches it on
  return 0;
}

extern "C" int e2b_set_audio_cond(e2b_handle* h, const float* cond_dev, const int* cond_lens_host, const int* audio_drop_host, e2b_stream stream) {
  if (!h) return fail(h, "null handle");
  if (!h->conditions_set) return fail(h, "e2b_set_audio_cond: call e2b_set_conditions first");
  if (!cond_dev) { h->audio_cond = false; return 0; }
  if (!h->proj_in2_w) return fail(h, "e2b_set_audio_cond: the loaded weights have no cond_proj_in (E2TTS(if_cond_proj_in=False))");
  if (!cond_lens_host) return fail(h, "e2b_set_audio_cond: cond_lens is required");
  cudaStream_t st = (cudaStream_t)stream;
  const e2b_config& c = h->cfg;
  const size_t per_pass = (size_t)h->B * h->n;
  std::vector<int> cl(h->B), eff(h->B);
  for (int b = 0; b < h->B; ++b) {
    if (cond_lens_host[b] < 0 || cond_lens_host[b] > h->n) return fail(h, "cond_lens[%d]=%d outside [0,%d]", b, cond_lens_host[b], h->n);
    cl[b] = cond_lens_host[b];
    eff[b] = (audio_drop_host && audio_drop_host[b]) ? 0 : cl[b];      // audio_drop_prompt zeroes the clip's condition (X3:2019-2020)
  }
  // network input: where(cond_mask, cond, 0) with dropped clips zeroed, through h->pred as fp32 scratch -> bf16 per pass
  CU(cudaMemcpyAsync(h->cond_lens_dev, eff.data(), h->B * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(e2b_mask_rows_launch(cond_dev, h->pred, h->cond_lens_dev, h->B, h->n, c.num_channels, st));
  for (int p = 0; p < h->P; ++p) {
    bf16* dst = h->condbf + p * per_pass * ldb(h, c.num_channels);
    if (h->pass_flags[p] & E2B_DROP_AUDIO) {
      CU(cudaMemsetAsync(dst, 0, per_pass * ldb(h, c.num_channels) * sizeof(bf16), st));
    } else if (cast_act(h, h->pred, c.num_channels, dst, c.num_channels, c.num_channels, per_pass, st)) {
      return -1;
    }
  }
  CU(cudaStreamSynchronize(st));   // eff (host vector) must outlive the async copy; cond_lens_dev is rewritten below
  // output select: the condition itself below cond_lens (dropped clips included: the reference selects from the caller's cond)
  CU(cudaMemcpyAsync(h->cond_lens_dev, cl.data(), h->B * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->condm, cond_dev, per_pass * c.num_channels * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CU(cudaStreamSynchronize(st));
  h->audio_cond = true;
  return 0;
}

extern "C" int e2b_forward(e2b_handle* h, const float* x_dev, float t, float* pred_dev, e2b_stream stream) {
  if (!h) return fail(h, "null handle");
  if (!h->conditions_set) return fail(h, "e2b_forward: call e2b_set_conditions first");
  cudaStream_t st = (cudaStream_t)stream;
  const e2b_config& c = h->cfg;
  const size_t per_pass = (size_t)h->B * h->n;
  if (compute_time_tables(h, &t, 1, st)) return -1;
  CU(cudaStreamSynchronize(st));   // &t is a stack address
  for (int p = 0; p < h->P; ++p)
    if (cast_act(h, x_dev, c.num_channels, h->ybf + p * per_pass * ldb(h, c.num_channels), c.num_channels, c.num_channels, per_pass, st)) return -1;
  if (init_streams_from_state(h, st)) return -1;
  if (forward_core(h, GamRef{h->gam, 0}, st)) return -1;
  return pred_head(h, pred_dev, st);
}

// the step loop of e2b_sample (time tables already computed)
static int run_sample_steps(e2b_handle* h, float* y_dev, const float* t_grid_host, int steps, const float* guidance_w_host, int apg,
                            float keep_parallel, cudaStream_t st) {
  const e2b_config& c = h->cfg;
  const size_t per_pass = (size_t)h->B * h->n;
  const long long per_sample = (long long)h->n * c.num_channels;
  for (int p = 0; p < h->P; ++p)
    if (cast_act(h, y_dev, c.num_channels, h->ybf + p * per_pass * ldb(h, c.num_channels), c.num_channels, c.num_channels, per_pass, st)) return -1;
  for (int s = 0; s < steps - 1; ++s) {
    e2b::NvtxRange range("e2b.euler_update");
    const float dt = t_grid_host[s + 1] - t_grid_host[s];
    if (init_streams_from_state(h, st)) return -1;
    if (forward_core(h, GamRef{h->gam + (size_t)s * h->nmat * c.dim, 0}, st)) return -1;
    if (pred_head(h, h->pred, st)) return -1;
    // bf16 mode: the Euler kernel also emits the bf16 copies of y that feed the next proj_in; the error-compensated mode
    // needs (hi, lo) pairs, produced by a separate split cast
    const bool select = h->audio_cond && s == steps - 2;      // out = where(cond_mask, cond, out) folded into the last update
    CK(e2b_guided_euler_inpaint_launch(y_dev, h->pred, h->P, h->B, per_sample, guidance_w_host, dt, apg, keep_parallel, h->apg_scratch,
                                       h->f32 ? nullptr : h->ybf, h->f32 ? 0 : h->P, select ? h->condm : nullptr,
                                       select ? h->cond_lens_dev : nullptr, select ? c.num_channels : 0, st));
    if (h->f32 && s + 1 < steps - 1)
      for (int p = 0; p < h->P; ++p)
        if (cast_act(h, y_dev, c.num_channels, h->ybf + p * per_pass * ldb(h, c.num_channels), c.num_channels, c.num_channels, per_pass, st)) return -1;
  }
  return 0;
}

static bool graphs_enabled() {
  static const bool on = [] {
    const char* e = getenv("E2B_GRAPH");
    return !(e && e[0] == '0');
  }();
  return on;
}

extern "C" int e2b_sample(e2b_handle* h, float* y_dev, const float* t_grid_host, int steps, const float* guidance_w_host, int apg,
                          float keep_parallel, e2b_stream stream) {
  if (!h) return fail(h, "null handle");
  if (!h->conditions_set) return fail(h, "e2b_sample: call e2b_set_conditions first");
  if (steps < 1) return fail(h, "e2b_sample: steps must be >= 1");
  if (steps - 1 > h->gam_capacity) return fail(h, "e2b_sample: too many steps");
  if (h->P > 1 && !guidance_w_host) return fail(h, "e2b_sample: guidance weights required");
  cudaStream_t st = (cudaStream_t)stream;
  if (steps == 1) return 0;
  if (compute_time_tables(h, t_grid_host, steps - 1, st)) return -1;      // host -> device copy of the grid: outside the graph
  if (!graphs_enabled() || e2b_prof_is_on_()) return run_sample_steps(h, y_dev, t_grid_host, steps, guidance_w_host, apg, keep_parallel, st);

  // signature of everything that shapes the launch sequence or is baked into kernel parameters
  std::vector<unsigned char> key;
  auto put = [&](const void* p, size_t n) { key.insert(key.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
  put(&steps, sizeof steps); put(&apg, sizeof apg); put(&keep_parallel, sizeof keep_parallel);
  put(&h->epoch, sizeof h->epoch); put(&h->P, sizeof h->P); put(h->pass_flags, sizeof(int) * h->P);
  put(&h->audio_cond, sizeof h->audio_cond);
  put(t_grid_host, sizeof(float) * steps);
  if (h->P > 1) put(guidance_w_host, sizeof(float) * (h->P - 1));

  if (!h->gexec || key != h->gkey) {
    if (key != h->gcand) {                     // first sighting: run eagerly (this also performs every one-time kernel setup)
      h->gcand = key;
      return run_sample_steps(h, y_dev, t_grid_host, steps, guidance_w_host, apg, keep_parallel, st);
    }
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; h->gkey.clear(); }
    if (!h->gstream) {
      CU(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&h->gev0, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&h->gev1, cudaEventDisableTiming));
    }
    const long long l0 = h->launches;
    CU(cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeThreadLocal));
    const int rc = run_sample_steps(h, h->ystate, t_grid_host, steps, guidance_w_host, apg, keep_parallel, h->gstream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->gstream, &graph);
    h->glaunches = h->launches - l0;
    h->launches = l0;
    if (rc != 0 || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      if (rc != 0) return -1;
      return fail(h, "e2b_sample: graph capture failed: %s", cudaGetErrorString(ce));
    }
    const cudaError_t ie = cudaGraphInstantiate(&h->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { h->gexec = nullptr; return fail(h, "e2b_sample: graph instantiation failed: %s", cudaGetErrorString(ie)); }
    h->gkey = key;
  }
  // run the graph on the library's stream, ordered after / before the caller's stream
  const size_t ybytes = (size_t)h->B * h->n * h->cfg.num_channels * sizeof(float);
  CU(cudaMemcpyAsync(h->ystate, y_dev, ybytes, cudaMemcpyDeviceToDevice, st));
  CU(cudaEventRecord(h->gev0, st));
  CU(cudaStreamWaitEvent(h->gstream, h->gev0, 0));
  CU(cudaGraphLaunch(h->gexec, h->gstream));
  CU(cudaEventRecord(h->gev1, h->gstream));
  CU(cudaStreamWaitEvent(st, h->gev1, 0));
  CU(cudaMemcpyAsync(y_dev, h->ystate, ybytes, cudaMemcpyDeviceToDevice, st));
  h->launches += h->glaunches;
  return 0;
}

extern "C" int e2b_transformer_forward(e2b_handle* h, const float* x_dev, const float* times_host, const int* lens_host,
                                       const float* text_dev, const float* frames_dev, const float* ctx_dev, const int* ctx_lens_host,
                                       float* out_dev, e2b_stream stream) {
  if (!h) return fail(h, "null handle");
  if (!h->weights_loaded || !h->B) return fail(h, "e2b_transformer_forward: load weights and prepare first");
  if (h->P != 1) return fail(h, "e2b_transformer_forward: prepare with P=1");
  if (!x_dev || !times_host || !text_dev || !frames_dev || !ctx_dev) return fail(h, "e2b_transformer_forward: all inputs are required");
  cudaStream_t st = (cudaStream_t)stream;
  const e2b_config& c = h->cfg;
  const int R = c.num_registers;
  h->Pctx = 1;
  if (set_lens(h, lens_host, st)) return -1;
  if (set_ctx(h, ctx_dev, ctx_lens_host, st)) return -1;
  if (compute_time_tables(h, times_host, h->B, st)) return -1;
  CK(e2b_init_stream_split_launch(h->x[0], h->xb, ldb(h, c.dim), spl(h, c.dim), h->registers, x_dev, h->B, nullptr, h->abs_pos, h->B, h->n, R,
                                  c.dim, st));
  CK(e2b_init_stream_launch(h->text[0], nullptr, h->t_registers, text_dev, h->B, nullptr, nullptr, h->B, h->n, R, c.dim_text, st));
  CK(e2b_init_stream_launch(h->frames[0], nullptr, h->f_registers, frames_dev, h->B, nullptr, nullptr, h->B, h->n, R, c.dim_frames, st));
  if (forward_core(h, GamRef{h->gam, h->nmat * c.dim}, st)) return -1;
  CK(e2b_rmsnorm_launch(h->x[0], c.dim, out_dev, c.dim, h->final_g, 0, h->B, h->N, R, c.dim, 1, st));
  h->conditions_set = false;
  return 0;
}

extern "C" int e2b_guided_euler(float* y_dev, const float* pred_dev, int P, int B, long long per_sample, const float* w_host, float dt,
                                int apg, float keep_parallel, double* scratch_dev, e2b_stream stream) {
  e2b_handle* h = nullptr;
  if (e2b_guided_euler_launch(y_dev, pred_dev, P, B, per_sample, w_host, dt, apg, keep_parallel, scratch_dev, nullptr, 0, (cudaStream_t)stream))
    return fail(h, "e2b_guided_euler: %s", e2b_kernel_last_error());
  return 0;
}

extern "C" double e2b_forward_flops(e2b_handle* h) {
  if (!h || !h->B) return 0.0;
  const e2b_config& c = h->cfg;
  const double N = h->N, n = h->n, nc = h->nc;
  const double Bt = h->Bt, Bc = (double)h->Pctx * h->B;
  auto side = [&](double C, double heads, double inner) {
    return 2 * N * C * (3 * heads * 64 + heads) + 4 * N * N * heads * 64 + 2 * N * heads * 64 * C + 2 * N * C * 3 * inner + 2 * N * C * 31;
  };
  double per = 2 * n * (c.num_channels * (double)c.dim + 64.0 * c.dim_frames + (double)c.dim * c.num_channels);
  per += c.depth * (side(c.dim_text, c.heads, h->inner_t) + side(c.dim_frames, c.frames_heads, h->inner_f));
  per += c.depth * 2 * N * (c.dim + c.dim_text + c.dim_frames) * c.dim;
  per += (c.depth - 1) * (2 * N * (c.dim + c.dim_text) * c.dim_text + 2 * N * (c.dim + c.dim_frames) * c.dim_frames);
  per += (c.depth / 2) * 2 * N * 2 * c.dim * c.dim;
  per += c.depth * (side(c.dim, c.heads, h->inner));
  double cross = c.depth * (2 * N * c.dim * (h->HD + c.heads) + 4 * N * nc * h->HD + 2 * N * h->HD * c.dim);
  return Bt * per + Bc * cross;
}
