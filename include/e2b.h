/* libe2b -- C-ABI of the B200-native CFM sampling path (E2TTS.sample -> E2 Transformer).
 *
 * This is the boundary a reference maintainer binds instead of the PyTorch-eager code in
 * src/e2_tts_pytorch/e2_tts_crossatt3.py ("X3").  Plain pointers and sizes only; every pointer named `dev` is a CUDA
 * device pointer owned by the caller (torch tensors on the host side), `host` pointers are ordinary host memory.
 * All calls are asynchronous on the given stream unless stated; none allocates after e2b_prepare().
 * Every function returns 0 on success and a negative value on error (message: e2b_last_error).
 *
 * Reference interface each entry point replaces:
 *   e2b_create / e2b_load_weights   E2TTS.__init__ + load_state_dict            X3:1275-1523, inference_v2a.py:117-124
 *   e2b_prepare / e2b_set_conditions  the step-invariant part of sample():       X3:2162-2216 (masks, frames_embed,
 *                                     proj_frames X3:2069, T5 context K/V of every attn2 X3:1131, CLIP stream X3:2040)
 *   e2b_set_audio_cond              step_cond / cond_proj_in / final where of the in-painting mode   X3:2029-2035, 2224-2228, 2259-2260
 *   e2b_forward                     transformer_with_pred_head for all guidance passes   X3:1993-2088
 *   e2b_transformer_forward         Transformer.forward                                  X3:941-1143
 *   e2b_guided_euler                cfg_transformer_with_pred_head combine + project + one torchdiffeq Euler update
 *                                   X3:2090-2113, 162-173, 2255
 *   e2b_sample                      the odeint loop of E2TTS.sample                      X3:2221-2256
 *   e2b_melspec                     MelSpec.forward                                      X3:375-417
 *   e2b_conv1d_cl / e2b_lstm_layer  EncodecWrapper.decode -> transformers EncodecDecoder                          X3:434-437
 *   e2b_frame_windows / e2b_roll_expand   E2TTS.encode_frames around the Video2RollNet call                X3:1525-1555
 *   e2b_stage_clip                  E2TTS.encode_video with a feature cache present: nearest-video-frame resampling of the
 *                                   cached CLIP embeddings to the latent frame rate, zero padding           X3:1802-1826
 */
#ifndef E2B_H_
#define E2B_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct e2b_handle e2b_handle;
typedef void* e2b_stream;      /* cudaStream_t */

typedef struct e2b_config {
  int depth, dim, dim_text, dim_frames;
  int heads, dim_head, frames_heads;      /* dim_head must be 64 */
  int num_channels, num_registers, kernel_size, notes, max_seq_len;
  int ff_mult;
  int precision;                          /* 0 = bf16 tensor-core path (rel-L2 <= 1e-2); 1 = error-compensated "fp32" mode
                                             (bf16 hi/lo operand pairs, 3 MMAs per product, exact fp32 attention; <= 1e-4) */
} e2b_config;

typedef struct e2b_tensor {
  const char* name;            /* reference state-dict key, e.g. "transformer.layers.0.0.3.to_q.weight" */
  const float* dev;            /* fp32, contiguous, device */
  int ndim;
  long long shape[4];
} e2b_tensor;

/* pass flags: which conditions a guidance pass drops (pass 0 must be 0 = full conditioning) */
enum { E2B_DROP_CLIP = 1, E2B_DROP_CTX = 2, E2B_DROP_ROLL = 4, E2B_DROP_AUDIO = 8 };

int e2b_create(const e2b_config* cfg, e2b_handle** out);
void e2b_destroy(e2b_handle* h);
const char* e2b_last_error(e2b_handle* h);   /* h may be NULL: last error of the calling thread */

/* Copies + repacks (bf16, fused QKV+gate, GEGLU interleave, transposed conv taps) into library-owned memory. */
int e2b_load_weights(e2b_handle* h, const e2b_tensor* tensors, int n, e2b_stream stream);

/* Workspace for B clips of n frames, nc context tokens (padded length) and P guidance passes (1 + K). */
int e2b_prepare(e2b_handle* h, int B, int n, int nc, int P);

/* Step-invariant inputs.  clip_dev [B,n,dim_text], roll_dev [B,n,notes] (NULL = zeros), ctx_dev [B,nc,dim];
 * lens_host[B] valid frames per clip, ctx_lens_host[B] valid context tokens, pass_flags_host[P]. */
int e2b_set_conditions(e2b_handle* h, const float* clip_dev, const float* roll_dev, const float* ctx_dev,
                       const int* lens_host, const int* ctx_lens_host, const int* pass_flags_host, e2b_stream stream);

/* Audio-conditioned / in-painting mode of sample() (lens < duration; X3:2029-2035 cond_proj_in, 2224-2228 step_cond,
 * 2259-2260 final select; audiocond_snr None).  cond_dev [B,n,num_channels]: the clip's existing latent; cond_lens_host[B]: frames
 * that carry it (cond_mask = lens_to_mask(lens)); audio_drop_host[B] (or NULL): clips whose condition is zeroed in every pass
 * (audio_drop_prompt, X3:2019-2020).  Passes with E2B_DROP_AUDIO see a zero condition (the null pass of the reference's CFG).
 * The network input of every pass gains cond_proj_in(where(cond_mask, cond, 0)) and the last Euler update of e2b_sample writes
 * where(cond_mask, cond, y).  Needs "cond_proj_in.weight" among the loaded tensors; call after e2b_set_conditions, which
 * switches the mode off again (cond_dev = NULL does too). */
int e2b_set_audio_cond(e2b_handle* h, const float* cond_dev, const int* cond_lens_host, const int* audio_drop_host, e2b_stream stream);

/* pred_dev [P,B,n,num_channels] = velocity of every pass at time t for state x_dev [B,n,num_channels]. */
int e2b_forward(e2b_handle* h, const float* x_dev, float t, float* pred_dev, e2b_stream stream);

/* y_dev [B,n,num_channels] is advanced over the grid t_grid_host[0..steps-1] (steps-1 Euler updates):
 *   v = p0 + sum_k w[k] (p0 - p_k)   (apg: the single cfg update is projected orthogonal to p0, keep_parallel) */
int e2b_sample(e2b_handle* h, float* y_dev, const float* t_grid_host, int steps, const float* guidance_w_host,
               int apg, float keep_parallel, e2b_stream stream);

/* Transformer.forward: x [b,n,dim] (already projected), times_host[b], lens_host[b], text_embed [b,n,dim_text],
 * frames_embed [b,n,dim_frames] (already projected), ctx [b,nc,dim] -> out [b,n,dim] fp32.  Uses the workspace of
 * e2b_prepare(B=b, n, nc, P=1). */
int e2b_transformer_forward(e2b_handle* h, const float* x_dev, const float* times_host, const int* lens_host,
                            const float* text_dev, const float* frames_dev, const float* ctx_dev,
                            const int* ctx_lens_host, float* out_dev, e2b_stream stream);

/* One fused guided Euler update on caller-owned buffers (pred_dev [P,B,per_sample]). scratch_dev: 2*B doubles for apg. */
int e2b_guided_euler(float* y_dev, const float* pred_dev, int P, int B, long long per_sample, const float* w_host,
                     float dt, int apg, float keep_parallel, double* scratch_dev, e2b_stream stream);

/* wav_dev [B,nw] -> out_dev [B,n_mels,nw/hop+1] = log(clamp(mel(|STFT|), 1e-5)); tables are caller-provided device
 * arrays: window[n_fft], fb[n_fft/2+1, n_mels] (torchaudio melscale_fbanks layout). */
int e2b_melspec(const float* wav_dev, int B, int nw, int n_fft, int hop, int n_mels, const float* window_dev,
                const float* fb_dev, float* out_dev, e2b_stream stream);

/* Condition staging: out_dev [B,l,d] fp32 = for every clip b and latent frame k < count_b the cached embedding row
 *   j = min(round_half_even((start_b + k*frame_size + frame_size/2) / sampling_rate / (duration_b / (F_b - 1))), F_b - 1)
 * (double arithmetic, the reference's operation order), zeros for k >= count_b.  emb_dev: all clips' embeddings back to back
 * [sum F_b, d] fp32; meta_dev [B,4] int64 = {row offset, F_b, count_b, start_sample_b}; duration_dev [B] double (seconds). */
int e2b_stage_clip(const float* emb_dev, const long long* meta_dev, const double* duration_dev, int B, int l, int d,
                   int sampling_rate, int frame_size, float* out_dev, e2b_stream stream);

/* EnCodec (SEANet) decoder building blocks, fp32, channels-last [B,T,C]; replace the HuggingFace modules behind
 * EncodecWrapper.decode (X3:434-437): EncodecConv1d / EncodecConvTranspose1d / EncodecResnetBlock -> e2b_conv1d_cl,
 * EncodecLSTM -> e2b_lstm_layer.  (video-to-audio-and-piano-rp_b200/e2_tts_pytorch/encodec.py strings them together.)
 *
 * y[b,t,co] (+)= bias[co] + sum_{k<K, ci<Ci} w[(k*Ci + ci)*Co + co] * act(x[b, t-(K-1)+k, ci]);  rows before the sequence are
 * zero (flags & 4: reflected, HF _pad1d);  flags & 1: act = ELU, & 2: accumulate into y.  ldy = row stride of y (>= Co). */
int e2b_conv1d_cl(const float* x_dev, const float* w_dev, const float* bias_dev, float* y_dev, int B, int T, int Ci, int Co, int K,
                  int ldy, int flags, e2b_stream stream);
/* One nn.LSTM layer.  gx_dev [B,T,4H] = x W_ih^T + b_ih + b_hh (gates i,f,g,o); whh_packed_dev [H/4][H][4 units][4 gates];
 * hseq_dev [B,T,H] = h_t (+ skip_dev[b,t,:] when not NULL); hbuf_dev: 2*H*B floats of scratch; counter_dev: one unsigned. */
int e2b_lstm_layer(const float* gx_dev, const float* whh_packed_dev, const float* skip_dev, float* hseq_dev, float* hbuf_dev,
                   unsigned* counter_dev, int B, int T, int H, e2b_stream stream);

/* Piano-roll front end around Video2RollNet (E2TTS.encode_frames, X3:1525-1555).
 * e2b_frame_windows: the network input -- for every frame i of x_dev [b, t, frame_elems] its `window` (5) clamped neighbours
 *   out_dev[(b*t + i), j, :] = x[b, clamp(i + j - window/2, 0, t-1), :]   (replaces the Python double loop at X3:1530-1538).
 * e2b_roll_expand: logits_dev [b*t, notes] (network output) -> roll_dev [b, l, notes] = sigmoid, every row repeated `repeat` (3)
 *   times, cut or zero-padded to l frames (X3:1541-1554). */
int e2b_frame_windows(const float* x_dev, float* out_dev, int b, int t, long long frame_elems, int window, e2b_stream stream);
int e2b_roll_expand(const float* logits_dev, float* roll_dev, int b, int t, int l, int notes, int repeat, e2b_stream stream);

/* sizeof(e2b_config) as this library was compiled: lets a binding check its own struct definition before e2b_create reads it */
int e2b_config_size(void);

/* algorithmic FLOPs of one e2b_forward at the prepared shape as executed (skipped null-pass attn2 not counted) */
double e2b_forward_flops(e2b_handle* h);
/* number of kernels launched by this library since the handle was created */
long long e2b_launch_count(e2b_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* E2B_H_ */
