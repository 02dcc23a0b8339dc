// EnCodec (SEANet) decoder building blocks, SURVEY.md section 8f row N1: the vocoder immediately downstream of the sampler
// (EncodecWrapper.decode, e2_tts_crossatt3.py:434-437 -> transformers EncodecDecoder).  Everything is fp32 and channels-last
// ([batch, time, channels]): the decoder is ~30 GFLOP per 10 s clip against ~69 TFLOP for the sampler, so it is written for
// exactness and streaming access, not for the tensor cores.
//
//   conv1d_cl   causal Conv1d, stride 1:  y[b,t,co] (+)= bias[co] + sum_{k,ci} w[k,ci,co] * act(x[b, t - (K-1) + k, ci])
//               with zero or reflect left padding and an optional ELU on the input.  ConvTranspose1d(kernel 2s, stride s) with
//               the causal right trim is the same kernel with K = 2 and s * Co output channels (phase-major weights): output
//               position t*s + r only receives x[t] * W[:,:,r] and x[t-1] * W[:,:,r+s].
//   lstm_layer  one nn.LSTM layer over a [B,T,4H] tensor of input-side gate pre-activations: persistent kernel, H/4 CTAs, each
//               owns 4 hidden units (its 16 rows of W_hh live in shared memory for all T steps), h_{t-1} is exchanged through a
//               unit-major global buffer and a grid barrier per step.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "kernels.h"
#include "prof.h"

namespace e2b {

constexpr int CV_THREADS = 128;
constexpr int CV_TPT = 16;          // time steps per thread
constexpr int CV_NCO = 2;           // output channels per thread (4 x 8 time steps was slower: 16 weight loads per step)

// threads: co_l = tid % COB (output channel inside the tile), tg = tid / COB (time group); a block covers CV_TPT * (128 / COB)
// time steps of CV_NCO * COB output channels.
__global__ void __launch_bounds__(CV_THREADS, 4) conv1d_cl_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y, int T, int Ci,
                                                              int Co, int K, int ldy, int cob, int flags) {
  extern __shared__ __align__(16) float xs[];          // [(TT + K - 1)][Ci]
  const int ng = CV_THREADS / cob, TT = CV_TPT * ng;
  const int t0 = blockIdx.x * TT, b = blockIdx.z;
  const int rows = TT + K - 1, ci4 = Ci >> 2;
  const float* xb = x + (size_t)b * T * Ci;
  const bool elu = flags & 1, reflect = flags & 4;
  for (int idx = threadIdx.x; idx < rows * ci4; idx += CV_THREADS) {
    const int i = idx / ci4, c4 = idx - i * ci4;
    int ti = t0 + i - (K - 1);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ti < 0 && reflect) ti = -ti;                   // reflect without repeating the edge; beyond the signal: zeros (HF _pad1d)
    if (ti >= 0 && ti < T) v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)ti * Ci) + c4);
    if (elu) {
      v.x = v.x > 0.f ? v.x : expm1f(v.x);
      v.y = v.y > 0.f ? v.y : expm1f(v.y);
      v.z = v.z > 0.f ? v.z : expm1f(v.z);
      v.w = v.w > 0.f ? v.w : expm1f(v.w);
    }
    reinterpret_cast<float4*>(xs)[idx] = v;
  }
  __syncthreads();
  // Each thread produces CV_TPT time steps of CV_NCO output channels (co, co + cob, ...): the broadcast 16-byte shared-memory
  // load of four input values then feeds 4 * CV_NCO FMAs -- with one channel per thread the kernel sat at 20 TFLOP/s, bound by
  // the load/store unit's return path (one LDS.128 per 4 FFMA), not by the FMA pipe.
  const int co_l = threadIdx.x % cob, tg = threadIdx.x / cob;
  const int co = blockIdx.y * CV_NCO * cob + co_l;
  if (co >= Co) return;
  int cc[CV_NCO];
  bool ok[CV_NCO];
#pragma unroll
  for (int c = 0; c < CV_NCO; ++c) {
    ok[c] = co + c * cob < Co;
    cc[c] = ok[c] ? co + c * cob : co;                  // channels beyond Co alias the first one and are not stored
  }
  const int tl0 = tg * CV_TPT;                          // first local time step of this thread
  float acc[CV_NCO][CV_TPT];
  float* yb = y + ((size_t)b * T + t0 + tl0) * ldy;
#pragma unroll
  for (int c = 0; c < CV_NCO; ++c) {
    const float bv = bias ? __ldg(bias + cc[c]) : 0.f;
#pragma unroll
    for (int t = 0; t < CV_TPT; ++t) acc[c][t] = ((flags & 2) && t0 + tl0 + t < T) ? yb[(size_t)t * ldy + cc[c]] + bv : bv;
  }
  // Taps and input channels flatten into one loop: with channels-last rows the K x Ci window of output t is the contiguous run
  // xs[(tl0 + t) * Ci + j], j < K * Ci, and the weights are w[j * Co + co].
  const int KC = K * Ci;
  const float* x0 = xs + (size_t)tl0 * Ci;
  for (int j = 0; j < KC; j += 4) {
    const float* wj = w + (size_t)j * Co;
    float wv[CV_NCO][4];
#pragma unroll
    for (int c = 0; c < CV_NCO; ++c) {
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[c][u] = __ldg(wj + (size_t)u * Co + cc[c]);
    }
#pragma unroll
    for (int t = 0; t < CV_TPT; ++t) {
      const float4 xv = *reinterpret_cast<const float4*>(x0 + (size_t)t * Ci + j);
#pragma unroll
      for (int c = 0; c < CV_NCO; ++c) {
        acc[c][t] = fmaf(wv[c][0], xv.x, acc[c][t]);
        acc[c][t] = fmaf(wv[c][1], xv.y, acc[c][t]);
        acc[c][t] = fmaf(wv[c][2], xv.z, acc[c][t]);
        acc[c][t] = fmaf(wv[c][3], xv.w, acc[c][t]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CV_NCO; ++c) {
    if (!ok[c]) continue;
#pragma unroll
    for (int t = 0; t < CV_TPT; ++t)
      if (t0 + tl0 + t < T) yb[(size_t)t * ldy + cc[c]] = acc[c][t];
  }
}

// ---------------------------------------------------------------------------------------------------- LSTM layer
constexpr int LS_UPC = 4;           // hidden units per CTA
constexpr int LS_THREADS = 256;

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 28)) {      // a protocol bug (or a CTA that never became resident) must fail the launch, not hang the GPU
        printf("e2b: lstm grid barrier timed out (block %d)\n", blockIdx.x);
        __trap();
      }
    } while (v < target);
  }
  __syncthreads();
}

// gx [B,T,4H] = x W_ih^T + b_ih + b_hh (gate order i,f,g,o); whh packed [H/4 ctas][H j][4 units][4 gates]; hbuf 2 x [H][B]
// (unit-major); hseq [B,T,H] = h_t (+ skip[b,t,:]).
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_layer_kernel(const float* __restrict__ gx, const float* __restrict__ whh,
                                                                   const float* __restrict__ skip, float* __restrict__ hseq,
                                                                   float* __restrict__ hbuf, unsigned* __restrict__ counter, int B,
                                                                   int T, int H) {
  extern __shared__ __align__(16) float sm[];
  float4* w_s = reinterpret_cast<float4*>(sm);               // [H][4 units] float4 (4 gates)
  float* h_s = sm + (size_t)H * LS_UPC * 4;                  // [H][B]
  const int u0 = blockIdx.x * LS_UPC;
  for (int i = threadIdx.x; i < H * LS_UPC; i += LS_THREADS) w_s[i] = __ldg(reinterpret_cast<const float4*>(whh) + (size_t)blockIdx.x * H * LS_UPC + i);
  const int pairs = B * LS_UPC;                              // (b, u) pairs of this CTA
  // Few sequences (a single clip = 16 pairs): the 512-long dot products would run on half a warp.  The H range is then split
  // into `nslice` slices over otherwise idle threads and the partial gate sums are combined through shared memory.
  int nslice = 1;
  while (nslice < 16 && 2 * nslice * pairs <= LS_THREADS && H % (2 * nslice) == 0) nslice <<= 1;
  float4* red = reinterpret_cast<float4*>(h_s + (size_t)H * B);          // [nslice][pairs] partial (i, f, g, o)
  constexpr int MAXP = 4;                                    // up to 256 batch items per launch
  float c[MAXP];
#pragma unroll
  for (int q = 0; q < MAXP; ++q) c[q] = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* hprev = hbuf + (size_t)((t + 1) & 1) * H * B;
    float* hcur = hbuf + (size_t)(t & 1) * H * B;
    if (t > 0) {
      for (int i = threadIdx.x; i < H * B / 4; i += LS_THREADS) reinterpret_cast<float4*>(h_s)[i] = __ldcg(reinterpret_cast<const float4*>(hprev) + i);
    }
    __syncthreads();
    auto finish = [&](int u, int b, float ai, float af, float ag, float ao, float& cs) {
      const float ig = 1.f / (1.f + expf(-ai)), fg = 1.f / (1.f + expf(-af)), gg = tanhf(ag), og = 1.f / (1.f + expf(-ao));
      cs = fg * cs + ig * gg;
      const float h = og * tanhf(cs);
      hcur[(size_t)(u0 + u) * B + b] = h;
      const size_t o = ((size_t)b * T + t) * H + u0 + u;
      hseq[o] = skip ? h + __ldg(skip + o) : h;
    };
    if (nslice > 1) {
      const int p = threadIdx.x % pairs, sl = threadIdx.x / pairs;
      const int u = p / B, b = p - u * B;
      float ai = 0.f, af = 0.f, ag = 0.f, ao = 0.f;
      if (sl == 0) {
        const float* g = gx + ((size_t)b * T + t) * 4 * H + u0 + u;
        ai = __ldg(g); af = __ldg(g + H); ag = __ldg(g + 2 * H); ao = __ldg(g + 3 * H);
      }
      if (t > 0 && sl < nslice) {
        const int j1 = (sl + 1) * (H / nslice);
        for (int j = sl * (H / nslice); j < j1; ++j) {
          const float hv = h_s[(size_t)j * B + b];
          const float4 wv = w_s[j * LS_UPC + u];
          ai = fmaf(wv.x, hv, ai);
          af = fmaf(wv.y, hv, af);
          ag = fmaf(wv.z, hv, ag);
          ao = fmaf(wv.w, hv, ao);
        }
        if (sl > 0) red[(size_t)sl * pairs + p] = make_float4(ai, af, ag, ao);
      }
      if (t > 0) __syncthreads();
      if (sl == 0) {
        if (t > 0) {
          for (int k = 1; k < nslice; ++k) {
            const float4 r = red[(size_t)k * pairs + p];
            ai += r.x; af += r.y; ag += r.z; ao += r.w;
          }
        }
        finish(u, b, ai, af, ag, ao, c[0]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < MAXP; ++q) {
        const int p = threadIdx.x + q * LS_THREADS;
        if (p < pairs) {
          const int u = p / B, b = p - u * B;                // consecutive lanes = consecutive batch items
          const float* g = gx + ((size_t)b * T + t) * 4 * H + u0 + u;
          float ai = __ldg(g), af = __ldg(g + H), ag = __ldg(g + 2 * H), ao = __ldg(g + 3 * H);
          if (t > 0) {
            for (int j = 0; j < H; ++j) {
              const float hv = h_s[(size_t)j * B + b];
              const float4 wv = w_s[j * LS_UPC + u];
              ai = fmaf(wv.x, hv, ai);
              af = fmaf(wv.y, hv, af);
              ag = fmaf(wv.z, hv, ag);
              ao = fmaf(wv.w, hv, ao);
            }
          }
          finish(u, b, ai, af, ag, ao, c[q]);
        }
      }
    }
    if (t + 1 < T) grid_barrier(counter, (unsigned)(t + 1) * gridDim.x);
  }
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("%s launch: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // namespace e2b

using namespace e2b;

extern "C" int e2b_conv1d_cl(const float* x_dev, const float* w_dev, const float* bias_dev, float* y_dev, int B, int T, int Ci, int Co, int K,
                             int ldy, int flags, cudaStream_t stream) {
  if (B <= 0 || T <= 0) return 0;
  if (Ci <= 0 || Ci % 4 || Co <= 0 || K <= 0 || ldy < Co) { e2b_set_kernel_error("conv1d_cl: bad shape Ci=%d Co=%d K=%d ldy=%d", Ci, Co, K, ldy); return -1; }
  if (B > 65535) { e2b_set_kernel_error("conv1d_cl: at most 65535 sequences per call"); return -1; }
  int cob = 4;                                        // threads along the channel axis; a block covers CV_NCO * cob output channels
  while (CV_NCO * cob < Co && cob < CV_THREADS) cob <<= 1;
  const int TT = CV_TPT * (CV_THREADS / cob);
  const size_t smem = (size_t)(TT + K - 1) * Ci * sizeof(float);
  if (smem > 200 * 1024) { e2b_set_kernel_error("conv1d_cl: input window of %zu bytes does not fit in shared memory", smem); return -1; }
  static bool configured[E2B_MAX_DEVICES] = {false};
  bool& conf = configured[e2b_device_slot()];
  if (!conf) {
    if (cudaFuncSetAttribute(conv1d_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      e2b_set_kernel_error("conv1d_cl: shared memory attribute failed");
      return -1;
    }
    conf = true;
  }
  ProfScope ps(stream, "conv1d_cl", (long long)B * T, Co, K * Ci, 2.0 * B * T * (double)Co * K * Ci, 4.0 * B * T * ((double)Ci + Co));
  dim3 grid((T + TT - 1) / TT, (Co + CV_NCO * cob - 1) / (CV_NCO * cob), B);
  conv1d_cl_kernel<<<grid, CV_THREADS, smem, stream>>>(x_dev, w_dev, bias_dev, y_dev, T, Ci, Co, K, ldy, cob, flags);
  return check_launch("conv1d_cl");
}

extern "C" int e2b_lstm_layer(const float* gx_dev, const float* whh_packed_dev, const float* skip_dev, float* hseq_dev, float* hbuf_dev,
                              unsigned* counter_dev, int B, int T, int H, cudaStream_t stream) {
  if (B <= 0 || T <= 0) return 0;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (H % LS_UPC || H / LS_UPC > sms) { e2b_set_kernel_error("lstm_layer: hidden size %d needs H %% 4 == 0 and H / 4 <= %d SMs", H, sms); return -1; }
  if (B % 4 || B > 256) { e2b_set_kernel_error("lstm_layer: batch %d must be a multiple of 4, at most 256 per call", B); return -1; }
  const size_t smem = ((size_t)H * LS_UPC * 4 + (size_t)H * B) * sizeof(float) + LS_THREADS * 16 /*partial sums of sliced dot products*/;
  if (smem > 220 * 1024) { e2b_set_kernel_error("lstm_layer: %zu bytes of shared memory needed (reduce the batch per call)", smem); return -1; }
  static bool configured[E2B_MAX_DEVICES] = {false};
  bool& conf = configured[e2b_device_slot()];
  if (!conf) {
    if (cudaFuncSetAttribute(lstm_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
      e2b_set_kernel_error("lstm_layer: shared memory attribute failed");
      return -1;
    }
    conf = true;
  }
  if (cudaMemsetAsync(counter_dev, 0, sizeof(unsigned), stream) != cudaSuccess) { e2b_set_kernel_error("lstm_layer: counter reset failed"); return -1; }
  ProfScope ps(stream, "lstm_layer", (long long)B * T, 4 * H, H, 2.0 * B * T * 4.0 * H * H, 4.0 * B * T * 6.0 * H);
  // every CTA must be resident at once (grid barrier): H / 4 <= SM count and one CTA per SM by its shared-memory footprint
  // (cooperative launch: the runtime refuses the launch instead of letting the barrier deadlock if the CTAs cannot all be resident)
  void* kargs[] = {(void*)&gx_dev, (void*)&whh_packed_dev, (void*)&skip_dev, (void*)&hseq_dev, (void*)&hbuf_dev, (void*)&counter_dev, (void*)&B, (void*)&T, (void*)&H};
  const cudaError_t le = cudaLaunchCooperativeKernel((const void*)lstm_layer_kernel, dim3(H / LS_UPC), dim3(LS_THREADS), kargs, smem, stream);
  if (le != cudaSuccess) { e2b_set_kernel_error("lstm_layer cooperative launch: %s", cudaGetErrorString(le)); return -1; }
  return check_launch("lstm_layer");
}
