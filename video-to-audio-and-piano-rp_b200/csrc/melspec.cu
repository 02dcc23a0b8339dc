// STFT + mel front end (reference MelSpec, e2_tts_crossatt3.py:375-417 + log :293-294): torchaudio MelSpectrogram with
// center=True / reflect padding, periodic Hann window, power=1 (magnitude), HTK mel filterbank without norm, then
// log(clamp(x, 1e-5)).  One CTA transforms FR consecutive frames of one clip: windowed frame -> shared memory
// (bit-reversed) -> in-place radix-2 FFT -> |X| -> filterbank dot products -> log -> [B, n_mels, T] tile store.
#include <math.h>

#include <map>
#include <vector>

#include "../../include/e2b.h"
#include "kernels.h"
#include "prof.h"

namespace e2b {

constexpr int MEL_THREADS = 256;
constexpr int MEL_FR = 8;        // frames per CTA (one 32-byte output segment per mel row)

// First / one-past-last frequency bin with a non-zero weight, per mel filter (triangular filters touch 3-45 of the 513 bins:
// the dense [bins x mels] product of the first version spent most of its FMAs on zeros).  One warp per filter (lanes stride over
// the bins: 17 independent loads per lane instead of 513 dependent ones in one thread, 39 -> ~4 us), run once per call.
__global__ void mel_ranges_kernel(const float* __restrict__ fb, int nbins, int n_mels, int2* __restrict__ range) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= n_mels) return;
  int lo = nbins, hi = 0;
  for (int k = lane; k < nbins; k += 32)
    if (fb[(size_t)k * n_mels + m] != 0.f) { lo = min(lo, k); hi = max(hi, k + 1); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) range[m] = make_int2(lo, hi > lo ? hi : lo);
}

// Two real frames share one complex FFT: z = x1 + i x2 (windowed), Z = FFT(z), and by conjugate symmetry
//   X1[k] = (Z[k] + conj(Z[N-k])) / 2 ,  X2[k] = (Z[k] - conj(Z[N-k])) / (2i)   ->  |X1|, |X2| for k = 0..N/2.
__global__ void __launch_bounds__(MEL_THREADS) melspec_kernel(const float* __restrict__ wav, int nw, int n_fft, int log2n, int hop,
                                                              int n_mels, int T, const float* __restrict__ window,
                                                              const float* __restrict__ fb, const float2* __restrict__ tw,
                                                              const int2* __restrict__ range, float* __restrict__ out, float log_eps) {
  extern __shared__ float sm[];
  float2* z = reinterpret_cast<float2*>(sm);          // n_fft complex
  float* mag = sm + 2 * n_fft;                        // 2 x (n_fft/2 + 1): the two frames of a pair
  const int nbins = n_fft / 2 + 1;
  float* tile = mag + 2 * nbins;                      // n_mels * MEL_FR
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * MEL_FR;
  const float* x = wav + (size_t)b * nw;
  const int half = n_fft / 2;

  for (int fr = 0; fr < MEL_FR; fr += 2) {
    const int t = t0 + fr;
    if (t >= T) break;      // uniform
    const bool two = t + 1 < T;
    const int start = t * hop - half;
    for (int i = threadIdx.x; i < n_fft; i += MEL_THREADS) {
      int s1 = start + i;
      if (s1 < 0) s1 = -s1;
      if (s1 >= nw) s1 = 2 * (nw - 1) - s1;
      int s2 = start + hop + i;
      if (s2 < 0) s2 = -s2;
      if (s2 >= nw) s2 = 2 * (nw - 1) - s2;
      const float w = window[i];
      const int r = __brev((unsigned)i) >> (32 - log2n);
      z[r] = make_float2(x[s1] * w, two ? x[s2] * w : 0.f);
    }
    __syncthreads();
    for (int st = 0; st < log2n; ++st) {
      const int hs = 1 << st;                         // half butterfly span
      for (int j = threadIdx.x; j < half; j += MEL_THREADS) {
        const int grp = j >> st, k = j & (hs - 1);
        const int i0 = (grp << (st + 1)) + k, i1 = i0 + hs;
        const float2 w = tw[k << (log2n - 1 - st)];   // exp(-2 pi i k / (2 hs))
        const float2 a = z[i0], c = z[i1];
        const float2 wc = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
        z[i0] = make_float2(a.x + wc.x, a.y + wc.y);
        z[i1] = make_float2(a.x - wc.x, a.y - wc.y);
      }
      __syncthreads();
    }
    for (int k = threadIdx.x; k < nbins; k += MEL_THREADS) {
      const float2 p = z[k], q = z[(n_fft - k) & (n_fft - 1)];
      const float ar = p.x + q.x, ai = p.y - q.y;     // Z[k] + conj(Z[N-k]) = 2 X1[k]
      const float br = p.x - q.x, bi = p.y + q.y;     // Z[k] - conj(Z[N-k]) = 2i X2[k]
      mag[k] = 0.5f * sqrtf(ar * ar + ai * ai);
      mag[nbins + k] = 0.5f * sqrtf(br * br + bi * bi);
    }
    __syncthreads();
    // filterbank over the non-zero bins of every filter: thread = (frame of the pair, mel)
    for (int i = threadIdx.x; i < 2 * n_mels; i += MEL_THREADS) {
      const int which = i / n_mels, m = i - which * n_mels;
      if (which == 0 || two) {
        const int2 rg = range[m];
        const float* mg = mag + which * nbins;
        float acc = 0.f;
        for (int k = rg.x; k < rg.y; ++k) acc = fmaf(mg[k], __ldg(fb + (size_t)k * n_mels + m), acc);
        tile[m * MEL_FR + fr + which] = logf(fmaxf(acc, log_eps));
      }
    }
    __syncthreads();
  }
  const int nfr = min(MEL_FR, T - t0);
  for (int i = threadIdx.x; i < n_mels * MEL_FR; i += MEL_THREADS) {
    const int m = i / MEL_FR, fr = i % MEL_FR;
    if (fr < nfr) out[((size_t)b * n_mels + m) * T + t0 + fr] = tile[i];
  }
}

// per-device tables: twiddles per n_fft, and a scratch array for the filter ranges (rewritten by every call, on the call's stream)
struct MelTables {
  std::map<int, float2*> twiddles;
  int2* ranges = nullptr;
  int ranges_cap = 0;
};
static MelTables g_mel[E2B_MAX_DEVICES];

static float2* twiddles_for(int n_fft) {
  MelTables& tb = g_mel[e2b_device_slot()];
  auto it = tb.twiddles.find(n_fft);
  if (it != tb.twiddles.end()) return it->second;
  std::vector<float2> h(n_fft / 2);
  for (int k = 0; k < n_fft / 2; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)n_fft;
    h[k] = make_float2((float)cos(a), (float)sin(a));
  }
  float2* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(float2)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  tb.twiddles[n_fft] = d;
  return d;
}

static int2* ranges_for(int n_mels) {
  MelTables& tb = g_mel[e2b_device_slot()];
  if (tb.ranges_cap < n_mels) {
    if (tb.ranges) cudaFree(tb.ranges);
    tb.ranges = nullptr;
    tb.ranges_cap = 0;
    if (cudaMalloc(&tb.ranges, sizeof(int2) * (size_t)n_mels) != cudaSuccess) return nullptr;
    tb.ranges_cap = n_mels;
  }
  return tb.ranges;
}

}  // namespace e2b

using namespace e2b;

extern "C" int e2b_melspec_launch(const float* wav, int B, int nw, int n_fft, int hop, int n_mels, const float* window, const float* fb,
                                  const float* twiddle, float* out, float log_eps, cudaStream_t stream) {
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  if ((1 << log2n) != n_fft || n_fft < 64 || n_fft > 4096) { e2b_set_kernel_error("melspec: n_fft=%d must be a power of two in [64,4096]", n_fft); return -1; }
  if (nw <= n_fft / 2) { e2b_set_kernel_error("melspec: waveform shorter than the reflect padding"); return -1; }
  if (n_mels <= 0 || B <= 0 || hop <= 0) { e2b_set_kernel_error("melspec: bad arguments"); return -1; }
  const int T = nw / hop + 1;
  const size_t smem = (2 * n_fft + 2 * (n_fft / 2 + 1) + n_mels * MEL_FR) * sizeof(float);
  int2* range = ranges_for(n_mels);
  if (!range) { e2b_set_kernel_error("melspec: filter range table allocation failed"); return -1; }
  mel_ranges_kernel<<<(n_mels + 3) / 4, 128, 0, stream>>>(fb, n_fft / 2 + 1, n_mels, range);
  static size_t configured_bytes[E2B_MAX_DEVICES] = {0};
  size_t& configured = configured_bytes[e2b_device_slot()];
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(melspec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      e2b_set_kernel_error("melspec: shared memory request %zu failed", smem);
      return -1;
    }
    configured = smem;
  }
  dim3 grid((T + MEL_FR - 1) / MEL_FR, B);
  ProfScope ps(stream, "melspec", B, nw, n_mels, 0.0, 4.0 * B * ((double)nw + (double)n_mels * T));
  melspec_kernel<<<grid, MEL_THREADS, smem, stream>>>(wav, nw, n_fft, log2n, hop, n_mels, T, window, fb,
                                                      reinterpret_cast<const float2*>(twiddle), range, out, log_eps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("melspec launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

extern "C" int e2b_melspec(const float* wav_dev, int B, int nw, int n_fft, int hop, int n_mels, const float* window_dev,
                           const float* fb_dev, float* out_dev, e2b_stream stream) {
  float2* tw = twiddles_for(n_fft);
  if (!tw) { e2b_set_kernel_error("melspec: twiddle table allocation failed"); return -1; }
  return e2b_melspec_launch(wav_dev, B, nw, n_fft, hop, n_mels, window_dev, fb_dev, reinterpret_cast<const float*>(tw), out_dev, 1e-5f,
                            (cudaStream_t)stream);
}
