// Piano-roll front end around the (third-party, torch) Video2RollNet: SURVEY.md 8f row N3, E2TTS.encode_frames
// (e2_tts_crossatt3.py:1525-1555).  The reference builds the network input with a Python loop over every video frame
// (5 clamped neighbours concatenated per frame) and post-processes the logits with sigmoid / repeat / cat; here both sides
// are one HBM-bound kernel each:
//   frame_windows   out[(b*t + i), j, :] = x[b, clamp(i + j - 2, 0, t-1), :]           j = 0..4   (:1531-1538)
//   roll_expand     roll[b, r, :] = r / repeat < t ? sigmoid(logits[b*t + r / repeat, :]) : 0       (:1540-1554)
#include "../../include/e2b.h"
#include "kernels.h"

namespace e2b {

// One CTA copies one (destination frame, window slot): frame_elems fp32 as float4 (100 x 900 = 90 000 floats per frame).
__global__ void __launch_bounds__(256) frame_windows_kernel(const float4* __restrict__ x, float4* __restrict__ out, int t, int win, size_t f4) {
  const int half = win / 2;
  const size_t slot = blockIdx.x;                 // (b * t + i) * win + j
  const int j = (int)(slot % win);
  const size_t bi = slot / win;
  const int i = (int)(bi % t);
  const size_t b = bi / t;
  const int src = min(max(i + j - half, 0), t - 1);
  const float4* s = x + (b * t + src) * f4;
  float4* d = out + slot * f4;
  for (size_t k = (size_t)blockIdx.y * blockDim.x + threadIdx.x; k < f4; k += (size_t)gridDim.y * blockDim.x) d[k] = __ldg(s + k);
}

__global__ void __launch_bounds__(256) roll_expand_kernel(const float* __restrict__ logits, float* __restrict__ roll, int t, int l, int notes,
                                                          int repeat, size_t total) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % notes);
    const size_t row = idx / notes;
    const int r = (int)(row % l);
    const size_t b = row / l;
    const int src = r / repeat;
    // torch.sigmoid in fp32: 1 / (1 + exp(-x))
    roll[idx] = src < t ? 1.0f / (1.0f + expf(-logits[(b * t + src) * notes + c])) : 0.0f;
  }
}

}  // namespace e2b

extern "C" int e2b_frame_windows(const float* x_dev, float* out_dev, int b, int t, long long frame_elems, int window, e2b_stream stream) {
  if (b <= 0 || t <= 0) return 0;
  if (window < 1 || window % 2 == 0) { e2b_set_kernel_error("frame_windows: window must be odd (got %d)", window); return -1; }
  if (frame_elems <= 0 || frame_elems % 4 || (reinterpret_cast<uintptr_t>(x_dev) & 15) || (reinterpret_cast<uintptr_t>(out_dev) & 15)) {
    e2b_set_kernel_error("frame_windows: frames must be 16-byte aligned multiples of 4 floats (frame_elems=%lld)", frame_elems);
    return -1;
  }
  const size_t f4 = (size_t)frame_elems / 4;
  const unsigned gy = (unsigned)((f4 + 256 * 8 - 1) / (256 * 8));
  dim3 grid((unsigned)((size_t)b * t * window), gy ? gy : 1);
  e2b::frame_windows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x_dev), reinterpret_cast<float4*>(out_dev), t,
                                                                    window, f4);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("frame_windows launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

extern "C" int e2b_roll_expand(const float* logits_dev, float* roll_dev, int b, int t, int l, int notes, int repeat, e2b_stream stream) {
  if (b <= 0 || l <= 0) return 0;
  if (t <= 0 || notes <= 0 || repeat <= 0) { e2b_set_kernel_error("roll_expand: bad shape t=%d notes=%d repeat=%d", t, notes, repeat); return -1; }
  const size_t total = (size_t)b * l * notes;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  e2b::roll_expand_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(logits_dev, roll_dev, t, l, notes, repeat, total);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("roll_expand launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
