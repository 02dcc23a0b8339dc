// Condition staging: the per-latent-frame CLIP stream of E2TTS.encode_video
// (e2_tts_crossatt3.py:1802-1826) from cached per-video-frame embeddings.
//
//   out[b, k, :] = emb_b[min(rint((start_b + k * frame_size + frame_size / 2) / sampling_rate / (duration_b / (F_b - 1))), F_b - 1), :]
//                  for k < count_b, zeros after
//
// The index is evaluated in IEEE double with the reference's operation order and round-half-even (Python `round`), so it is
// bit-identical to the reference's; the rows are copied, never interpolated.  HBM-bound gather: 4 B read + 4 B written per
// output element (reads of a repeated frame hit L2).
#include "kernels.h"
#include "prof.h"

namespace e2b {

// meta[b] = {embedding row offset, F, count, start_sample}
__global__ void __launch_bounds__(256) stage_clip_kernel(const float* __restrict__ emb, const long long* __restrict__ meta,
                                                         const double* __restrict__ duration, int l, int d4, int sampling_rate,
                                                         int frame_size, float* __restrict__ out) {
  const int k = blockIdx.x, b = blockIdx.y;
  const long long off = meta[4 * b], nf = meta[4 * b + 1], count = meta[4 * b + 2], start = meta[4 * b + 3];
  float4* dst = reinterpret_cast<float4*>(out) + ((size_t)b * l + k) * d4;
  if (k >= count) {
    for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const long long i = start + (long long)k * frame_size;
  // (i + frame_size // 2) / sampling_rate / (duration / (F - 1)): int/int true division, then double/double
  const double pos = (double)(i + frame_size / 2) / (double)sampling_rate / (duration[b] / (double)(nf - 1));
  long long j = (long long)rint(pos);
  if (j > nf - 1) j = nf - 1;
  const float4* src = reinterpret_cast<const float4*>(emb) + (size_t)(off + j) * d4;
  for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = __ldg(src + c);
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("%s launch: %s", what, cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // namespace e2b

extern "C" int e2b_stage_clip(const float* emb_dev, const long long* meta_dev, const double* duration_dev, int B, int l, int d,
                              int sampling_rate, int frame_size, float* out_dev, cudaStream_t stream) {
  if (B <= 0 || l <= 0) return 0;
  if (d <= 0 || d % 4) { e2b_set_kernel_error("stage_clip: embedding width %d must be a positive multiple of 4", d); return -1; }
  if (sampling_rate <= 0 || frame_size <= 0) { e2b_set_kernel_error("stage_clip: sampling_rate and frame_size must be positive"); return -1; }
  if (B > 65535) { e2b_set_kernel_error("stage_clip: at most 65535 clips per call"); return -1; }
  e2b::ProfScope ps(stream, "stage_clip", (long long)B * l, d, 0, 0.0, 8.0 * B * l * (double)d);
  e2b::stage_clip_kernel<<<dim3(l, B), 256, 0, stream>>>(emb_dev, meta_dev, duration_dev, l, d / 4, sampling_rate, frame_size, out_dev);
  return e2b::check_launch("stage_clip");
}
