// Exact fp32 attention for the error-compensated ("fp32") mode: same semantics as attention.cu
//     sim = 50 tanh(q k^T / (8 * 50)) ; key-length mask ; softmax ; out = attn v ; out *= sigmoid gate
// but every operand and every operation in fp32 with accurate tanhf / expf.  SIMT, one warp per (query row, head, batch):
// lane j walks keys j, j+32, ... with the query row and the 64-wide output accumulator in registers, then the 32 partial
// accumulators are reduced with shuffles.  Validation mode (rel-L2 <= 1e-4 against the fp32 reference), not the throughput path.
#include "kernels.h"
#include "prof.h"
#include "ptx.cuh"

namespace e2b {

constexpr int AF_WARPS = 8;

__global__ void __launch_bounds__(32 * AF_WARPS) attention_f32_kernel(const e2b_attn_f32_desc d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qpos = blockIdx.x * AF_WARPS + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (qpos >= d.q_rows_per_batch) return;
  const int kvb = d.kv_batch_mod > 0 ? b % d.kv_batch_mod : b;
  int kv_len = d.kv_lens ? (__ldg(d.kv_lens + kvb) + d.kv_lens_add) : d.kv_rows_per_batch;
  kv_len = min(kv_len, d.kv_rows_per_batch);
  const size_t qrow = (size_t)b * d.q_rows_per_batch + qpos;
  const float4* qp = reinterpret_cast<const float4*>(d.q + qrow * d.ldq + d.q_col0 + h * 64);
  float4 q[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) q[i] = __ldg(qp + i);
  float4 o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float l = 0.f;
  const float clamp = d.softclamp;
  for (int j = lane; j < kv_len; j += 32) {
    const size_t krow = (size_t)kvb * d.kv_rows_per_batch + j;
    const float4* kp = reinterpret_cast<const float4*>(d.k + krow * d.ldk + d.k_col0 + h * 64);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 kk = __ldg(kp + i);
      s = fmaf(q[i].x, kk.x, s); s = fmaf(q[i].y, kk.y, s); s = fmaf(q[i].z, kk.z, s); s = fmaf(q[i].w, kk.w, s);
    }
    const float p = expf(tanhf(s / clamp) * clamp);       // |logit| <= clamp: no running maximum needed in fp32
    l += p;
    const float4* vp = reinterpret_cast<const float4*>(d.v + krow * d.ldv + d.v_col0 + h * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 vv = __ldg(vp + i);
      o[i].x = fmaf(p, vv.x, o[i].x); o[i].y = fmaf(p, vv.y, o[i].y); o[i].z = fmaf(p, vv.z, o[i].z); o[i].w = fmaf(p, vv.w, o[i].w);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, off);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i].x += __shfl_xor_sync(0xffffffffu, o[i].x, off);
      o[i].y += __shfl_xor_sync(0xffffffffu, o[i].y, off);
      o[i].z += __shfl_xor_sync(0xffffffffu, o[i].z, off);
      o[i].w += __shfl_xor_sync(0xffffffffu, o[i].w, off);
    }
  }
  float scale = l > 0.f ? 1.0f / l : 0.f;
  if (d.hgate) scale *= __ldg(d.hgate + qrow * d.hgate_ld + h);
  // lanes 0..15 each store one float4 (4 channels) of the 64-wide head output
  if (lane < 16) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i == lane) r = o[i];
    r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(d.out) + qrow * d.ldo + h * 64 + lane * 4;
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(r.x, r.y), h1 = __floats2bfloat162_rn(r.z, r.w);
    uint2 hi;
    hi.x = *reinterpret_cast<const uint32_t*>(&h0);
    hi.y = *reinterpret_cast<const uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(op) = hi;
    if (d.out_split > 0) {
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      *reinterpret_cast<uint2*>(op + d.out_split) = make_uint2(pack_bf16(r.x - f0.x, r.y - f0.y), pack_bf16(r.z - f1.x, r.w - f1.y));
    }
  }
}

}  // namespace e2b

extern "C" int e2b_attention_f32_launch(const e2b_attn_f32_desc* d, cudaStream_t stream) {
  if (d->batch <= 0 || d->heads <= 0 || d->q_rows_per_batch <= 0) return 0;
  if ((d->ldq | d->ldk | d->ldv | d->q_col0 | d->k_col0 | d->v_col0 | d->ldo | d->out_split) % 4) {
    e2b_set_kernel_error("attention_f32: leading dimensions / column offsets must be multiples of 4");
    return -1;
  }
  dim3 grid((d->q_rows_per_batch + e2b::AF_WARPS - 1) / e2b::AF_WARPS, d->heads, d->batch);
  e2b::ProfScope ps(stream, "attention_f32", (long long)d->batch * d->q_rows_per_batch, d->kv_rows_per_batch, d->heads,
                    4.0 * d->batch * d->heads * (double)d->q_rows_per_batch * d->kv_rows_per_batch * 64.0, 0.0);
  e2b::attention_f32_kernel<<<grid, 32 * e2b::AF_WARPS, 0, stream>>>(*d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { e2b_set_kernel_error("attention_f32 launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
