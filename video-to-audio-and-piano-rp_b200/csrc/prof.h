// Optional per-launch CUDA-event profiler (off by default).  bench.py switches it on for one profiled pass to get the
// live per-kernel durations its roofline block reports; it is never on inside a timed region.
//
// Tracing: every launch sits inside an NVTX range named after its kernel class (ProfScope: gemm_qkv,
// gemm_resid, gemm_geglu, attention, dwconv, rmsnorm, guided_euler, ...) and engine.cu wraps the phases of the forward in
// NvtxRange scopes (e2b.text_stream, e2b.audio.self_attn, ...), so an Nsight Systems / ncu --nvtx timeline groups the ~400
// launches of an Euler update by layer and by class.  NVTX v3 is header-only: without a tool attached a range is a null
// function-pointer test.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

namespace e2b {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct ProfScope {
  ProfScope(cudaStream_t st, const char* kind, long long m, long long n, long long k, double flops, double bytes);
  ~ProfScope();
  int idx;
  cudaStream_t st;
};
}  // namespace e2b

extern "C" {
void e2b_prof_enable(int on);                       // clears previous records
bool e2b_prof_is_on_();                             // (events cannot be recorded inside a graph capture)
// Writes lines "kind m n k count total_ms flops_per_launch bytes_per_launch\n" into buf; returns bytes needed.
int e2b_prof_report(char* buf, int buflen);
}
