// Optional per-launch CUDA-event profiler (off by default).  bench.py switches it on for one profiled pass to get the
// live per-kernel durations its roofline block reports; it is never on inside a timed region.
#pragma once
#include <cuda_runtime.h>

namespace e2b {
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
