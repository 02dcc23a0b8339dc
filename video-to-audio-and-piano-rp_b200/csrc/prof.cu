#include "prof.h"

#include <cstdio>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace e2b {
namespace {
struct Rec {
  cudaEvent_t a, b;
  std::string kind;
  long long m, n, k;
  double flops, bytes;
};
bool g_on = false;
std::vector<Rec> g_recs;
}  // namespace

ProfScope::ProfScope(cudaStream_t s, const char* kind, long long m, long long n, long long k, double flops, double bytes) : idx(-1), st(s) {
  nvtxRangePushA(kind);
  if (!g_on) return;
  Rec r;
  r.kind = kind; r.m = m; r.n = n; r.k = k; r.flops = flops; r.bytes = bytes;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  idx = (int)g_recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_recs[idx].b, st);
  nvtxRangePop();
}
}  // namespace e2b

using namespace e2b;

extern "C" bool e2b_prof_is_on_() { return e2b::g_on; }

extern "C" void e2b_prof_enable(int on) {
  for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_recs.clear();
  g_on = on != 0;
}

extern "C" int e2b_prof_report(char* buf, int buflen) {
  cudaDeviceSynchronize();
  struct Agg { long long count = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::tuple<std::string, long long, long long, long long>, Agg> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    Agg& a = agg[std::make_tuple(r.kind, r.m, r.n, r.k)];
    a.count++; a.ms += ms; a.flops = r.flops; a.bytes = r.bytes;
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %lld %lld %lld %.6f %.6e %.6e\n", std::get<0>(kv.first).c_str(), std::get<1>(kv.first),
             std::get<2>(kv.first), std::get<3>(kv.first), kv.second.count, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += line;
  }
  if (buf && buflen > 0) snprintf(buf, buflen, "%s", out.c_str());
  return (int)out.size() + 1;
}
