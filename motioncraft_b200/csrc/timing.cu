#include "timing.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace mcm {
namespace {
struct Rec { int kind; cudaEvent_t a, b; double flops; };
std::atomic<bool> g_on{false};
std::mutex g_mu;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

LaunchTimer::LaunchTimer(int kind, cudaStream_t st, double flops) : idx_(-1), st_(st) {
  if (!g_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  Rec r{kind, get_event(), get_event(), flops};
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  idx_ = (int)g_recs.size() - 1;
}
LaunchTimer::~LaunchTimer() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[idx_].b, st_);
}

void timing_enable(bool on) { g_on.store(on); }
bool timing_enabled() { return g_on.load(std::memory_order_relaxed); }

int timing_collect(double* ms, unsigned long long* launches, double* flops) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_mu);
  for (int k = 0; k < LK_COUNT; ++k) { ms[k] = 0.0; launches[k] = 0; flops[k] = 0.0; }
  for (auto& r : g_recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.kind] += t;
    launches[r.kind] += 1;
    flops[r.kind] += r.flops;
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}
}  // namespace mcm
