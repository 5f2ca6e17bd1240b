// motioncraft_b200 -- optional per-launch CUDA-event timing (bench.py's roofline leg).
// When enabled every kernel launch of this library is bracketed by a pair of events recorded on the
// launching stream; timing_collect() synchronises and sums the device time per kernel class.
// Disabled (the default) it costs one relaxed atomic load per launch.
#pragma once
#include <cuda_runtime.h>

namespace mcm {
enum LaunchKind : int { LK_GEMM = 0, LK_ROW = 1, LK_FUSED = 2, LK_COUNT = 3 };

class LaunchTimer {
 public:
  LaunchTimer(int kind, cudaStream_t st, double flops = 0.0);
  ~LaunchTimer();
 private:
  int idx_;
  cudaStream_t st_;
};

void timing_enable(bool on);
bool timing_enabled();
// returns 0; fills ms / launches / flops per LaunchKind and clears the record
int timing_collect(double* ms, unsigned long long* launches, double* flops);
}  // namespace mcm
