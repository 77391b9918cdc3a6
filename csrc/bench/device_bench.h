// C++ device benchmark loop (the "stimulator benchmark" of the reference, re-done over device
// timers).  Reference: scaelum/dynamics/benchmarker.py:49-71 + estimator.py:15-34 time a Conv2d
// proxy with wall clock + cuda.synchronize and no warm-up; here every rank times the REAL
// transformer-block GEMM chain (QKV, attn-out, FFN1+GELU, FFN2 on the tcgen05 kernel) with CUDA
// events after warm-up, so the measured speed predicts the real per-layer cost.
#pragma once
#include <utility>

namespace sky {
// returns (total seconds for `iterations` proxy blocks incl. simulated slowdown, free HBM in MiB)
std::pair<double, double> device_benchmark(int tokens, int hidden, int intermediate, int iterations,
                                           int warmup, double slowdown);
}  // namespace sky
