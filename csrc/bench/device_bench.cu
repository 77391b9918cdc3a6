#include "device_bench.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "../kernels/api.h"

namespace sky {

namespace {
void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
void ckrc(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + " rc=" + std::to_string(rc));
}
}  // namespace

std::pair<double, double> device_benchmark(int tokens, int hidden, int intermediate, int iterations,
                                           int warmup, double slowdown) {
  const size_t M = tokens, H = hidden, I = intermediate;
  __nv_bfloat16 *x, *wqkv, *qkv, *wo, *h1, *w1, *inter, *w2, *out;
  float* bias;
  uint64_t* tslot;
  ck(cudaMalloc(&x, M * H * 2), "malloc");
  ck(cudaMalloc(&wqkv, 3 * H * H * 2), "malloc");
  ck(cudaMalloc(&qkv, M * 3 * H * 2), "malloc");
  ck(cudaMalloc(&wo, H * H * 2), "malloc");
  ck(cudaMalloc(&h1, M * H * 2), "malloc");
  ck(cudaMalloc(&w1, I * H * 2), "malloc");
  ck(cudaMalloc(&inter, M * I * 2), "malloc");
  ck(cudaMalloc(&w2, H * I * 2), "malloc");
  ck(cudaMalloc(&out, M * H * 2), "malloc");
  ck(cudaMalloc(&bias, (3 * H > I ? 3 * H : I) * sizeof(float)), "malloc");
  ck(cudaMalloc(&tslot, sizeof(uint64_t)), "malloc");
  // small non-zero operands (value pattern irrelevant for timing, avoid denormals/NaN)
  ck(cudaMemset(x, 0x3c, M * H * 2), "memset");
  ck(cudaMemset(wqkv, 0x3c, 3 * H * H * 2), "memset");
  ck(cudaMemset(wo, 0x3c, H * H * 2), "memset");
  ck(cudaMemset(w1, 0x3c, I * H * 2), "memset");
  ck(cudaMemset(w2, 0x3c, H * I * 2), "memset");
  ck(cudaMemset(bias, 0, (3 * H > I ? 3 * H : I) * sizeof(float)), "memset");

  cudaStream_t s;
  ck(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "stream");
  cudaEvent_t e0, e1;
  ck(cudaEventCreate(&e0), "event");
  ck(cudaEventCreate(&e1), "event");

  auto block = [&]() {
    GemmArgs g;
    g.M = tokens;
    g.bias = bias;
    // QKV
    g.A = x; g.lda = hidden; g.B = wqkv; g.ldb = hidden; g.N = 3 * hidden; g.K = hidden;
    g.out = qkv; g.ldo = 3 * hidden; g.act = ACT_NONE;
    ckrc(launch_gemm(g, s), "gemm qkv");
    // attention output projection (A = first H columns of qkv as a stand-in for ctx)
    g.A = qkv; g.lda = 3 * hidden; g.B = wo; g.ldb = hidden; g.N = hidden; g.K = hidden;
    g.out = h1; g.ldo = hidden;
    ckrc(launch_gemm(g, s), "gemm attn-out");
    // FFN1 + GELU
    g.A = h1; g.lda = hidden; g.B = w1; g.ldb = hidden; g.N = intermediate; g.K = hidden;
    g.out = inter; g.ldo = intermediate; g.act = ACT_GELU;
    ckrc(launch_gemm(g, s), "gemm ffn1");
    // FFN2
    g.A = inter; g.lda = intermediate; g.B = w2; g.ldb = intermediate; g.N = hidden;
    g.K = intermediate; g.out = out; g.ldo = hidden; g.act = ACT_NONE;
    ckrc(launch_gemm(g, s), "gemm ffn2");
  };

  for (int i = 0; i < warmup; ++i) block();
  ck(cudaStreamSynchronize(s), "sync");
  ck(cudaEventRecord(e0, s), "record");
  for (int i = 0; i < iterations; ++i) {
    if (slowdown > 0) ckrc(launch_record_time(tslot, s), "record_time");
    block();
    // simulated slow device: spin on the GPU for slowdown x the block's own duration
    if (slowdown > 0) ckrc(launch_spin_factor(tslot, static_cast<float>(slowdown), s), "spin");
  }
  ck(cudaEventRecord(e1, s), "record");
  ck(cudaEventSynchronize(e1), "sync");
  float ms = 0.f;
  ck(cudaEventElapsedTime(&ms, e0, e1), "elapsed");

  cudaFree(x); cudaFree(wqkv); cudaFree(qkv); cudaFree(wo); cudaFree(h1); cudaFree(w1);
  cudaFree(inter); cudaFree(w2); cudaFree(out); cudaFree(bias); cudaFree(tslot);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
  size_t free_b = 0, total_b = 0;
  ck(cudaMemGetInfo(&free_b, &total_b), "meminfo");
  return {static_cast<double>(ms) * 1e-3, static_cast<double>(free_b) / (1024.0 * 1024.0)};
}

}  // namespace sky
