// Memory-bound sm_100a kernels: LayerNorm fwd/bwd (optionally gated on cross-GPU panel flags),
// bias-gradient column sums, fused embedding gather+LN+dropout and its scatter backward, the
// latency-bound pooler / classifier / cross-entropy heads, fused multi-tensor SGD, and the small
// runtime kernels (RNG step, flag epochs, device-side throttle, peer copy + signal).
//
// Reference parity (what these replace):
//   BertLayerNorm            scaelum/model/bert_layers.py:143-168   -> layernorm_{fwd,bwd}
//   BertEmbeddings.forward   scaelum/model/bert_layers.py:191-212   -> embed_{fwd,bwd}
//   BertPooler / classifier  scaelum/model/bert_layers.py:366-395   -> small_linear_{fwd,bwd}
//   nn.CrossEntropyLoss      scaelum/runner/runner.py:51-52,131     -> softmax_ce
//   optim.SGD step           experiment/launch.py:152-155           -> sgd_multi
//   time.sleep slowdown      scaelum/builder/module_wrapper.py:124-126,265-266 -> spin_* kernels
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "api.h"
#include "launch_util.h"
#include "sm100_ptx.cuh"

namespace sky {

namespace {

constexpr uint64_t kFlagTimeoutNs = 4000000000ull;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void warp_wait_panel(const uint32_t* flags, int panel, uint32_t target,
                                                int* error_flag, int lane) {
  if (lane == 0) {
    if (!wait_flag_ge(flags + panel, target, kFlagTimeoutNs)) {
      if (error_flag) atomicExch(error_flag, 2);
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, 8 bf16 (16B) per lane per step, row held in registers.
// H % 8 == 0, H <= 2048.  NV = number of 16B vectors per lane (compile-time).
// ------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 8;  // H <= 8*256
constexpr int kLnRowsPerCta = 8;  // one row per warp; a 128-row panel = 16 signals

template <int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ y,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int M, int H,
                     float eps, const uint32_t* wait_flags, const uint32_t* wait_epoch,
                     uint32_t wait_mult, int* error_flag, uint32_t* signal_flags) {
  pdl_wait();
  pdl_launch_dependents();
  // Each CTA owns kLnRowsPerCta consecutive rows (8 warps x 4 rows) so that a finished CTA can
  // publish "32 rows of panel p are written" with one release.sys add (y may be peer memory).
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = H / 8;  // 16B vectors per row
  const float inv_h = 1.f / static_cast<float>(H);
  const uint32_t target = wait_flags ? (*wait_epoch) * wait_mult : 0u;
  int last_panel = -1;
  const int row_base = blockIdx.x * kLnRowsPerCta;

  for (int r = warp; r < kLnRowsPerCta; r += 8) {
    const int row = row_base + r;
    if (row >= M) break;
    if (wait_flags != nullptr) {
      const int panel = row >> 7;
      if (panel != last_panel) {
        warp_wait_panel(wait_flags, panel, target, error_flag, lane);
        last_panel = panel;
      }
    }
    const uint4* zr = reinterpret_cast<const uint4*>(z + static_cast<long long>(row) * H);
    float x[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint4 u = zr[v];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_bf16x2(w[t]);
          x[i][2 * t] = f.x;
          x[i][2 * t + 1] = f.y;
          s += f.x + f.y;
        }
      }
    }
    const float mu = warp_sum(s) * inv_h;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float d = x[i][t] - mu;
          sq += d * d;
        }
      }
    }
    const float var = warp_sum(sq) * inv_h;
    const float rs = rsqrtf(var + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rs;
    }
    uint4* yr = reinterpret_cast<uint4*>(y + static_cast<long long>(row) * H);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        uint4 o;
        o.x = pack_bf16x2((x[i][0] - mu) * rs * g0.x + b0.x, (x[i][1] - mu) * rs * g0.y + b0.y);
        o.y = pack_bf16x2((x[i][2] - mu) * rs * g0.z + b0.z, (x[i][3] - mu) * rs * g0.w + b0.w);
        o.z = pack_bf16x2((x[i][4] - mu) * rs * g1.x + b1.x, (x[i][5] - mu) * rs * g1.y + b1.y);
        o.w = pack_bf16x2((x[i][6] - mu) * rs * g1.z + b1.z, (x[i][7] - mu) * rs * g1.w + b1.w);
        yr[v] = o;
      }
    }
  }
  if (signal_flags != nullptr) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) red_release_sys_add(signal_flags + (row_base >> 7), 1u);
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward (input gradient).  One warp per row; parameter gradients are produced by
// ln_param_grad_kernel below (column-oriented, no per-lane column partials -> low registers).
// ------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dz,
                     __nv_bfloat16* __restrict__ dz_dropped, int M, int H, float dropout_p,
                     const uint64_t* rng_state, uint32_t rng_stream, const uint32_t* wait_flags,
                     const uint32_t* wait_epoch, uint32_t wait_mult, int* error_flag) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  const int nvec = H / 8;
  const float inv_h = 1.f / static_cast<float>(H);
  const uint32_t target = wait_flags ? (*wait_epoch) * wait_mult : 0u;
  int last_panel = -1;
  const bool has_dropout = dropout_p > 0.f && dz_dropped != nullptr;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float drop_scale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(rng_state, rng_stream);
    thr16 = static_cast<uint32_t>(dropout_p * 65536.f);
    drop_scale = 1.f / (1.f - dropout_p);
  }

  for (int row = warp_global; row < M; row += total_warps) {
    if (wait_flags != nullptr) {
      const int panel = row >> 7;
      if (panel != last_panel) {
        warp_wait_panel(wait_flags, panel, target, error_flag, lane);
        last_panel = panel;
      }
    }
    const float mu = mean[row];
    const float rs = rstd[row];
    const uint4* zr = reinterpret_cast<const uint4*>(z + static_cast<long long>(row) * H);
    const uint4* dr = reinterpret_cast<const uint4*>(dy + static_cast<long long>(row) * H);
    float xh[NV][8], g[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const uint4 uz = zr[v];
        const uint4 ud = dr[v];
        const uint32_t wz[4] = {uz.x, uz.y, uz.z, uz.w};
        const uint32_t wd[4] = {ud.x, ud.y, ud.z, ud.w};
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 fz = unpack_bf16x2(wz[t]);
          const float2 fd = unpack_bf16x2(wd[t]);
          const float x0 = (fz.x - mu) * rs, x1 = (fz.y - mu) * rs;
          xh[i][2 * t] = x0;
          xh[i][2 * t + 1] = x1;
          const float q0 = fd.x * gm[2 * t], q1 = fd.y * gm[2 * t + 1];
          g[i][2 * t] = q0;
          g[i][2 * t + 1] = q1;
          s1 += q0 + q1;
          s2 += q0 * x0 + q1 * x1;
        }
      }
    }
    const float c1 = warp_sum(s1) * inv_h;
    const float c2 = warp_sum(s2) * inv_h;
    uint4* dzr = reinterpret_cast<uint4*>(dz + static_cast<long long>(row) * H);
    uint4* dzd = has_dropout
                     ? reinterpret_cast<uint4*>(dz_dropped + static_cast<long long>(row) * H)
                     : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = (g[i][t] - c1 - xh[i][t] * c2) * rs;
        dzr[v] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                            pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        if (has_dropout) {
          const uint64_t base = (static_cast<uint64_t>(row) * H + v * 8) >> 2;
          const uint32_t m0 = dropout_keep4(seed, base, thr16);
          const uint32_t m1 = dropout_keep4(seed, base + 1, thr16);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            o[t] = (m0 >> t) & 1u ? o[t] * drop_scale : 0.f;
            o[4 + t] = (m1 >> t) & 1u ? o[4 + t] * drop_scale : 0.f;
          }
          dzd[v] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                              pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
      }
    }
  }
}

// dgamma[c] += sum_r dy[r,c] * xhat[r,c];  dbeta[c] += sum_r dy[r,c]
// block = 32 column groups (8 columns each) x 8 row lanes; grid.y splits the rows.
__global__ void __launch_bounds__(256)
ln_param_grad_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int H,
                     const __nv_bfloat16* __restrict__ x2, float* __restrict__ out2) {
  // x2 / out2 (optional): out2[c] += sum_r x2[r,c] - the bias gradient of the dense layer in front
  // of this LayerNorm (column sum of the LayerNorm input gradient), fused here to save a launch
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float redg[8][256 + 8];
  __shared__ float redb[8][256 + 8];
  __shared__ float red2[8][256 + 8];
  float a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int cg = threadIdx.x & 31;
  const int rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < H) {
    for (int r = blockIdx.y * 8 + rl; r < M; r += gridDim.y * 8) {
      const float mu = mean[r], rs = rstd[r];
      const uint4 ud = *reinterpret_cast<const uint4*>(dy + static_cast<long long>(r) * H + col);
      const uint4 uz = *reinterpret_cast<const uint4*>(z + static_cast<long long>(r) * H + col);
      const uint32_t wd[4] = {ud.x, ud.y, ud.z, ud.w};
      const uint32_t wz[4] = {uz.x, uz.y, uz.z, uz.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 fd = unpack_bf16x2(wd[t]);
        const float2 fz = unpack_bf16x2(wz[t]);
        ag[2 * t] += fd.x * (fz.x - mu) * rs;
        ag[2 * t + 1] += fd.y * (fz.y - mu) * rs;
        ab[2 * t] += fd.x;
        ab[2 * t + 1] += fd.y;
      }
      if (x2 != nullptr) {
        const uint4 ux = *reinterpret_cast<const uint4*>(x2 + static_cast<long long>(r) * H + col);
        const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 fx = unpack_bf16x2(wx[t]);
          a2[2 * t] += fx.x;
          a2[2 * t + 1] += fx.y;
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    redg[rl][cg * 8 + t] = ag[t];
    redb[rl][cg * 8 + t] = ab[t];
    red2[rl][cg * 8 + t] = a2[t];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float sg = 0.f, sb = 0.f, s2 = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    sg += redg[r][c];
    sb += redb[r][c];
    s2 += red2[r][c];
  }
  const int gc = blockIdx.x * 256 + c;
  if (gc < H) {
    if (dgamma) atomicAdd(&dgamma[gc], sg);
    if (dbeta) atomicAdd(&dbeta[gc], sb);
    if (out2) atomicAdd(&out2[gc], s2);
  }
}

// ------------------------------------------------------------------------------------------
// Column sums of a bf16 [M,N] matrix into fp32 (+=).  block = (32 col-groups of 8) x 8 row lanes
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, int M, int N, long long ldx,
              float* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31;
  const int rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (int r = blockIdx.y * 8 + rl; r < M; r += gridDim.y * 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + r * ldx + col);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = unpack_bf16x2(w[t]);
        acc[2 * t] += f.x;
        acc[2 * t + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) red[rl][cg * 8 + t] = acc[t];
  __syncthreads();
  const int c = threadIdx.x;  // 256 columns per block
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) s += red[r][c];
  const int gc = blockIdx.x * 256 + c;
  if (gc < N) atomicAdd(&out[gc], s);
}

// ------------------------------------------------------------------------------------------
// Embeddings
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tts,
                 const int64_t* __restrict__ amask, const float* __restrict__ word,
                 const float* __restrict__ pos, const float* __restrict__ type,
                 const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ ext_mask,
                 float* __restrict__ mean_out, float* __restrict__ rstd_out, int BS, int S, int H,
                 float eps, float dropout_p, const uint64_t* rng_state, uint32_t rng_stream) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int row0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  const float inv_h = 1.f / static_cast<float>(H);
  const bool has_dropout = dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float drop_scale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(rng_state, rng_stream);
    thr16 = static_cast<uint32_t>(dropout_p * 65536.f);
    drop_scale = 1.f / (1.f - dropout_p);
  }
  for (int row = row0; row < BS; row += total_warps) {
    const long long id = ids[row];
    const long long tt = tts[row];
    const int s = row % S;
    if (lane == 0 && ext_mask != nullptr)
      ext_mask[row] = amask ? (1.0f - static_cast<float>(amask[row])) * -10000.0f : 0.f;
    const float* wr = word + id * H;
    const float* pr = pos + static_cast<long long>(s) * H;
    const float* tr = type + tt * H;
    float sum = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(wr + c);
      const float4 b = *reinterpret_cast<const float4*>(pr + c);
      const float4 d = *reinterpret_cast<const float4*>(tr + c);
      sum += (a.x + b.x + d.x) + (a.y + b.y + d.y) + (a.z + b.z + d.z) + (a.w + b.w + d.w);
    }
    const float mu = warp_sum(sum) * inv_h;
    float sq = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(wr + c);
      const float4 b = *reinterpret_cast<const float4*>(pr + c);
      const float4 d = *reinterpret_cast<const float4*>(tr + c);
      const float e0 = a.x + b.x + d.x - mu, e1 = a.y + b.y + d.y - mu;
      const float e2 = a.z + b.z + d.z - mu, e3 = a.w + b.w + d.w - mu;
      sq += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    }
    const float rs = rsqrtf(warp_sum(sq) * inv_h + eps);
    if (lane == 0) {
      mean_out[row] = mu;
      rstd_out[row] = rs;
    }
    __nv_bfloat16* orow = out + static_cast<long long>(row) * H;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(wr + c);
      const float4 b = *reinterpret_cast<const float4*>(pr + c);
      const float4 d = *reinterpret_cast<const float4*>(tr + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
      float o0 = (a.x + b.x + d.x - mu) * rs * g.x + be.x;
      float o1 = (a.y + b.y + d.y - mu) * rs * g.y + be.y;
      float o2 = (a.z + b.z + d.z - mu) * rs * g.z + be.z;
      float o3 = (a.w + b.w + d.w - mu) * rs * g.w + be.w;
      if (has_dropout) {
        const uint32_t m = dropout_keep4(seed, (static_cast<uint64_t>(row) * H + c) >> 2, thr16);
        o0 = (m & 1u) ? o0 * drop_scale : 0.f;
        o1 = (m & 2u) ? o1 * drop_scale : 0.f;
        o2 = (m & 4u) ? o2 * drop_scale : 0.f;
        o3 = (m & 8u) ? o3 * drop_scale : 0.f;
      }
      uint2 pk;
      pk.x = pack_bf16x2(o0, o1);
      pk.y = pack_bf16x2(o2, o3);
      *reinterpret_cast<uint2*>(orow + c) = pk;
    }
  }
}

// backward: recompute xhat from the tables, LN backward, scatter-add into the fp32 grad tables.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const int64_t* __restrict__ ids,
                 const int64_t* __restrict__ tts, const float* __restrict__ word,
                 const float* __restrict__ pos, const float* __restrict__ type,
                 const float* __restrict__ gamma, const float* __restrict__ mean,
                 const float* __restrict__ rstd, float* __restrict__ dword,
                 float* __restrict__ dpos, float* __restrict__ dtype_, float* __restrict__ dgamma,
                 float* __restrict__ dbeta, int BS, int S, int H, float dropout_p,
                 const uint64_t* rng_state, uint32_t rng_stream, const uint32_t* wait_flags,
                 const uint32_t* wait_epoch, uint32_t wait_mult, int* error_flag,
                 int type_rows_in_smem) {
  // Heavily shared targets (LayerNorm gamma/beta: every row; token-type rows: ~half of all
  // rows each) are accumulated in shared memory and flushed once per CTA; the first version
  // issued one global atomic per element and spent 400 us serialising on ~4K addresses.
  extern __shared__ float acc_smem[];  // [dgamma H][dbeta H][dtype type_rows_in_smem * H]
  float* s_dgamma = acc_smem;
  float* s_dbeta = acc_smem + H;
  float* s_dtype = acc_smem + 2 * H;
  for (int i = threadIdx.x; i < (2 + type_rows_in_smem) * H; i += blockDim.x) acc_smem[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int row0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int total_warps = gridDim.x * warps_per_block;
  const float inv_h = 1.f / static_cast<float>(H);
  const bool has_dropout = dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float drop_scale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(rng_state, rng_stream);
    thr16 = static_cast<uint32_t>(dropout_p * 65536.f);
    drop_scale = 1.f / (1.f - dropout_p);
  }
  const uint32_t target = wait_flags ? (*wait_epoch) * wait_mult : 0u;
  int last_panel = -1;
  for (int row = row0; row < BS; row += total_warps) {
    if (wait_flags != nullptr) {
      const int panel = row >> 7;
      if (panel != last_panel) {
        warp_wait_panel(wait_flags, panel, target, error_flag, lane);
        last_panel = panel;
      }
    }
    const long long id = ids[row];
    const long long tt = tts[row];
    const int s = row % S;
    const float mu = mean[row], rs = rstd[row];
    const float* wr = word + id * H;
    const float* pr = pos + static_cast<long long>(s) * H;
    const float* tr = type + tt * H;
    const __nv_bfloat16* drow = dout + static_cast<long long>(row) * H;
    float s1 = 0.f, s2 = 0.f;
    // pass 1: row statistics of g = dy*gamma
    for (int c = lane * 4; c < H; c += 128) {
      const uint2 u = *reinterpret_cast<const uint2*>(drow + c);
      float2 d01 = unpack_bf16x2(u.x), d23 = unpack_bf16x2(u.y);
      float dy[4] = {d01.x, d01.y, d23.x, d23.y};
      if (has_dropout) {
        const uint32_t m = dropout_keep4(seed, (static_cast<uint64_t>(row) * H + c) >> 2, thr16);
#pragma unroll
        for (int t = 0; t < 4; ++t) dy[t] = (m >> t) & 1u ? dy[t] * drop_scale : 0.f;
      }
      const float4 a = *reinterpret_cast<const float4*>(wr + c);
      const float4 b = *reinterpret_cast<const float4*>(pr + c);
      const float4 d = *reinterpret_cast<const float4*>(tr + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float e[4] = {a.x + b.x + d.x, a.y + b.y + d.y, a.z + b.z + d.z, a.w + b.w + d.w};
      const float gm[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float xh = (e[t] - mu) * rs;
        const float q = dy[t] * gm[t];
        s1 += q;
        s2 += q * xh;
        atomicAdd(&s_dgamma[c + t], dy[t] * xh);
        atomicAdd(&s_dbeta[c + t], dy[t]);
      }
    }
    const float c1 = warp_sum(s1) * inv_h;
    const float c2 = warp_sum(s2) * inv_h;
    for (int c = lane * 4; c < H; c += 128) {
      const uint2 u = *reinterpret_cast<const uint2*>(drow + c);
      float2 d01 = unpack_bf16x2(u.x), d23 = unpack_bf16x2(u.y);
      float dy[4] = {d01.x, d01.y, d23.x, d23.y};
      if (has_dropout) {
        const uint32_t m = dropout_keep4(seed, (static_cast<uint64_t>(row) * H + c) >> 2, thr16);
#pragma unroll
        for (int t = 0; t < 4; ++t) dy[t] = (m >> t) & 1u ? dy[t] * drop_scale : 0.f;
      }
      const float4 a = *reinterpret_cast<const float4*>(wr + c);
      const float4 b = *reinterpret_cast<const float4*>(pr + c);
      const float4 d = *reinterpret_cast<const float4*>(tr + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float e[4] = {a.x + b.x + d.x, a.y + b.y + d.y, a.z + b.z + d.z, a.w + b.w + d.w};
      const float gm[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float xh = (e[t] - mu) * rs;
        const float de = (dy[t] * gm[t] - c1 - xh * c2) * rs;
        atomicAdd(&dword[id * H + c + t], de);
        atomicAdd(&dpos[static_cast<long long>(s) * H + c + t], de);
        if (tt < type_rows_in_smem)
          atomicAdd(&s_dtype[tt * H + c + t], de);
        else
          atomicAdd(&dtype_[tt * H + c + t], de);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_dgamma[i]);
    atomicAdd(&dbeta[i], s_dbeta[i]);
  }
  for (int i = threadIdx.x; i < type_rows_in_smem * H; i += blockDim.x)
    atomicAdd(&dtype_[i], s_dtype[i]);
}

// ------------------------------------------------------------------------------------------
// Small dense layers (pooler, classifier): one warp per output column n, looping over rows.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_x(const void* x, bool x_bf16, long long idx) {
  return x_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx])
                : reinterpret_cast<const float*>(x)[idx];
}
__device__ __forceinline__ float drop1(uint64_t seed, uint64_t idx, uint32_t thr16, float scale,
                                       float v) {
  const uint32_t m = dropout_keep4(seed, idx >> 2, thr16);
  return ((m >> (idx & 3)) & 1u) ? v * scale : 0.f;
}

__global__ void __launch_bounds__(128)
small_linear_fwd_kernel(const void* __restrict__ x, bool x_bf16, long long ldx,
                        const float* __restrict__ w, const float* __restrict__ b,
                        float* __restrict__ y, int M, int N, int K, int act_tanh, float dropout_p,
                        const uint64_t* rng_state, uint32_t rng_stream) {
  // one warp per output element (m, n): lanes stride the reduction dimension
  const int lane = threadIdx.x & 31;
  const long long o = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= static_cast<long long>(M) * N) return;
  const int m = static_cast<int>(o / N), n = static_cast<int>(o % N);
  const bool has_dropout = dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float scale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(rng_state, rng_stream);
    thr16 = static_cast<uint32_t>(dropout_p * 65536.f);
    scale = 1.f / (1.f - dropout_p);
  }
  const float* wr = w + static_cast<long long>(n) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float xv = load_x(x, x_bf16, m * ldx + k);
    if (has_dropout) xv = drop1(seed, static_cast<uint64_t>(m) * K + k, thr16, scale, xv);
    acc += xv * wr[k];
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (b ? b[n] : 0.f);
    if (act_tanh) v = tanhf(v);
    y[o] = v;
  }
}

// dx[m,k] = sum_n dy'[m,n] w[n,k]; one thread per k, 8 rows at a time.
__global__ void __launch_bounds__(128)
small_linear_dx_kernel(const float* __restrict__ w, const float* __restrict__ y,
                       const float* __restrict__ dy, void* __restrict__ dx, bool dx_bf16,
                       long long lddx, int M, int N, int K, int act_tanh, float dropout_p,
                       const uint64_t* rng_state, uint32_t rng_stream) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int m0 = blockIdx.y * 8;
  if (k >= K) return;
  const bool has_dropout = dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float scale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(rng_state, rng_stream);
    thr16 = static_cast<uint32_t>(dropout_p * 65536.f);
    scale = 1.f / (1.f - dropout_p);
  }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = 0; n < N; ++n) {
    const float wv = w[static_cast<long long>(n) * K + k];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + i;
      if (m < M) {
        float g = dy[static_cast<long long>(m) * N + n];
        if (act_tanh) {
          const float yv = y[static_cast<long long>(m) * N + n];
          g *= (1.f - yv * yv);
        }
        acc[i] += g * wv;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + i;
    if (m < M) {
      float v = acc[i];
      if (has_dropout) v = drop1(seed, static_cast<uint64_t>(m) * K + k, thr16, scale, v);
      if (dx_bf16)
        reinterpret_cast<__nv_bfloat16*>(dx)[m * lddx + k] = __float2bfloat16(v);
      else
        reinterpret_cast<float*>(dx)[m * lddx + k] = v;
    }
  }
}

// dw[n,k] += sum_m dy'[m,n] x[m,k];  db[n] += sum_m dy'[m,n]
__global__ void __launch_bounds__(128)
small_linear_dw_kernel(const void* __restrict__ x, bool x_bf16, long long ldx,
                       const float* __restrict__ y, const float* __restrict__ dy,
                       float* __restrict__ dw, float* __restrict__ db, int M, int N, int K,
                       int act_tanh, float dropout_p, const uint64_t* rng_state,
                       uint32_t rng_stream) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  const bool has_dropout = dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float scale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(rng_state, rng_stream);
    thr16 = static_cast<uint32_t>(dropout_p * 65536.f);
    scale = 1.f / (1.f - dropout_p);
  }
  float acc = 0.f, accb = 0.f;
  for (int m = 0; m < M; ++m) {
    float g = dy[static_cast<long long>(m) * N + n];
    if (act_tanh) {
      const float yv = y[static_cast<long long>(m) * N + n];
      g *= (1.f - yv * yv);
    }
    accb += g;
    if (k < K) {
      float xv = load_x(x, x_bf16, m * ldx + k);
      if (has_dropout) xv = drop1(seed, static_cast<uint64_t>(m) * K + k, thr16, scale, xv);
      acc += g * xv;
    }
  }
  if (k < K) dw[static_cast<long long>(n) * K + k] += acc;
  if (db != nullptr && k == 0) db[n] += accb;
}

// softmax cross entropy, mean reduction; single block.
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                  float* __restrict__ loss, float* __restrict__ dlogits, int M, int C,
                  float grad_scale, float* __restrict__ loss_acc) {
  __shared__ float red[256];
  float local = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const float* l = logits + static_cast<long long>(m) * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, l[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(l[c] - mx);
    const float lse = mx + logf(se);
    const int lab = static_cast<int>(labels[m]);
    local += lse - l[lab];
    if (dlogits != nullptr) {
      for (int c = 0; c < C; ++c) {
        const float pc = expf(l[c] - lse);
        dlogits[static_cast<long long>(m) * C + c] =
            (pc - (c == lab ? 1.f : 0.f)) * grad_scale / static_cast<float>(M);
      }
    }
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mean = red[0] / static_cast<float>(M);
    if (loss != nullptr) loss[0] = mean;
    // micro-batched training: the step loss is the sum of the scaled micro-batch losses
    if (loss_acc != nullptr) loss_acc[0] += mean * grad_scale;
  }
}

// ------------------------------------------------------------------------------------------
// Optimizer
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgd_multi_kernel(const SgdTensor* __restrict__ tensors, float lr, float momentum,
                 float weight_decay, float grad_scale, int zero_grad) {
  const SgdTensor t = tensors[blockIdx.y];
  const long long n4 = t.numel >> 2;
  float4* p4 = reinterpret_cast<float4*>(t.p);
  float4* g4 = reinterpret_cast<float4*>(t.g);
  float4* m4 = reinterpret_cast<float4*>(t.mom);
  uint2* b4 = reinterpret_cast<uint2*>(t.p_bf16);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 p = p4[i];
    float4 g = g4[i];
    g.x = g.x * grad_scale + weight_decay * p.x;
    g.y = g.y * grad_scale + weight_decay * p.y;
    g.z = g.z * grad_scale + weight_decay * p.z;
    g.w = g.w * grad_scale + weight_decay * p.w;
    if (m4 != nullptr) {
      float4 m = m4[i];
      m.x = momentum * m.x + g.x;
      m.y = momentum * m.y + g.y;
      m.z = momentum * m.z + g.z;
      m.w = momentum * m.w + g.w;
      m4[i] = m;
      g = m;
    }
    p.x -= lr * g.x;
    p.y -= lr * g.y;
    p.z -= lr * g.z;
    p.w -= lr * g.w;
    p4[i] = p;
    if (b4 != nullptr) b4[i] = make_uint2(pack_bf16x2(p.x, p.y), pack_bf16x2(p.z, p.w));
    if (zero_grad && !t.skip_zero) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // scalar tail
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < t.numel; i += blockDim.x) {
      float p = t.p[i];
      float g = t.g[i] * grad_scale + weight_decay * p;
      if (t.mom != nullptr) {
        const float m = momentum * t.mom[i] + g;
        t.mom[i] = m;
        g = m;
      }
      p -= lr * g;
      t.p[i] = p;
      if (t.p_bf16 != nullptr) reinterpret_cast<__nv_bfloat16*>(t.p_bf16)[i] = __float2bfloat16(p);
      if (zero_grad) t.g[i] = 0.f;
    }
  }
}

// Adam / AdamW over the same descriptors (torch.optim.Adam semantics, experiment/launch.py:152-155
// accepts any torch.optim name): one launch per stage, moments + master + bf16 shadow + gradient
// zeroing in one pass = 28 B / parameter.  The bias corrections come from a device-side step
// counter, so a captured graph replays the right correction every step.
__global__ void __launch_bounds__(256)
adam_multi_kernel(const SgdTensor* __restrict__ tensors, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int decoupled, const uint64_t* __restrict__ step,
                  float grad_scale, int zero_grad) {
  const SgdTensor t = tensors[blockIdx.y];
  const float tstep = static_cast<float>(*step);
  const float bc1 = 1.f - powf(beta1, tstep);
  const float bc2_rsqrt = rsqrtf(1.f - powf(beta2, tstep));
  const float step_size = lr / bc1;
  const float decay = decoupled ? 1.f - lr * weight_decay : 1.f;
  const float l2 = decoupled ? 0.f : weight_decay;
  auto upd = [&](float& p, float gi, float& m, float& v) {
    p *= decay;
    const float g = gi * grad_scale + l2 * p;
    m = beta1 * m + (1.f - beta1) * g;
    v = beta2 * v + (1.f - beta2) * g * g;
    p -= step_size * m / (sqrtf(v) * bc2_rsqrt + eps);
  };
  const long long n4 = t.numel >> 2;
  float4* p4 = reinterpret_cast<float4*>(t.p);
  float4* g4 = reinterpret_cast<float4*>(t.g);
  float4* m4 = reinterpret_cast<float4*>(t.mom);
  float4* v4 = reinterpret_cast<float4*>(t.mom2);
  uint2* b4 = reinterpret_cast<uint2*>(t.p_bf16);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    const float4 g = g4[i];
    upd(p.x, g.x, m.x, v.x);
    upd(p.y, g.y, m.y, v.y);
    upd(p.z, g.z, m.z, v.z);
    upd(p.w, g.w, m.w, v.w);
    p4[i] = p;
    m4[i] = m;
    v4[i] = v;
    if (b4 != nullptr) b4[i] = make_uint2(pack_bf16x2(p.x, p.y), pack_bf16x2(p.z, p.w));
    if (zero_grad && !t.skip_zero) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < t.numel; i += blockDim.x) {
      float p = t.p[i], m = t.mom[i], v = t.mom2[i];
      upd(p, t.g[i], m, v);
      t.p[i] = p;
      t.mom[i] = m;
      t.mom2[i] = v;
      if (t.p_bf16 != nullptr) reinterpret_cast<__nv_bfloat16*>(t.p_bf16)[i] = __float2bfloat16(p);
      if (zero_grad) t.g[i] = 0.f;
    }
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d,
                                     long long n) {
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(s)[i];
    reinterpret_cast<uint2*>(d)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x)
      d[i] = __float2bfloat16(s[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d,
                                     long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    d[i] = __bfloat162float(s[i]);
}

__global__ void __launch_bounds__(256)
dgelu_mul_kernel(const uint4* __restrict__ g, const uint4* __restrict__ h, uint4* __restrict__ y,
                 long long n8) {
  pdl_wait();
  pdl_launch_dependents();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 a = g[i];
    const uint4 b = h[i];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = unpack_bf16x2(aw[t]);
      const float2 z = unpack_bf16x2(bw[t]);
      o[t] = pack_bf16x2(x.x * dgelu_erf(z.x), x.y * dgelu_erf(z.y));
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------
// runtime helpers
// ------------------------------------------------------------------------------------------
__global__ void advance_counter_kernel(uint64_t* c, uint64_t inc) { *c += inc; }
__global__ void advance_epoch_kernel(uint32_t* c, uint32_t inc) { *c += inc; }
__global__ void signal_flags_kernel(uint32_t* flags, int n, uint32_t inc) {
  __threadfence_system();
  for (int i = threadIdx.x; i < n; i += blockDim.x) red_release_sys_add(flags + i, inc);
}
__global__ void wait_flags_kernel(const uint32_t* flags, int n, const uint32_t* epoch,
                                  uint32_t mult, int* error_flag) {
  const uint32_t target = (*epoch) * mult;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!wait_flag_ge(flags + i, target, kFlagTimeoutNs)) {
      if (error_flag) atomicExch(error_flag, 3);
    }
  }
}
__global__ void spin_ns_kernel(uint64_t ns) {
  const uint64_t t0 = globaltimer_ns();
  while (globaltimer_ns() - t0 < ns) __nanosleep(200);
}
__global__ void record_time_kernel(uint64_t* slot) { *slot = globaltimer_ns(); }
__global__ void spin_factor_kernel(const uint64_t* t_start, float factor) {
  const uint64_t now = globaltimer_ns();
  const uint64_t elapsed = now - *t_start;
  uint64_t budget = static_cast<uint64_t>(static_cast<double>(elapsed) * factor);
  if (budget > 2000000000ull) budget = 2000000000ull;  // never hold the GPU > 2 s
  while (globaltimer_ns() - now < budget) __nanosleep(200);
}
__global__ void __launch_bounds__(256)
peer_copy_signal_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16,
                        uint32_t* flags, int n_flags, uint32_t inc, unsigned int* done_counter) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *done_counter = 0;
      __threadfence_system();
      for (int i = 0; i < n_flags; ++i) red_release_sys_add(flags + i, inc);
    }
  }
}

inline int grid_for_rows(int rows, int warps_per_block, int max_blocks) {
  int g = (rows + warps_per_block - 1) / warps_per_block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return g;
}

}  // namespace

#define SKY_LAUNCH_CHECK() return static_cast<int>(cudaGetLastError())

int launch_layernorm_fwd(const LayerNormFwdArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return 0;
  if (a.H % 8 != 0 || a.H > kLnMaxVec * 256) return 910;
  const int grid = (a.M + kLnRowsPerCta - 1) / kLnRowsPerCta;
  const int nv = (a.H / 8 + 31) / 32;
#define SKY_LN_FWD(NV)                                                                          \
  launch_pdl(layernorm_fwd_kernel<NV>, dim3(grid), dim3(256), 0, stream,                        \
      reinterpret_cast<const __nv_bfloat16*>(a.z), reinterpret_cast<__nv_bfloat16*>(a.y), a.mean, \
      a.rstd, a.gamma, a.beta, a.M, a.H, a.eps, a.wait_flags, a.wait_epoch, a.wait_mult,        \
      a.error_flag, a.signal_flags)
  if (nv <= 1) SKY_LN_FWD(1);
  else if (nv <= 2) SKY_LN_FWD(2);
  else if (nv <= 4) SKY_LN_FWD(4);
  else SKY_LN_FWD(8);
#undef SKY_LN_FWD
  SKY_LAUNCH_CHECK();
}

int launch_layernorm_bwd(const LayerNormBwdArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return 0;
  if (a.H % 8 != 0 || a.H > kLnMaxVec * 256) return 911;
  const int grid = grid_for_rows(a.M, 8, 148 * 8);
  const int nv = (a.H / 8 + 31) / 32;
#define SKY_LN_BWD(NV)                                                                           \
  launch_pdl(layernorm_bwd_kernel<NV>, dim3(grid), dim3(256), 0, stream,                         \
      reinterpret_cast<const __nv_bfloat16*>(a.dy), reinterpret_cast<const __nv_bfloat16*>(a.z), \
      a.mean, a.rstd, a.gamma, reinterpret_cast<__nv_bfloat16*>(a.dz),                           \
      reinterpret_cast<__nv_bfloat16*>(a.dz_dropped), a.M, a.H, a.dropout_p, a.rng_state,        \
      a.rng_stream, a.wait_flags, a.wait_epoch, a.wait_mult, a.error_flag)
  if (nv <= 1) SKY_LN_BWD(1);
  else if (nv <= 2) SKY_LN_BWD(2);
  else if (nv <= 4) SKY_LN_BWD(4);
  else SKY_LN_BWD(8);
#undef SKY_LN_BWD
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  // NOTE: when gated on peer flags, the row kernel above has already waited for every panel.
  if (a.dgamma != nullptr || a.dbeta != nullptr)
    return launch_ln_param_grad(a.dy, a.z, a.mean, a.rstd, a.dgamma, a.dbeta, a.M, a.H, nullptr,
                                nullptr, stream);
  SKY_LAUNCH_CHECK();
}

// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy.  A separate entry point so that the engine
// can take it off the input-gradient critical path (it is queued with the weight gradients).
int launch_ln_param_grad(const void* dy, const void* z, const float* mean, const float* rstd,
                         float* dgamma, float* dbeta, int M, int H, const void* x2, float* out2,
                         cudaStream_t stream) {
  if (M <= 0) return 0;
  dim3 g2((H + 255) / 256, 1);
  int ry = (M + 63) / 64;
  if (ry > 64) ry = 64;
  if (ry < 1) ry = 1;
  g2.y = ry;
  launch_pdl(ln_param_grad_kernel, g2, dim3(256), 0, stream,
             reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(z),
             mean, rstd, dgamma, dbeta, M, H, reinterpret_cast<const __nv_bfloat16*>(x2), out2);
  SKY_LAUNCH_CHECK();
}

int launch_dgelu_mul(const void* g, const void* h, void* y, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n % 8 != 0) return 914;
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(dgelu_mul_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream,
      reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(h),
      reinterpret_cast<uint4*>(y), n / 8);
  SKY_LAUNCH_CHECK();
}

int launch_colsum(const void* x, int M, int N, int ldx, float* out, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (N % 8 != 0) return 912;
  dim3 grid((N + 255) / 256, 1);
  int ry = (M + 63) / 64;
  if (ry > 64) ry = 64;
  grid.y = ry;
  launch_pdl(colsum_kernel, grid, dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), M, N, static_cast<long long>(ldx),
                                          out);
  SKY_LAUNCH_CHECK();
}

int launch_embed_fwd(const EmbedArgs& a, cudaStream_t stream) {
  const int BS = a.B * a.S;
  if (BS <= 0) return 0;
  if (a.H % 4 != 0) return 913;
  const int grid = grid_for_rows(BS, 8, 148 * 8);
  embed_fwd_kernel<<<grid, 256, 0, stream>>>(
      a.input_ids, a.token_type, a.attn_mask, a.word, a.pos, a.type, a.gamma, a.beta,
      reinterpret_cast<__nv_bfloat16*>(a.out), a.ext_mask, a.mean, a.rstd, BS, a.S, a.H, a.eps,
      a.dropout_p, a.rng_state, a.rng_stream);
  SKY_LAUNCH_CHECK();
}

int launch_embed_bwd(const EmbedBwdArgs& a, cudaStream_t stream) {
  const int BS = a.B * a.S;
  if (BS <= 0) return 0;
  if (a.H % 4 != 0) return 913;
  const int grid = grid_for_rows(BS, 8, 148 * 2);
  int t_smem = a.type_rows < 4 ? a.type_rows : 4;
  if (t_smem < 0) t_smem = 0;
  const size_t smem = static_cast<size_t>(2 + t_smem) * a.H * sizeof(float);
  if (smem > 48 * 1024) {
    t_smem = 0;
  }
  const size_t smem_used = static_cast<size_t>(2 + t_smem) * a.H * sizeof(float);
  if (smem_used > 48 * 1024) return 915;
  embed_bwd_kernel<<<grid, 256, smem_used, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a.dout), a.input_ids, a.token_type, a.word, a.pos,
      a.type, a.gamma, a.mean, a.rstd, a.dword, a.dpos, a.dtype_, a.dgamma, a.dbeta, BS, a.S, a.H,
      a.dropout_p, a.rng_state, a.rng_stream, a.wait_flags, a.wait_epoch, a.wait_mult,
      a.error_flag, t_smem);
  SKY_LAUNCH_CHECK();
}

int launch_small_linear_fwd(const void* x, bool x_bf16, int ldx, const float* w, const float* b,
                            float* y, int M, int N, int K, int act_tanh, float dropout_p,
                            const uint64_t* rng_state, uint32_t rng_stream, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  const long long outs = static_cast<long long>(M) * N;
  const int grid = static_cast<int>((outs + 3) / 4);
  small_linear_fwd_kernel<<<grid, 128, 0, stream>>>(x, x_bf16, ldx, w, b, y, M, N, K, act_tanh,
                                                    dropout_p, rng_state, rng_stream);
  SKY_LAUNCH_CHECK();
}

int launch_small_linear_bwd(const void* x, bool x_bf16, int ldx, const float* w, const float* y,
                            const float* dy, void* dx, bool dx_bf16, int lddx, float* dw, float* db,
                            int M, int N, int K, int act_tanh, float dropout_p,
                            const uint64_t* rng_state, uint32_t rng_stream, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (dx != nullptr) {
    dim3 grid((K + 127) / 128, (M + 7) / 8);
    small_linear_dx_kernel<<<grid, 128, 0, stream>>>(w, y, dy, dx, dx_bf16, lddx, M, N, K,
                                                     act_tanh, dropout_p, rng_state, rng_stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (dw != nullptr) {
    dim3 grid((K + 127) / 128, N);
    small_linear_dw_kernel<<<grid, 128, 0, stream>>>(x, x_bf16, ldx, y, dy, dw, db, M, N, K,
                                                     act_tanh, dropout_p, rng_state, rng_stream);
  }
  SKY_LAUNCH_CHECK();
}

int launch_softmax_ce(const float* logits, const int64_t* labels, float* loss, float* dlogits,
                      int M, int C, float grad_scale, float* loss_acc, cudaStream_t stream) {
  softmax_ce_kernel<<<1, 256, 0, stream>>>(logits, labels, loss, dlogits, M, C, grad_scale,
                                           loss_acc);
  SKY_LAUNCH_CHECK();
}

int launch_sgd_multi(const SgdTensor* d_tensors, int n_tensors, long long max_numel, float lr,
                     float momentum, float weight_decay, float grad_scale, bool zero_grad,
                     cudaStream_t stream) {
  if (n_tensors <= 0) return 0;
  long long blocks = (max_numel / 4 + 255) / 256;
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), n_tensors);
  sgd_multi_kernel<<<grid, 256, 0, stream>>>(d_tensors, lr, momentum, weight_decay, grad_scale,
                                             zero_grad ? 1 : 0);
  SKY_LAUNCH_CHECK();
}

int launch_adam_multi(const SgdTensor* d_tensors, int n_tensors, long long max_numel, float lr,
                      float beta1, float beta2, float eps, float weight_decay, bool decoupled,
                      const uint64_t* step, float grad_scale, bool zero_grad, cudaStream_t stream) {
  if (n_tensors <= 0) return 0;
  if (step == nullptr) return 906;
  long long blocks = (max_numel / 4 + 255) / 256;
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), n_tensors);
  adam_multi_kernel<<<grid, 256, 0, stream>>>(d_tensors, lr, beta1, beta2, eps, weight_decay,
                                              decoupled ? 1 : 0, step, grad_scale,
                                              zero_grad ? 1 : 0);
  SKY_LAUNCH_CHECK();
}

int launch_cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  SKY_LAUNCH_CHECK();
}
int launch_cast_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  cast_bf16_f32_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  SKY_LAUNCH_CHECK();
}

int launch_advance_counter(uint64_t* counter, uint64_t inc, cudaStream_t stream) {
  advance_counter_kernel<<<1, 1, 0, stream>>>(counter, inc);
  SKY_LAUNCH_CHECK();
}
int launch_advance_epoch(uint32_t* counter, uint32_t inc, cudaStream_t stream) {
  advance_epoch_kernel<<<1, 1, 0, stream>>>(counter, inc);
  SKY_LAUNCH_CHECK();
}
int launch_signal_flags(uint32_t* flags, int n, uint32_t inc, cudaStream_t stream) {
  signal_flags_kernel<<<1, 64, 0, stream>>>(flags, n, inc);
  SKY_LAUNCH_CHECK();
}
int launch_wait_flags(const uint32_t* flags, int n, const uint32_t* epoch, uint32_t mult,
                      int* error_flag, cudaStream_t stream) {
  wait_flags_kernel<<<1, 64, 0, stream>>>(flags, n, epoch, mult, error_flag);
  SKY_LAUNCH_CHECK();
}
int launch_spin_ns(uint64_t ns, cudaStream_t stream) {
  spin_ns_kernel<<<1, 1, 0, stream>>>(ns);
  SKY_LAUNCH_CHECK();
}
int launch_record_time(uint64_t* slot, cudaStream_t stream) {
  record_time_kernel<<<1, 1, 0, stream>>>(slot);
  SKY_LAUNCH_CHECK();
}
int launch_spin_factor(const uint64_t* t_start_slot, float factor, cudaStream_t stream) {
  spin_factor_kernel<<<1, 1, 0, stream>>>(t_start_slot, factor);
  SKY_LAUNCH_CHECK();
}

int launch_peer_copy_signal(const void* src, void* dst, long long nbytes, uint32_t* flags,
                            int n_flags, uint32_t inc, cudaStream_t stream) {
  static unsigned int* d_counter = nullptr;
  if (d_counter == nullptr) {
    if (cudaMalloc(&d_counter, sizeof(unsigned int)) != cudaSuccess) return 920;
    cudaMemset(d_counter, 0, sizeof(unsigned int));
  }
  if (nbytes % 16 != 0) return 921;
  long long n16 = nbytes / 16;
  long long blocks = (n16 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  peer_copy_signal_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), n16, flags, n_flags, inc,
      d_counter);
  SKY_LAUNCH_CHECK();
}

}  // namespace sky
