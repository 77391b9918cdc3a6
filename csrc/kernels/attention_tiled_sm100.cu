// Self-attention for ANY sequence length (S % 8 == 0, head_dim = 64) on tcgen05: flash-style
// tiling over 128-key blocks, so the native path covers everything the reference's eager attention
// does (scaelum/model/bert_layers.py:249-275 materialises [B, h, S, S] for any S <= 512).  S = 128
// keeps its specialised single-tile kernels (attention_sm100.cu); this file is the general case.
//
//   forward   one CTA per (batch, head, 128-query tile), loop over key tiles:
//               S = Q K_j^T (UMMA 128x128x64) -> TMEM -> online softmax in registers (running row
//               max m, row sum l, + additive mask, + dropout) -> P_j (bf16, swizzled smem) ->
//               O_j = P_j V_j (UMMA 128x64x128) -> registers: O = O * 2^(m_old - m_new) + O_j.
//               K / V tiles are double buffered: tile j+1 streams in while tile j is processed.
//   backward  two kernels, both recompute P from the saved log-sum-exp:
//               dQ  : one CTA per query tile, loop over key tiles: dS_j -> dQ += dS_j K_j
//               dKV : one CTA per key tile, loop over query tiles: dV += P^T dO, dK += dS^T Q
//                     (accumulated in TMEM across the loop: 4 accumulators = 384 columns)
//             no atomics, deterministic; the price is computing S and dP twice.
//
// Ragged tiles: Q / K / V / dO boxes are fetched through a 3D tensor map (column, position,
// batch), so rows past the end of a sequence arrive as zeros; keys past S get an additive -1e30
// (probability exactly 0), queries past S are computed on zeros and never stored.
// Dropout indexing is the same flat [B, heads, S, S] element index as the S = 128 kernels and the
// host replica (skycomputing_b200/ops/dropout_ref.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "api.h"
#include "launch_util.h"
#include "sm100_ptx.cuh"

namespace sky {

int make_tmap_bf16_3d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t middle,
                      uint64_t outer, uint64_t ld_middle_elems, uint64_t ld_outer_elems,
                      uint32_t box_inner, uint32_t box_middle);

namespace {

constexpr int kT = 128;              // queries / keys per tile
constexpr int kD = 64;               // head dim
constexpr int kTile = kT * kD * 2;   // 16 KB: one [128 x 64] bf16 tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNegBig = -1.0e30f;  // additive mask of keys past the end of the sequence

__device__ __forceinline__ uint32_t sw128_offset(int row, int col) {
  const uint32_t chunk = static_cast<uint32_t>(col >> 3) ^ static_cast<uint32_t>(row & 7);
  return static_cast<uint32_t>(row) * 128u + (chunk << 4) + static_cast<uint32_t>(col & 7) * 2u;
}

struct AttnTDev {
  const float* mask;            // [B, S] additive or null
  __nv_bfloat16* ctx;           // fwd: output [B*S, H]
  float* lse;                   // [B*heads, S] log2-domain log-sum-exp
  const __nv_bfloat16* ctx_in;  // bwd: saved forward output
  const __nv_bfloat16* dctx;    // bwd: gradient of the attention output
  __nv_bfloat16* dqkv;          // bwd: [B*S, 3H]
  int S, heads, H, n_tiles;
  float scale, dropout_p;
  const uint64_t* rng_state;
  uint32_t rng_stream;
};

__device__ __forceinline__ void store_row64(__nv_bfloat16* dst, const float (&o)[64], float mul) {
#pragma unroll
  for (int j = 0; j < 64; j += 8) {
    uint4 pk;
    pk.x = pack_bf16x2(o[j] * mul, o[j + 1] * mul);
    pk.y = pack_bf16x2(o[j + 2] * mul, o[j + 3] * mul);
    pk.z = pack_bf16x2(o[j + 4] * mul, o[j + 5] * mul);
    pk.w = pack_bf16x2(o[j + 6] * mul, o[j + 7] * mul);
    *reinterpret_cast<uint4*>(dst + j) = pk;
  }
}

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
constexpr int kFwdSmem = 7 * kTile + 512 + 128 + 1024;  // Q, 2 x (K, V), P (2 blocks), mask, barriers

__global__ void __launch_bounds__(128, 1)
attention_fwd_tiled_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnTDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kTile;      // [2 buffers][K, V]
  uint8_t* sP = smem + 5 * kTile;   // two [128 x 64] K-blocks
  float* sMask = reinterpret_cast<float*>(smem + 7 * kTile);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTile + 512);
  uint64_t* bar_q = &bars[0];
  uint64_t* bar_kv = &bars[1];      // [2]
  uint64_t* bar_s = &bars[3];
  uint64_t* bar_o = &bars[4];
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int qt = blockIdx.x % p.n_tiles;
  const int bh = blockIdx.x / p.n_tiles;
  const int b = bh / p.heads;
  const int h = bh % p.heads;
  const int q = qt * kT + tid;      // my query position
  const bool q_ok = q < p.S;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  pdl_wait();
  pdl_launch_dependents();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  const bool issuer = warp == 0 && elect_one_sync();
  if (issuer) {
    mbar_expect_tx(bar_q, kTile);
    tma_load_3d(sQ, &tmap_qkv, bar_q, h * kD, qt * kT, b);
    mbar_expect_tx(&bar_kv[0], 2 * kTile);
    tma_load_3d(sKV, &tmap_qkv, &bar_kv[0], p.H + h * kD, 0, b);
    tma_load_3d(sKV + kTile, &tmap_qkv, &bar_kv[0], 2 * p.H + h * kD, 0, b);
  }

  const float sc = p.scale * kLog2e;
  const bool has_dropout = p.dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  if (has_dropout) {
    seed = dropout_seed(p.rng_state, p.rng_stream);
    thr16 = static_cast<uint32_t>(p.dropout_p * 65536.f);
  }
  const uint64_t drop_row = (static_cast<uint64_t>(bh) * p.S + (q_ok ? q : 0)) * p.S;

  float o_acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) o_acc[j] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < p.n_tiles; ++j) {
    const int buf = j & 1;
    uint8_t* sK = sKV + buf * 2 * kTile;
    uint8_t* sV = sK + kTile;
    // additive mask of this key tile (log2 domain); keys past the sequence end get -1e30
    {
      const int k = j * kT + tid;
      sMask[tid] = k < p.S ? (p.mask ? p.mask[b * p.S + k] * kLog2e : 0.f) : kNegBig;
    }
    if (issuer) {
      if (j + 1 < p.n_tiles) {   // the other buffer's tile (j - 1) is fully consumed: prefetch
        uint8_t* nK = sKV + (buf ^ 1) * 2 * kTile;
        mbar_expect_tx(&bar_kv[buf ^ 1], 2 * kTile);
        tma_load_3d(nK, &tmap_qkv, &bar_kv[buf ^ 1], p.H + h * kD, (j + 1) * kT, b);
        tma_load_3d(nK + kTile, &tmap_qkv, &bar_kv[buf ^ 1], 2 * p.H + h * kD, (j + 1) * kT, b);
      }
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[buf], (j >> 1) & 1);
      tcgen05_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 128, false, false);
      const uint64_t da = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t db = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, k);
      umma_commit(bar_s);
    }
    __syncthreads();             // sMask visible
    mbar_wait(bar_s, j & 1);
    tcgen05_fence_after();

    // ---- online softmax over my row ----
    float mx = m_run;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(v[t]) * sc + sMask[c * 32 + t]);
    }
    const float alpha = exp2f(m_run - mx);   // 0 on the first tile (m_run = -inf, mx finite)
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + c * 32, v);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        e[t] = exp2f(__uint_as_float(v[t]) * sc + sMask[c * 32 + t] - mx);
        sum += e[t];
      }
      if (has_dropout) {
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          const uint32_t m =
              dropout_keep4(seed, (drop_row + static_cast<uint64_t>(j * kT + c * 32 + t)) >> 2, thr16);
          e[t] = (m & 1u) ? e[t] : 0.f;
          e[t + 1] = (m & 2u) ? e[t + 1] : 0.f;
          e[t + 2] = (m & 4u) ? e[t + 2] : 0.f;
          e[t + 3] = (m & 8u) ? e[t + 3] : 0.f;
        }
      }
      uint8_t* blk = sP + (c >> 1) * kTile;
#pragma unroll
      for (int t = 0; t < 32; t += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(e[t], e[t + 1]);
        pk.y = pack_bf16x2(e[t + 2], e[t + 3]);
        pk.z = pack_bf16x2(e[t + 4], e[t + 5]);
        pk.w = pack_bf16x2(e[t + 6], e[t + 7]);
        *reinterpret_cast<uint4*>(blk + sw128_offset(tid, (c & 1) * 32 + t)) = pk;
      }
    }
    l_run = l_run * alpha + sum;
    m_run = mx;
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();

    if (issuer) {
      tcgen05_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 64, false, true);
#pragma unroll
      for (int k = 0; k < kT / 16; ++k) {
        const uint64_t da =
            make_smem_desc_sw128(smem_u32(sP + (k >> 2) * kTile) + (k & 3) * 32, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(sV) + k * 2048, kTile, 1024);
        umma_bf16_ss(tmem_base + 128, da, db, idesc, k);
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, j & 1);
    tcgen05_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {   // fully unrolled: o_acc must stay in registers
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + 128 + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t)
        o_acc[c * 32 + t] = fmaf(o_acc[c * 32 + t], alpha, __uint_as_float(v[t]));
    }
    tcgen05_fence_before();
  }

  if (q_ok) {
    const float inv = (has_dropout ? 1.f / (1.f - p.dropout_p) : 1.f) / l_run;
    store_row64(p.ctx + (static_cast<long long>(b) * p.S + q) * p.H + h * kD, o_acc, inv);
    if (p.lse != nullptr) p.lse[static_cast<long long>(bh) * p.S + q] = m_run + log2f(l_run);
  }
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ----------------------------------------------------------------------------------------------
// backward: shared pieces
// ----------------------------------------------------------------------------------------------
// P and dS of one [128 queries x 128 keys] tile from S and dP in TMEM (columns 0..127 / 128..255):
// P = 2^(s * sc + mask - lse), dS = P * (dP * keep / (1-p) - delta) * scale; the dropped / scaled
// P (-> dV) goes to sP, dS to sdS, both as two swizzled [128 x 64] K-blocks indexed [query][key].
struct BwdRow {
  float lse, delta;
  uint64_t drop_row;
};

__device__ __forceinline__ void softmax_bwd_tile(uint32_t trow, const float* sMask, const BwdRow& r,
                                                 int key0, float sc, float scale, bool has_dropout,
                                                 uint64_t seed, uint32_t thr16, float dscale,
                                                 int tid, uint8_t* sP, uint8_t* sdS) {
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {   // 16 keys per step
    uint32_t vs[16], vd[16];
    tmem_ld_32x32b_x16(trow + c * 16, vs);
    tmem_ld_32x32b_x16(trow + 128 + c * 16, vd);
    tmem_ld_wait();
    float pd[16], ds[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      pd[t] = exp2f(__uint_as_float(vs[t]) * sc + sMask[c * 16 + t] - r.lse);
      ds[t] = __uint_as_float(vd[t]);
    }
    if (has_dropout) {
#pragma unroll
      for (int t = 0; t < 16; t += 4) {
        const uint32_t m =
            dropout_keep4(seed, (r.drop_row + static_cast<uint64_t>(key0 + c * 16 + t)) >> 2, thr16);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool keep = (m >> u) & 1u;
          const float pj = pd[t + u];
          ds[t + u] = pj * ((keep ? ds[t + u] * dscale : 0.f) - r.delta) * scale;
          pd[t + u] = keep ? pj * dscale : 0.f;
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t) ds[t] = pd[t] * (ds[t] - r.delta) * scale;
    }
    uint8_t* bp = sP != nullptr ? sP + (c >> 2) * kTile : nullptr;
    uint8_t* bd = sdS + (c >> 2) * kTile;
#pragma unroll
    for (int t = 0; t < 16; t += 8) {
      const uint32_t off = sw128_offset(tid, (c & 3) * 16 + t);
      uint4 pk;
      if (bp != nullptr) {
        pk.x = pack_bf16x2(pd[t], pd[t + 1]);
        pk.y = pack_bf16x2(pd[t + 2], pd[t + 3]);
        pk.z = pack_bf16x2(pd[t + 4], pd[t + 5]);
        pk.w = pack_bf16x2(pd[t + 6], pd[t + 7]);
        *reinterpret_cast<uint4*>(bp + off) = pk;
      }
      pk.x = pack_bf16x2(ds[t], ds[t + 1]);
      pk.y = pack_bf16x2(ds[t + 2], ds[t + 3]);
      pk.z = pack_bf16x2(ds[t + 4], ds[t + 5]);
      pk.w = pack_bf16x2(ds[t + 6], ds[t + 7]);
      *reinterpret_cast<uint4*>(bd + off) = pk;
    }
  }
}

// lse and delta = rowsum(dO * O) of query `q` (zeros / +big for rows past the sequence end, which
// makes P and dS of such rows exactly 0)
__device__ __forceinline__ BwdRow load_bwd_row(const AttnTDev& p, int b, int h, int bh, int q) {
  BwdRow r;
  if (q >= p.S) {
    r.lse = 1.0e30f;
    r.delta = 0.f;
    r.drop_row = 0;
    return r;
  }
  const long long off = (static_cast<long long>(b) * p.S + q) * p.H + h * kD;
  const uint4* o4 = reinterpret_cast<const uint4*>(p.ctx_in + off);
  const uint4* d4 = reinterpret_cast<const uint4*>(p.dctx + off);
  float delta = 0.f;
#pragma unroll
  for (int i = 0; i < kD / 8; ++i) {
    const uint4 a = o4[i];
    const uint4 g = d4[i];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = unpack_bf16x2(aw[t]);
      const float2 y = unpack_bf16x2(gw[t]);
      delta += x.x * y.x + x.y * y.y;
    }
  }
  r.delta = delta;
  r.lse = p.lse[static_cast<long long>(bh) * p.S + q];
  r.drop_row = (static_cast<uint64_t>(bh) * p.S + q) * p.S;
  return r;
}

// ----------------------------------------------------------------------------------------------
// backward: dQ (one CTA per query tile, loop over key tiles)
// ----------------------------------------------------------------------------------------------
constexpr int kBwdQSmem = 6 * kTile + 512 + 128 + 1024;   // Q, dO, K, V, dS (2 blocks)

__global__ void __launch_bounds__(128, 1)
attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                        const __grid_constant__ CUtensorMap tmap_do, const AttnTDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + kTile;
  uint8_t* sK = smem + 2 * kTile;
  uint8_t* sV = smem + 3 * kTile;
  uint8_t* sdS = smem + 4 * kTile;
  float* sMask = reinterpret_cast<float*>(smem + 6 * kTile);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTile + 512);
  uint64_t* bar_q = &bars[0];
  uint64_t* bar_kv = &bars[1];
  uint64_t* bar_s = &bars[2];
  uint64_t* bar_o = &bars[3];
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int qt = blockIdx.x % p.n_tiles;
  const int bh = blockIdx.x / p.n_tiles;
  const int b = bh / p.heads;
  const int h = bh % p.heads;
  const int q = qt * kT + tid;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  pdl_wait();
  pdl_launch_dependents();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const bool issuer = warp == 0 && elect_one_sync();
  if (issuer) {
    mbar_expect_tx(bar_q, 2 * kTile);
    tma_load_3d(sQ, &tmap_qkv, bar_q, h * kD, qt * kT, b);
    tma_load_3d(sdO, &tmap_do, bar_q, h * kD, qt * kT, b);
  }
  const BwdRow row = load_bwd_row(p, b, h, bh, q);
  const float sc = p.scale * kLog2e;
  const bool has_dropout = p.dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float dscale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(p.rng_state, p.rng_stream);
    thr16 = static_cast<uint32_t>(p.dropout_p * 65536.f);
    dscale = 1.f / (1.f - p.dropout_p);
  }
  float dq[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) dq[j] = 0.f;

  for (int j = 0; j < p.n_tiles; ++j) {
    {
      const int k = j * kT + tid;
      sMask[tid] = k < p.S ? (p.mask ? p.mask[b * p.S + k] * kLog2e : 0.f) : kNegBig;
    }
    if (issuer) {
      mbar_expect_tx(bar_kv, 2 * kTile);
      tma_load_3d(sK, &tmap_qkv, bar_kv, p.H + h * kD, j * kT, b);
      tma_load_3d(sV, &tmap_qkv, bar_kv, 2 * p.H + h * kD, j * kT, b);
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, j & 1);
      tcgen05_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 128, false, false);
      const uint64_t dqd = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t dk = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t ddo = make_smem_desc_sw128(smem_u32(sdO), 16, 1024);
      const uint64_t dv = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tmem_base, dqd + 2 * k, dk + 2 * k, idesc, k);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k)
        umma_bf16_ss(tmem_base + 128, ddo + 2 * k, dv + 2 * k, idesc, k);
      umma_commit(bar_s);
    }
    __syncthreads();
    mbar_wait(bar_s, j & 1);
    tcgen05_fence_after();
    softmax_bwd_tile(trow, sMask, row, j * kT, sc, p.scale, has_dropout, seed, thr16, dscale, tid,
                     nullptr, sdS);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (issuer) {
      tcgen05_fence_after();
      // dQ_j[q, d] = sum_key dS[q, key] K[key, d]: A = dS (K-major, 2 K-blocks), B = K (MN-major)
      constexpr uint32_t idesc_q = make_idesc_bf16_f32(128, 64, false, true);
#pragma unroll
      for (int k = 0; k < kT / 16; ++k) {
        const uint64_t da =
            make_smem_desc_sw128(smem_u32(sdS + (k >> 2) * kTile) + (k & 3) * 32, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(sK) + k * 2048, kTile, 1024);
        umma_bf16_ss(tmem_base, da, db, idesc_q, k);   // S is dead: its columns take dQ_j
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, j & 1);
    tcgen05_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {   // fully unrolled: dq must stay in registers
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t) dq[c * 32 + t] += __uint_as_float(v[t]);
    }
    tcgen05_fence_before();
    __syncthreads();   // everybody has read dQ_j (and K / V): the next tile may overwrite them
  }
  if (q < p.S)
    store_row64(p.dqkv + (static_cast<long long>(b) * p.S + q) * (3 * p.H) + h * kD, dq, 1.f);
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ----------------------------------------------------------------------------------------------
// backward: dK, dV (one CTA per key tile, loop over query tiles, accumulators stay in TMEM)
// ----------------------------------------------------------------------------------------------
constexpr int kBwdKVSmem = 8 * kTile + 512 + 128 + 1024;   // K, V, Q, dO, P (2), dS (2)

__global__ void __launch_bounds__(128, 1)
attention_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                         const __grid_constant__ CUtensorMap tmap_do, const AttnTDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTile;
  uint8_t* sQ = smem + 2 * kTile;
  uint8_t* sdO = smem + 3 * kTile;
  uint8_t* sP = smem + 4 * kTile;
  uint8_t* sdS = smem + 6 * kTile;
  float* sMask = reinterpret_cast<float*>(smem + 8 * kTile);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * kTile + 512);
  uint64_t* bar_kv = &bars[0];
  uint64_t* bar_q = &bars[1];
  uint64_t* bar_s = &bars[2];
  uint64_t* bar_o = &bars[3];
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int kt = blockIdx.x % p.n_tiles;
  const int bh = blockIdx.x / p.n_tiles;
  const int b = bh / p.heads;
  const int h = bh % p.heads;
  constexpr uint32_t cdV = 256, cdK = 320;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  pdl_wait();
  pdl_launch_dependents();
  {
    const int k = kt * kT + tid;
    sMask[tid] = k < p.S ? (p.mask ? p.mask[b * p.S + k] * kLog2e : 0.f) : kNegBig;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const bool issuer = warp == 0 && elect_one_sync();
  if (issuer) {
    mbar_expect_tx(bar_kv, 2 * kTile);
    tma_load_3d(sK, &tmap_qkv, bar_kv, p.H + h * kD, kt * kT, b);
    tma_load_3d(sV, &tmap_qkv, bar_kv, 2 * p.H + h * kD, kt * kT, b);
  }
  const float sc = p.scale * kLog2e;
  const bool has_dropout = p.dropout_p > 0.f;
  uint64_t seed = 0;
  uint32_t thr16 = 0;
  float dscale = 1.f;
  if (has_dropout) {
    seed = dropout_seed(p.rng_state, p.rng_stream);
    thr16 = static_cast<uint32_t>(p.dropout_p * 65536.f);
    dscale = 1.f / (1.f - p.dropout_p);
  }

  for (int i = 0; i < p.n_tiles; ++i) {
    if (issuer) {
      mbar_expect_tx(bar_q, 2 * kTile);
      tma_load_3d(sQ, &tmap_qkv, bar_q, h * kD, i * kT, b);
      tma_load_3d(sdO, &tmap_do, bar_q, h * kD, i * kT, b);
    }
    const BwdRow row = load_bwd_row(p, b, h, bh, i * kT + tid);   // overlaps the loads
    if (issuer) {
      if (i == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, i & 1);
      tcgen05_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16_f32(128, 128, false, false);
      const uint64_t dqd = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t dk = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t ddo = make_smem_desc_sw128(smem_u32(sdO), 16, 1024);
      const uint64_t dv = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k) umma_bf16_ss(tmem_base, dqd + 2 * k, dk + 2 * k, idesc, k);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k)
        umma_bf16_ss(tmem_base + 128, ddo + 2 * k, dv + 2 * k, idesc, k);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, i & 1);
    tcgen05_fence_after();
    softmax_bwd_tile(trow, sMask, row, kt * kT, sc, p.scale, has_dropout, seed, thr16, dscale, tid,
                     sP, sdS);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (issuer) {
      tcgen05_fence_after();
      // dV[key, d] += sum_q P[q, key] dO[q, d]; dK[key, d] += sum_q dS[q, key] Q[q, d]:
      // A = P / dS read MN-major (key-chunks one tile apart), B = dO / Q (MN-major)
      constexpr uint32_t idesc_t = make_idesc_bf16_f32(128, 64, true, true);
#pragma unroll
      for (int k = 0; k < kT / 16; ++k) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sP) + k * 2048, kTile, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(sdO) + k * 2048, kTile, 1024);
        umma_bf16_ss(tmem_base + cdV, da, db, idesc_t, (i > 0 || k > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < kT / 16; ++k) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sdS) + k * 2048, kTile, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(sQ) + k * 2048, kTile, 1024);
        umma_bf16_ss(tmem_base + cdK, da, db, idesc_t, (i > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
    // Q / dO / P / dS are read by these MMAs: the next query tile may only overwrite them once
    // they have retired
    mbar_wait(bar_o, i & 1);
    tcgen05_fence_after();
    tcgen05_fence_before();
    __syncthreads();
  }

  tcgen05_fence_after();
  const int key = kt * kT + tid;
  {
    // tcgen05.ld is warp-collective (.sync.aligned): every lane loads, only rows of real keys store
    __nv_bfloat16* grow =
        p.dqkv + (static_cast<long long>(b) * p.S + (key < p.S ? key : 0)) * (3 * p.H) + h * kD;
#pragma unroll
    for (int t = 0; t < 2; ++t) {   // t = 0: dK -> K section, t = 1: dV -> V section
      float o[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + (t == 0 ? cdK : cdV) + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) o[c * 32 + u] = __uint_as_float(v[u]);
      }
      if (key < p.S) store_row64(grow + (t + 1) * p.H, o, 1.f);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int fill_dev(const AttnArgs& a, AttnTDev* d) {
  if (a.head_dim != kD || a.S <= 0 || (a.S % 8) != 0) return 930;
  if (a.dropout_p > 0.f && a.rng_state == nullptr) return 903;
  d->mask = a.mask;
  d->ctx = reinterpret_cast<__nv_bfloat16*>(a.ctx);
  d->lse = a.lse;
  d->ctx_in = reinterpret_cast<const __nv_bfloat16*>(a.ctx);
  d->dctx = reinterpret_cast<const __nv_bfloat16*>(a.dctx);
  d->dqkv = reinterpret_cast<__nv_bfloat16*>(a.dqkv);
  d->S = a.S;
  d->heads = a.heads;
  d->H = a.heads * kD;
  d->n_tiles = (a.S + kT - 1) / kT;
  d->scale = a.scale;
  d->dropout_p = a.dropout_p;
  d->rng_state = a.rng_state;
  d->rng_stream = a.rng_stream;
  return 0;
}

template <class K>
int set_smem(K kern, int bytes, bool* done) {
  if (*done) return 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  *done = true;
  return 0;
}

}  // namespace

bool attention_supported(int S, int head_dim) { return head_dim == kD && S > 0 && (S % 8) == 0; }

int launch_attention_fwd_tiled(const AttnArgs& a, cudaStream_t stream) {
  AttnTDev d;
  int rc = fill_dev(a, &d);
  if (rc) return rc;
  const uint64_t H = d.H;
  CUtensorMap tm;
  rc = make_tmap_bf16_3d(&tm, a.qkv, 3 * H, a.S, a.B, 3 * H, 3 * H * a.S, 64, 128);
  if (rc) return rc;
  static bool attr = false;
  rc = set_smem(attention_fwd_tiled_kernel, kFwdSmem, &attr);
  if (rc) return rc;
  cudaError_t le = launch_pdl(attention_fwd_tiled_kernel, dim3(a.B * a.heads * d.n_tiles),
                              dim3(128), kFwdSmem, stream, tm, d);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

int launch_attention_bwd_tiled(const AttnArgs& a, cudaStream_t stream) {
  AttnTDev d;
  int rc = fill_dev(a, &d);
  if (rc) return rc;
  if (a.lse == nullptr || a.ctx == nullptr || a.dctx == nullptr || a.dqkv == nullptr) return 931;
  const uint64_t H = d.H;
  CUtensorMap tm, tdo;
  rc = make_tmap_bf16_3d(&tm, a.qkv, 3 * H, a.S, a.B, 3 * H, 3 * H * a.S, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tdo, a.dctx, H, a.S, a.B, H, H * a.S, 64, 128);
  if (rc) return rc;
  static bool attr_q = false, attr_kv = false;
  rc = set_smem(attention_bwd_dq_kernel, kBwdQSmem, &attr_q);
  if (rc) return rc;
  rc = set_smem(attention_bwd_dkv_kernel, kBwdKVSmem, &attr_kv);
  if (rc) return rc;
  const dim3 grid(a.B * a.heads * d.n_tiles);
  cudaError_t le = launch_pdl(attention_bwd_dq_kernel, grid, dim3(128), kBwdQSmem, stream, tm, tdo, d);
  if (le != cudaSuccess) return static_cast<int>(le);
  le = launch_pdl(attention_bwd_dkv_kernel, grid, dim3(128), kBwdKVSmem, stream, tm, tdo, d);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace sky
