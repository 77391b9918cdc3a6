// Persistent warp-specialised tcgen05 GEMM for sm_100a (B200).
//
//   out[M,N] = epilogue( sum_k A[m,k] * B[n,k] )      bf16 x bf16 -> fp32 (TMEM) -> bf16 | fp32
//
//   * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a 4..6 deep smem ring,
//   * tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BLOCK_N x 16) issued by ONE thread,
//   * two TMEM accumulator buffers (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
//     main loop of tile i+1,
//   * epilogue warps read TMEM with tcgen05.ld, apply bias / erf-GELU / GELU' / dropout /
//     residual add, and store straight to global memory -- which may be the NEXT PIPELINE STAGE's
//     HBM (peer pointer over NVLink); a per-128-row-panel flag is then bumped with
//     red.release.sys so the consumer GPU can start on that panel while we still compute,
//   * symmetric consumer side: the TMA producer can ld.acquire.sys-poll such flags before it
//     loads the A rows of a panel (GEMM -> GEMM boundary), see GemmArgs in api.h.
//
// Both operands may be K-major or MN-major (UMMA descriptor "major" bits), so forward
// (x W^T), dgrad (dy W) and wgrad (dy^T x) all run on this one kernel without transposes.
//
// Reference parity: replaces every nn.Linear / LinearActivation call in
// scaelum/model/bert_layers.py:60-108,227-229,281,311-313,319 and the autograd backward of those.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "api.h"
#include "launch_util.h"
#include "sm100_ptx.cuh"

namespace sky {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-11 epilogue
constexpr int kEpiWarp0 = 4;
constexpr uint64_t kFlagTimeoutNs = 4000000000ull;  // 4 s

template <int BLOCK_N>
struct Cfg {
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 256 or 512 (power of two)
  static constexpr int kBiasBytes = 2 * BLOCK_N * 4;  // double-buffered bias tile
  static constexpr int kOutStageBytes = 8 * 32 * 80;  // per epilogue warp: 32 rows x (64+16) B
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ +
                                    256 /*barriers*/ + kBiasBytes + kOutStageBytes;
};

struct GemmDev {
  int M, N, K;
  void* out;
  void* out2;
  const float* bias;
  const __nv_bfloat16* aux;
  long long ldo, ldo2, ldaux;
  int act;
  int add_aux;
  int accumulate;
  float dropout_p;
  const uint64_t* rng_state;
  uint32_t rng_stream;
  uint32_t* signal_flags;
  const uint32_t* wait_flags;
  const uint32_t* wait_epoch;
  uint32_t wait_mult;
  int* error_flag;
  int debug;
};

template <int BLOCK_N, bool A_MN, bool B_MN, bool OUT_F32>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmDev p) {
  using C = Cfg<BLOCK_N>;
  constexpr int kStages = C::kStages;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tmem_full_bar = bars + 2 * kStages;    // [2]
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  float* s_bias = reinterpret_cast<float*>(smem + kStages * C::kStageBytes + 256);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_blks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n_blks = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m_blks * num_n_blks;
  const int num_k_blks = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, C::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above (barrier init, TMEM alloc, tensormap prefetch) overlapped the previous
  // kernel's tail; from here on we read what it produced
  pdl_wait();
  pdl_launch_dependents();

  if (warp_idx == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % num_m_blks;
      const int n_blk = tile / num_m_blks;
      if (p.wait_flags != nullptr) {
        const uint32_t target = (*p.wait_epoch) * p.wait_mult;
        if (!wait_flag_ge(p.wait_flags + m_blk, target, kFlagTimeoutNs)) {
          if (p.error_flag) atomicExch(p.error_flag, 1);
        }
        fence_proxy_async();  // peer generic-proxy writes -> our async-proxy (TMA) reads
      }
      for (int kb = 0; kb < num_k_blks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], C::kStageBytes);
        uint8_t* sa = smem_a + stage * C::kABytes;
        uint8_t* sb = smem_b + stage * C::kBBytes;
        if constexpr (!A_MN) {
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
        } else {
#pragma unroll
          for (int c = 0; c < BLOCK_M / 64; ++c)
            tma_load_2d(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[stage],
                        m_blk * BLOCK_M + c * 64, kb * BLOCK_K);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
        } else {
#pragma unroll
          for (int c = 0; c < BLOCK_N / 64; ++c)
            tma_load_2d(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[stage],
                        n_blk * BLOCK_N + c * 64, kb * BLOCK_K);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_idx == 1 && lane == 0) {
    // ===================================== MMA issuer ======================================
    constexpr uint32_t idesc = make_idesc_bf16_f32(BLOCK_M, BLOCK_N, A_MN, B_MN);
    // K-major  SW128: 8-row atoms 1024B apart (SBO), LBO unused (1); K advance = 32B per UMMA_K
    // MN-major SW128: 64-element MN chunks BLOCK_K*128B apart (LBO), 8-k-row groups 1024B apart
    //                 (SBO); K advance = 16 rows * 128B = 2048B per UMMA_K
    constexpr uint32_t a_lbo = A_MN ? BLOCK_K * 128 : 16;
    constexpr uint32_t b_lbo = B_MN ? BLOCK_K * 128 : 16;
    constexpr uint32_t a_kadv = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    constexpr uint32_t b_kadv = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_k_blks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint64_t desc_a =
            make_smem_desc_sw128(smem_u32(smem_a + stage * C::kABytes), a_lbo, 1024);
        const uint64_t desc_b =
            make_smem_desc_sw128(smem_u32(smem_b + stage * C::kBBytes), b_lbo, 1024);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          umma_bf16_ss(tmem_d, desc_a + k * a_kadv, desc_b + k * b_kadv, idesc,
                       (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        if (kb == num_k_blks - 1) umma_commit(&tmem_full_bar[acc]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_idx >= kEpiWarp0) {
    // ====================================== epilogue =======================================
    // 8 warps: warp e handles TMEM lane quadrant (e & 3) and column half (e >> 2) of the tile.
    // Latency hiding (profiles/gemm_epilogue_v1.md: the first version stalled ~80% on
    // long-scoreboard waits for bias/aux global loads with one warp per scheduler):
    //   * bias for the tile is staged once in shared memory (double buffered per accumulator),
    //   * aux (residual / pre-activation) chunks are prefetched one chunk ahead, the first one
    //     before the accumulator is even ready,
    //   * TMEM loads are double buffered: chunk c+1 is in flight while chunk c is processed,
    //   * fp32 gradient accumulation uses red.global.add.v4.f32 (no read-modify-write).
    const int e = warp_idx - kEpiWarp0;
    const int q = e & 3;  // == warp_idx % 4: the TMEM lanes this warp may access
    const int half = e >> 2;
    constexpr int HALF_N = BLOCK_N / 2;
    constexpr int NCH = HALF_N / 32;
    const int epi_tid = threadIdx.x - kEpiWarp0 * 32;  // 0..255
    uint8_t* s_stage = smem + kStages * C::kStageBytes + 256 + C::kBiasBytes + e * (32 * 80);
    const int row_in_tile = q * 32 + lane;
    const bool has_dropout = p.dropout_p > 0.f;
    const bool has_aux = p.aux != nullptr;
    uint64_t seed = 0;
    uint32_t thr16 = 0;
    float drop_scale = 1.f;
    if (has_dropout) {
      seed = dropout_seed(p.rng_state, p.rng_stream);
      thr16 = static_cast<uint32_t>(p.dropout_p * 65536.f);
      drop_scale = 1.f / (1.f - p.dropout_p);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile % num_m_blks;
      const int n_blk = tile / num_m_blks;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long row = static_cast<long long>(m_blk) * BLOCK_M + row_in_tile;
      const bool row_ok = row < p.M;
      const int colbase = n_blk * BLOCK_N + half * HALF_N;
      // ---- stage bias, prefetch first aux chunk (both overlap the MMA main loop) ----
      float* sb = s_bias + acc * BLOCK_N;
      if (epi_tid < BLOCK_N) {
        const int gc = n_blk * BLOCK_N + epi_tid;
        sb[epi_tid] = (p.bias != nullptr && gc < p.N) ? __ldg(p.bias + gc) : 0.f;
      }
      uint4 auxa[4], auxb[4];
      const __nv_bfloat16* aux_row = has_aux ? p.aux + row * p.ldaux + colbase : nullptr;
      auto load_aux = [&](uint4 (&dst)[4], int c) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = colbase + c * 32 + t * 8;
          dst[t] = (has_aux && row_ok && col < p.N)
                       ? *reinterpret_cast<const uint4*>(aux_row + c * 32 + t * 8)
                       : make_uint4(0, 0, 0, 0);
        }
      };
      if (has_aux) load_aux(auxa, 0);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // bias of this tile visible to all 8 warps

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + half * HALF_N;

      // bf16 stores go through a per-warp smem transpose: a lane owns one accumulator ROW, so
      // storing straight from registers makes every STG.128 touch 32 different rows (32 partial
      // sectors per request).  Staged, a warp store covers 8 rows x 64 contiguous bytes.
      // Row stride 80 B keeps both the row-wise writes and the 4-lanes-per-row reads (nearly)
      // bank-conflict free.
      const long long warp_row0 = static_cast<long long>(m_blk) * BLOCK_M + q * 32;
      auto store_bf16_chunk = [&](const float (&f)[32], __nv_bfloat16* gout, long long ld,
                                  int col0) {
        uint8_t* mine = s_stage + lane * 80;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          pk.x = pack_bf16x2(f[j], f[j + 1]);
          pk.y = pack_bf16x2(f[j + 2], f[j + 3]);
          pk.z = pack_bf16x2(f[j + 4], f[j + 5]);
          pk.w = pack_bf16x2(f[j + 6], f[j + 7]);
          *reinterpret_cast<uint4*>(mine + j * 2) = pk;
        }
        __syncwarp();
        const int piece = lane & 3;
        if (col0 + piece * 8 < p.N) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2);
            if (warp_row0 + r < p.M) {
              const uint4 val = *reinterpret_cast<const uint4*>(s_stage + r * 80 + piece * 16);
              *reinterpret_cast<uint4*>(gout + (warp_row0 + r) * ld + col0 + piece * 8) = val;
            }
          }
        }
        __syncwarp();
      };

      auto process = [&](uint32_t (&v)[32], uint4 (&ax)[4], int c) {
        const int col0 = colbase + c * 32;
        if (col0 >= p.N || p.debug == 1) return;  // warp-uniform
        float f[32];
        const float4* sb4 = reinterpret_cast<const float4*>(sb + half * HALF_N + c * 32);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = sb4[j >> 2];  // smem broadcast
          f[j] = __uint_as_float(v[j]) + bb.x;
          f[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
          f[j + 2] = __uint_as_float(v[j + 2]) + bb.z;
          f[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
        }
        if (p.act == ACT_GELU) {
          if (p.out2 != nullptr && p.debug != 2)
            store_bf16_chunk(f, reinterpret_cast<__nv_bfloat16*>(p.out2), p.ldo2, col0);
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
        } else if (p.act == ACT_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j]);
        }
        if (has_dropout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const uint64_t idx4 = (static_cast<uint64_t>(row) * p.N + col0 + j) >> 2;
            const uint32_t m = dropout_keep4(seed, idx4, thr16);
            f[j] = (m & 1u) ? f[j] * drop_scale : 0.f;
            f[j + 1] = (m & 2u) ? f[j + 1] * drop_scale : 0.f;
            f[j + 2] = (m & 4u) ? f[j + 2] * drop_scale : 0.f;
            f[j + 3] = (m & 8u) ? f[j + 3] * drop_scale : 0.f;
          }
        }
        if (has_aux) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t aw[4] = {ax[t].x, ax[t].y, ax[t].z, ax[t].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 x = unpack_bf16x2(aw[u]);
              const int j = t * 8 + u * 2;
              if (p.act == ACT_DGELU_MUL_AUX) {
                f[j] *= dgelu_fast(x.x);
                f[j + 1] *= dgelu_fast(x.y);
              } else if (p.add_aux) {
                f[j] += x.x;
                f[j + 1] += x.y;
              }
            }
          }
        }
        if (p.debug != 2) {
          if constexpr (OUT_F32) {
            if (row_ok) {
              float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (col0 + j < p.N) {
                  if (p.accumulate)
                    red_add_v4_f32(o + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
                  else
                    *reinterpret_cast<float4*>(o + j) =
                        make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
              }
            }
          } else {
            store_bf16_chunk(f, reinterpret_cast<__nv_bfloat16*>(p.out), p.ldo, col0);
          }
        }
      };

      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(taddr, va);
#pragma unroll 1
      for (int c = 0; c < NCH; c += 2) {
        tmem_ld_wait();                                // chunk c landed in va
        tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);  // chunk c+1 in flight
        if (has_aux) load_aux(auxb, c + 1);
        process(va, auxa, c);
        tmem_ld_wait();  // chunk c+1 landed in vb
        if (c + 2 < NCH) {
          tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
          if (has_aux) load_aux(auxa, c + 2);
        }
        process(vb, auxb, c + 1);
      }
      // release the TMEM buffer back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (p.signal_flags != nullptr) {
        // publish this tile: all 256 epilogue threads' stores -> barrier -> one release.sys
        __threadfence_system();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (epi_tid == 0) red_release_sys_add(p.signal_flags + m_blk, 1u);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e =
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

}  // namespace

// 2D bf16 tensor map: inner (contiguous) extent `inner`, `outer` rows of stride ld elements.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return 901;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + static_cast<int>(r);
}

int gemm_pick_block_n(int M, int N) {
  // 256-wide tiles halve B re-reads and give the best tensor-pipe duty cycle; fall back to 128
  // when that would leave most SMs idle or N is small.
  const long long tiles256 = static_cast<long long>((M + 127) / 128) * ((N + 255) / 256);
  if (N % 256 != 0 && N <= 128) return 128;
  if (tiles256 < 96) return 128;
  return 256;
}

int gemm_tiles_per_panel(int N, int block_n) { return (N + block_n - 1) / block_n; }

namespace {
static int g_num_sms = 0;

template <int BLOCK_N, bool A_MN, bool B_MN, bool OUT_F32>
int launch_inst(const GemmArgs& a, const GemmDev& dev, cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  CUtensorMap tma, tmb;
  int rc;
  if (!A_MN)
    rc = make_tmap_bf16_2d(&tma, a.A, a.K, a.M, a.lda, BLOCK_K, BLOCK_M);
  else
    rc = make_tmap_bf16_2d(&tma, a.A, a.M, a.K, a.lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!B_MN)
    rc = make_tmap_bf16_2d(&tmb, a.B, a.K, a.N, a.ldb, BLOCK_K, BLOCK_N);
  else
    rc = make_tmap_bf16_2d(&tmb, a.B, a.N, a.K, a.ldb, 64, BLOCK_K);
  if (rc) return rc;

  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, OUT_F32>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev_id);
  }
  const int num_tiles = ((a.M + BLOCK_M - 1) / BLOCK_M) * ((a.N + BLOCK_N - 1) / BLOCK_N);
  int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
  if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(kNumThreads), C::kSmemBytes, stream, tma, tmb, dev);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

template <int BLOCK_N>
int dispatch_major(const GemmArgs& a, const GemmDev& dev, cudaStream_t stream) {
  if (!a.a_mn && !a.b_mn)
    return a.out_f32 ? launch_inst<BLOCK_N, false, false, true>(a, dev, stream)
                     : launch_inst<BLOCK_N, false, false, false>(a, dev, stream);
  if (!a.a_mn && a.b_mn)
    return a.out_f32 ? launch_inst<BLOCK_N, false, true, true>(a, dev, stream)
                     : launch_inst<BLOCK_N, false, true, false>(a, dev, stream);
  if (a.a_mn && a.b_mn)
    return a.out_f32 ? launch_inst<BLOCK_N, true, true, true>(a, dev, stream)
                     : launch_inst<BLOCK_N, true, true, false>(a, dev, stream);
  return a.out_f32 ? launch_inst<BLOCK_N, true, false, true>(a, dev, stream)
                   : launch_inst<BLOCK_N, true, false, false>(a, dev, stream);
}
}  // namespace

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return 0;
  if ((a.N % 8) != 0 || (a.K % 8) != 0 || (a.lda % 8) != 0 || (a.ldb % 8) != 0) return 902;
  if (a.a_mn && (a.M % 8) != 0) return 902;
  if (a.dropout_p > 0.f && a.rng_state == nullptr) return 903;
  if (a.accumulate && !a.out_f32) return 904;
  GemmDev d;
  d.M = a.M;
  d.N = a.N;
  d.K = a.K;
  d.out = a.out;
  d.out2 = a.out2;
  d.bias = a.bias;
  d.aux = reinterpret_cast<const __nv_bfloat16*>(a.aux);
  d.ldo = a.ldo;
  d.ldo2 = a.ldo2;
  d.ldaux = a.ldaux;
  d.act = a.act;
  d.add_aux = a.add_aux ? 1 : 0;
  d.accumulate = a.accumulate ? 1 : 0;
  d.dropout_p = a.dropout_p;
  d.rng_state = a.rng_state;
  d.rng_stream = a.rng_stream;
  d.signal_flags = a.signal_flags;
  d.wait_flags = a.wait_flags;
  d.wait_epoch = a.wait_epoch;
  d.wait_mult = a.wait_mult;
  d.error_flag = a.error_flag;
  d.debug = a.debug;
  const int bn = a.block_n ? a.block_n : gemm_pick_block_n(a.M, a.N);
  if (bn == 256) return dispatch_major<256>(a, d, stream);
  if (bn == 128) return dispatch_major<128>(a, d, stream);
  return 905;
}

}  // namespace sky
