// esm_b200 — persistent warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   C[M,N] = A[M,K] (fp16, K-major) x B[N,K]^T (fp16, K-major = nn.Linear.weight layout), fp32 accumulate in TMEM.
//
// One CTA per SM, 128x256 output tile, K streamed in 64-element (128-byte, one swizzle atom) slabs through a
// 4-stage TMA->smem ring; a single thread issues tcgen05.mma (UMMA 128x256x16); the 2x256-column TMEM
// accumulator is double buffered so the 8 epilogue warps drain tile i while tile i+1 is being multiplied.
//
// Epilogues (reference lines they replace, /root/reference/esm/...):
//   QKV_ROPE      multihead_attention.py:258-261 (q/k/v Linear + bias, q *= d^-1/2) + :354-355 /
//                 rotary_embedding.py:11-20 (rotate-half RoPE on q,k) -> fp16 [M,3E]
//   BIAS_RESIDUAL multihead_attention.py:395 + modules.py:134, and modules.py:139-140
//                 (Linear + bias, residual add) -> fp32 residual stream updated in place
//   BIAS_GELU     modules.py:138 + :17-24 (fc1 + exact erf GELU) -> fp16 [M,F]
//   BIAS_F32      plain Linear + bias -> fp32 (LM-head dense, modules.py:308)
#pragma once

#include "common.cuh"

namespace esmb200 {

enum : int { EPI_QKV_ROPE = 0, EPI_BIAS_RESIDUAL = 1, EPI_BIAS_GELU = 2, EPI_BIAS_F32 = 3, EPI_BIAS_GELU_F32 = 4,
              EPI_NONE = 5 /* profiling only: accumulators are discarded */,
              EPI_LDONLY = 6 /* profiling only: accumulators are read from TMEM and discarded */,
              EPI_LD_X16 = 7, EPI_LD_4WARPS = 8, EPI_LD_BATCH = 9 /* profiling only: TMEM read pattern variants */,
              EPI_GELU_MATHONLY = 10, EPI_F16_STOREONLY = 11 /* profiling only: halves of the fc1 epilogue */,
              EPI_FMA_MATHONLY = 12 /* profiling only: 15 dependent FMAs per element instead of GELU, no MUFU */ };

struct GemmParams {
  int M, N, K;
  const float* bias;      // [N] fp32
  void* out;              // fp16 or fp32, row-major [M, ldo]
  int ldo;
  // EPI_QKV_ROPE only
  const float* rope_cos;  // [T, 32] fp32 (angle t * inv_freq[j], j < d/2)
  const float* rope_sin;
  int T;                  // tokens per sequence: position of row r is r % T
  int E;                  // embed dim: columns [0,E) = q, [E,2E) = k, [2E,3E) = v
  float q_scale;          // head_dim^-0.5
};

namespace gemm_cfg {
constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 fp16 = 128 B = one SWIZZLE_128B atom
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512
constexpr int NUM_THREADS = 384;                 // warps 0-3: TMA / MMA / TMEM alloc / spare, warps 4-11: epilogue
constexpr int EPI_THREADS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
}  // namespace gemm_cfg

// Exact-erf GELU x * 0.5 * (1 + erf(x / sqrt 2)) (esm/modules.py:17-24) with erf from Abramowitz & Stegun 7.1.26
// (|erf error| <= 1.5e-7, far below the fp16 rounding of the stored activation): 15 instructions, 2 MUFU,
// against ~25 for libdevice erff. For x >= 0: x - x*q, for x < 0: x*q with q = 0.5 * poly(t) * exp(-x^2/2),
// t = 1 / (1 + p|x|/sqrt 2).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  poly *= t;
  const float e = ex2_approx(z * (z * -1.4426950408889634f));
  const float q = poly * e;
  return x * (x >= 0.f ? 1.0f - q : q);
}

template <int EPI>
__global__ void __launch_bounds__(gemm_cfg::NUM_THREADS, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p) {
  using namespace gemm_cfg;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                        // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;              // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;          // [ACC_STAGES] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + ACC_STAGES;  // [ACC_STAGES] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES);

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;

  const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M, BLOCK_N, false);
      uint32_t stage = 0, phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const uint32_t as = iter & 1, aph = (iter >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES), 1024, 0);
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * B_STAGE_BYTES), 1024, 0);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // +32 bytes per UMMA_K inside the 128-byte swizzle atom -> +2 in the (addr >> 4) field
            umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (kb == num_kb - 1) tc_commit(&tfull_bar[as]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const uint32_t ew = warp - 4;
    const uint32_t quarter = warp % 4;  // TMEM lane quarter this warp may access
    const uint32_t chalf = ew / 4;      // which 128-column half of the tile
    const uint32_t row_local = quarter * 32 + lane;
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      const uint32_t as = iter & 1, aph = (iter >> 1) & 1;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + row_local;
      const bool row_ok = row < p.M;
      const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + as * BLOCK_N + chalf * 128;
      const int col0 = n_blk * BLOCK_N + chalf * 128;

      if constexpr (EPI == EPI_QKV_ROPE) {
        // 64-column groups = one attention head of q, k or v
        const int t = row_ok ? (row % p.T) : 0;
        const float4* cs4 = reinterpret_cast<const float4*>(p.rope_cos + (size_t)t * 32);
        const float4* sn4 = reinterpret_cast<const float4*>(p.rope_sin + (size_t)t * 32);
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const int col = col0 + g * 64;
          if (col >= p.N) break;  // warp-uniform
          uint32_t lo[32], hi[32];
          tmem_ld_32x32b_x32(taddr0 + g * 64, lo);
          tmem_ld_32x32b_x32(taddr0 + g * 64 + 32, hi);
          tmem_wait_ld();
          const int sect = col / p.E;  // 0 = q, 1 = k, 2 = v (E % 64 == 0 so a group never straddles)
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
          uint32_t out_lo[16], out_hi[16];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bl = __ldg(b4 + j4), bh = __ldg(b4 + 8 + j4);
            float x1[4] = {__uint_as_float(lo[4 * j4 + 0]) + bl.x, __uint_as_float(lo[4 * j4 + 1]) + bl.y,
                           __uint_as_float(lo[4 * j4 + 2]) + bl.z, __uint_as_float(lo[4 * j4 + 3]) + bl.w};
            float x2[4] = {__uint_as_float(hi[4 * j4 + 0]) + bh.x, __uint_as_float(hi[4 * j4 + 1]) + bh.y,
                           __uint_as_float(hi[4 * j4 + 2]) + bh.z, __uint_as_float(hi[4 * j4 + 3]) + bh.w};
            float y1[4], y2[4];
            const float sc = (sect == 0) ? p.q_scale : 1.0f;
            if (sect < 2 && p.rope_cos != nullptr) {
              const float4 c = __ldg(cs4 + j4), s = __ldg(sn4 + j4);
              const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = x1[e] * sc, b = x2[e] * sc;
                // rotary_embedding.py:16-20: x*cos + rotate_half(x)*sin, rotate_half = cat(-x2, x1)
                y1[e] = a * cc[e] - b * ss[e];
                y2[e] = b * cc[e] + a * ss[e];
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) { y1[e] = x1[e] * sc; y2[e] = x2[e] * sc; }
            }
            out_lo[2 * j4 + 0] = pack_half2(y1[0], y1[1]);
            out_lo[2 * j4 + 1] = pack_half2(y1[2], y1[3]);
            out_hi[2 * j4 + 0] = pack_half2(y2[0], y2[1]);
            out_hi[2 * j4 + 1] = pack_half2(y2[2], y2[3]);
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + (size_t)row * p.ldo + col);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              dst[v] = make_uint4(out_lo[4 * v], out_lo[4 * v + 1], out_lo[4 * v + 2], out_lo[4 * v + 3]);
              dst[4 + v] = make_uint4(out_hi[4 * v], out_hi[4 * v + 1], out_hi[4 * v + 2], out_hi[4 * v + 3]);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = col0 + c * 32;
          if (col >= p.N) break;  // warp-uniform (N % 32 == 0)
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr0 + c * 32, acc);
          tmem_wait_ld();
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
          if constexpr (EPI == EPI_BIAS_RESIDUAL) {
            if (row_ok) {
              float4* x4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col);
              float4 xin[8];
#pragma unroll
              for (int v = 0; v < 8; ++v) xin[v] = x4[v];
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                const float4 b = __ldg(b4 + v);
                float4 o;
                o.x = xin[v].x + (__uint_as_float(acc[4 * v + 0]) + b.x);
                o.y = xin[v].y + (__uint_as_float(acc[4 * v + 1]) + b.y);
                o.z = xin[v].z + (__uint_as_float(acc[4 * v + 2]) + b.z);
                o.w = xin[v].w + (__uint_as_float(acc[4 * v + 3]) + b.w);
                x4[v] = o;
              }
            }
          } else if constexpr (EPI == EPI_BIAS_GELU) {
            uint32_t o[16];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const float4 b = __ldg(b4 + v);
              const float g0 = gelu_erf(__uint_as_float(acc[4 * v + 0]) + b.x);
              const float g1 = gelu_erf(__uint_as_float(acc[4 * v + 1]) + b.y);
              const float g2 = gelu_erf(__uint_as_float(acc[4 * v + 2]) + b.z);
              const float g3 = gelu_erf(__uint_as_float(acc[4 * v + 3]) + b.w);
              o[2 * v] = pack_half2(g0, g1);
              o[2 * v + 1] = pack_half2(g2, g3);
            }
            if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + (size_t)row * p.ldo + col);
#pragma unroll
              for (int v = 0; v < 4; ++v) dst[v] = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
            }
          } else {  // EPI_BIAS_F32 / EPI_BIAS_GELU_F32
            if (row_ok) {
              float4* y4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col);
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                const float4 b = __ldg(b4 + v);
                float4 o;
                o.x = __uint_as_float(acc[4 * v + 0]) + b.x;
                o.y = __uint_as_float(acc[4 * v + 1]) + b.y;
                o.z = __uint_as_float(acc[4 * v + 2]) + b.z;
                o.w = __uint_as_float(acc[4 * v + 3]) + b.w;
                if constexpr (EPI == EPI_BIAS_GELU_F32) {
                  o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
                }
                y4[v] = o;
              }
            }
          }
        }
      }
      // all TMEM reads of this thread are complete (wait::ld above) -> hand the accumulator back
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EPI>
inline cudaError_t launch_gemm_epi(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                                   cudaStream_t stream) {
  using namespace gemm_cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_f16_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, p);
  return cudaGetLastError();
}

}  // namespace esmb200
