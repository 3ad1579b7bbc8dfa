// esm_b200 — attention forward v3 (sm_100a, head_dim 64): fully double-buffered 64-key pipeline.
//
// Replaces /root/reference/esm/multihead_attention.py:357-394 (same contract as attention.cuh / attention2.cuh).
//
// v2 profiling (profiles/r01_ncu_attention_v2.txt) showed the softmax pass itself MUFU-bound (MUFU.EX2 lines carry
// the stall samples, 564 instructions per 128x128 pass) but the XU pipe only ~49 % busy: with a single S buffer per
// CTA, QK^T of block j+1 cannot start before the softmax of block j has read S, so every block pays the tensor-core +
// mbarrier round trip, and the two co-resident CTAs convoy in phase.  v3 removes that dependency inside one CTA:
//
//   key blocks of 64: S_j (128x64 fp32) and P_j (128x64 fp16) are both double buffered in TMEM, so QK^T(j+1) runs
//   while the softmax warps exponentiate block j, and P.V(j) runs while they exponentiate block j+1;
//   K/V tiles (64 keys x 64, 8 KB each) stream through a 4-stage TMA ring (prefetch distance 3 blocks);
//   O accumulates in TMEM with the lazy reference-max rescale of v2 (exact softmax up to fp16 rounding of P).
//
// TMEM (256 columns, 2 CTAs/SM): S0 [0,64) S1 [64,128) | P0 [128,160) P1 [160,192) | O [192,256).
#pragma once

#include "attention.cuh"
#include "attention2.cuh"
#include "common.cuh"

namespace esmb200 {

namespace attn3_cfg {
constexpr int BLOCK_Q = 128;
constexpr int BLOCK_KV = 64;
constexpr int HEAD_DIM = 64;
constexpr int KV_STAGES = 4;
constexpr int Q_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int KV_BYTES = 64 * 64 * 2;   // 8 KB per K tile and per V tile
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * 2 * KV_BYTES + 1024 + 256;
}  // namespace attn3_cfg

__global__ void __launch_bounds__(attn3_cfg::NUM_THREADS, 2)
attention_fwd_kernel_v3(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                        const AttnParams p) {
  using namespace attn3_cfg;
  constexpr float LOG2E = attn_cfg::LOG2E;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + Q_BYTES;
  uint8_t* smem_v = smem + Q_BYTES + KV_STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + KV_STAGES * 2 * KV_BYTES);
  uint64_t* q_full = bars;          // [1]
  uint64_t* kv_full = bars + 1;     // [4]
  uint64_t* kv_empty = bars + 5;    // [4]
  uint64_t* s_full = bars + 9;      // [2] MMA -> softmax: S_j written
  uint64_t* s_free = bars + 11;     // [2] softmax -> MMA: S_j fully read (128 arrivals)
  uint64_t* p_full = bars + 13;     // [2] softmax -> MMA: P_j stored (128 arrivals)
  uint64_t* pv_done = bars + 15;    // [2] MMA -> softmax: P_j V_j accumulated (P_j buffer reusable, O current)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * BLOCK_Q;
  const int kvlen = p.kvlen[b];
  const int nblk = (kvlen + BLOCK_KV - 1) / BLOCK_KV;
  const int row_base = b * p.T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
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
  const uint32_t tmem_s = tmem_base;        // + 64 * buffer
  const uint32_t tmem_p = tmem_base + 128;  // + 32 * buffer
  const uint32_t tmem_o = tmem_base + 192;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && nblk > 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(smem_q, &tmap_q, q_full, h * HEAD_DIM, row_base + q0);
      for (int i = 0; i < nblk; ++i) {
        const int s = i % KV_STAGES;
        mbar_wait(&kv_empty[s], ((i / KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * KV_BYTES);
        tma_load_2d(smem_k + s * KV_BYTES, &tmap_kv, &kv_full[s], p.E + h * HEAD_DIM, row_base + i * BLOCK_KV);
        tma_load_2d(smem_v + s * KV_BYTES, &tmap_kv, &kv_full[s], 2 * p.E + h * HEAD_DIM, row_base + i * BLOCK_KV);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nblk > 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 64, false);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, true);
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem_q), 1024, 0);
      auto issue_qk = [&](int i) {
        const int s = i % KV_STAGES;
        mbar_wait(&kv_full[s], (i / KV_STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem_k + s * KV_BYTES), 1024, 0);
        const uint32_t d = tmem_s + (i & 1) * 64;
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        tc_commit(&s_full[i & 1]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      if (nblk > 1) issue_qk(1);
      for (int j = 0; j < nblk; ++j) {
        const int bf = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&p_full[bf], ph);
        tc_fence_after();
        const int s = j % KV_STAGES;
        // V tile: 64 keys (K dim) x 64 head-dim (N, contiguous) -> MN-major B operand, 16 keys = 2048 B per UMMA_K
        const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(smem_v + s * KV_BYTES), 1024, 8192);
#pragma unroll
        for (int k = 0; k < BLOCK_KV / 16; ++k)
          umma_ts(tmem_o, tmem_p + bf * 32 + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
        tc_commit(&pv_done[bf]);
        tc_commit(&kv_empty[s]);
        if (j + 2 < nblk) {
          mbar_wait(&s_free[bf], ph);  // softmax has read S_j out of buffer bf
          tc_fence_after();
          issue_qk(j + 2);
        }
      }
    }
  } else {
    // ===================== softmax / output warps: one thread per query row =====================
    const uint32_t quarter = warp % 4;
    const uint32_t row_local = quarter * 32 + lane;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    float m_ref = 0.f, l_run = 0.f;
    const uint32_t* kb_ptr = p.keybits + (size_t)b * p.words;

    for (int j = 0; j < nblk; ++j) {
      const int bf = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const uint2 kw2 = __ldg(reinterpret_cast<const uint2*>(kb_ptr + j * 2));
      const uint32_t kw[2] = {kw2.x, kw2.y};
      const uint32_t ts = tmem_s + lane_addr + bf * 64;
      const uint32_t tp = tmem_p + lane_addr + bf * 32;
      mbar_wait(&s_full[bf], ph);
      if (j >= 2) mbar_wait(&pv_done[bf], ph ^ 1);  // P.V(j-2) has finished reading P buffer bf
      tc_fence_after();

      if (j == 0) {  // exact row max of the first block seeds the reference
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32b_x32(ts + c * 32, sv);
          tmem_wait_ld_dep(sv);
          const uint32_t w = kw[c];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            mx[i & 3] = fmaxf(mx[i & 3], (w == 0xFFFFFFFFu || ((w >> i) & 1u)) ? __uint_as_float(sv[i]) : -INFINITY);
        }
        const float m0 = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        m_ref = (m0 == -INFINITY) ? 0.f : m0;
      }

      float rsum = 0.f;
      for (int trip = 0;; ++trip) {
        // ---- p = exp(s - m_ref) for the 64 keys of this block -> fp16 P buffer; row sum; block max
        const float mneg = -m_ref * LOG2E;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        uint32_t sv[2][32];
        tmem_ld_32x32b_x32(ts, sv[0]);
        tmem_ld_32x32b_x32(ts + 32, sv[1]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_wait_ld_dep(sv[c]);
          const uint32_t w = kw[c];
          uint32_t pk[16];
          if (w == 0xFFFFFFFFu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float s0 = __uint_as_float(sv[c][2 * i]), s1 = __uint_as_float(sv[c][2 * i + 1]);
              mx[i & 3] = fmaxf(mx[i & 3], fmaxf(s0, s1));
              const float p0 = ex2_approx(fmaf(s0, LOG2E, mneg));
              const float p1 = ex2_approx(fmaf(s1, LOG2E, mneg));
              sum[i & 3] += p0 + p1;
              pk[i] = pack_half2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const bool k0 = (w >> (2 * i)) & 1u, k1 = (w >> (2 * i + 1)) & 1u;
              const float s0 = k0 ? __uint_as_float(sv[c][2 * i]) : -INFINITY;
              const float s1 = k1 ? __uint_as_float(sv[c][2 * i + 1]) : -INFINITY;
              mx[i & 3] = fmaxf(mx[i & 3], fmaxf(s0, s1));
              const float p0 = ex2_approx(fmaf(s0, LOG2E, mneg));  // ex2(-inf) = 0 for masked keys
              const float p1 = ex2_approx(fmaf(s1, LOG2E, mneg));
              sum[i & 3] += p0 + p1;
              pk[i] = pack_half2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(tp + c * 16, pk);
        }
        rsum = (sum[0] + sum[1]) + (sum[2] + sum[3]);
        if (j == 0 || trip == 1) break;
        const float bmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        const bool raise = bmax > m_ref + attn2_cfg::RESCALE_TAU;
        if (!__any_sync(0xffffffffu, raise)) break;
        // rare: raise the reference of this warp's rows, rescale O (TMEM) and the row sum, redo the block
        const float m_new = fmaxf(m_ref, bmax);
        const float alpha = ex2_approx((m_ref - m_new) * LOG2E);
        mbar_wait(&pv_done[bf ^ 1], ((j - 1) >> 1) & 1);  // P.V(j-1) (and all earlier) have landed in O
        tc_fence_after();
#pragma unroll 1
        for (int q8 = 0; q8 < 4; ++q8) {
          uint32_t ov[16];
          tmem_ld_32x32b_x16(tmem_o + lane_addr + q8 * 16, ov);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st_32x32b_x16(tmem_o + lane_addr + q8 * 16, ov);
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      tc_fence_before();
      mbar_arrive(&s_free[bf]);  // every read of S_j by this thread has completed
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[bf]);
      l_run += rsum;
    }

    const int t = q0 + row_local;
    if (nblk > 0) {
      mbar_wait(&pv_done[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);
      tc_fence_after();
    }
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    uint32_t outv[32];
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      uint32_t ov[32];
      if (nblk > 0) {
        tmem_ld_32x32b_x32(tmem_o + lane_addr + hlf * 32, ov);
        tmem_wait_ld_dep(ov);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i)
        outv[hlf * 16 + i] = pack_half2(__uint_as_float(ov[2 * i]) * inv, __uint_as_float(ov[2 * i + 1]) * inv);
    }
    if (t < p.T) {
      if (p.row_max != nullptr) {
        const size_t si = ((size_t)b * p.H + h) * p.T + t;
        p.row_max[si] = m_ref;
        p.row_sum[si] = l_run;
      }
      uint4* dst = reinterpret_cast<uint4*>(p.ctx + (size_t)(row_base + t) * p.E + h * HEAD_DIM);
#pragma unroll
      for (int v = 0; v < 8; ++v) dst[v] = make_uint4(outv[4 * v], outv[4 * v + 1], outv[4 * v + 2], outv[4 * v + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

inline cudaError_t launch_attention_v3(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const AttnParams& p,
                                       cudaStream_t stream) {
  using namespace attn3_cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(attention_fwd_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.T + BLOCK_Q - 1) / BLOCK_Q, p.H, p.B);
  attention_fwd_kernel_v3<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmap_q, tmap_kv, p);
  return cudaGetLastError();
}

}  // namespace esmb200
