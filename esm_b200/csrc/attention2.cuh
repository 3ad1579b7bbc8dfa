// esm_b200 — attention forward v2 (sm_100a, head_dim 64): same contract as attention_fwd_kernel (attention.cuh),
// restructured around the real bottleneck — the 16 ex2/clk/SM MUFU rate of the softmax, not the tensor core.
//
// Replaces /root/reference/esm/multihead_attention.py:357-394.
//
// Changes against v1:
//   * O accumulates in TMEM across key blocks (tcgen05.mma accumulate flag); the per-block TMEM->register O merge
//     (64 FMA + 2 tcgen05.ld per row per block, 64 live registers, spills) is gone;
//   * ONE pass over S per key block: probabilities are taken against a running reference max m_ref that is only
//     raised when a block exceeds it by more than 2^8 (lazy rescale, as in FlashAttention-4); the rare raise
//     rescales O in TMEM and redoes that block, so the result is the exact softmax up to fp16 rounding of P;
//   * P lives in its own TMEM columns, so S is free as soon as the softmax warps have READ it: QK^T of block j+1 is
//     issued while block j's exponentials are still being computed (s_free barrier), taking the QK^T latency off the
//     per-block critical path;
//   * the next 32-column chunk of S is prefetched from TMEM while the current one is exponentiated; 4 independent
//     max/sum chains per thread.
//
// TMEM (256 columns): S fp32 [0,128) | P fp16 [128,192) | O fp32 [192,256).   2 CTAs per SM.
#pragma once

#include "attention.cuh"
#include "common.cuh"

namespace esmb200 {

namespace attn2_cfg {
constexpr float RESCALE_TAU = 8.0f / 1.4426950408889634f;  // raise m_ref only when exp(s - m_ref) could exceed 2^8
}

__global__ void __launch_bounds__(attn_cfg::NUM_THREADS, 2)
attention_fwd_kernel_v2(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  using namespace attn_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + TILE_BYTES;
  uint8_t* smem_v = smem + TILE_BYTES * (1 + KV_STAGES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TILE_BYTES * (1 + 2 * KV_STAGES));
  uint64_t* q_full = bars;        // [1]
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // MMA -> softmax: S_j written
  uint64_t* s_free = bars + 6;    // softmax -> MMA: S_j fully read (128 arrivals)
  uint64_t* p_full = bars + 7;    // softmax -> MMA: P_j stored (128 arrivals)
  uint64_t* o_done = bars + 8;    // MMA -> softmax: P_j V_j accumulated into O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * BLOCK_Q;
  const int kvlen = p.kvlen[b];
  const int nblk = (kvlen + BLOCK_KV - 1) / BLOCK_KV;
  const int row_base = b * p.T;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
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
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_p = tmem_base + 128;
  const uint32_t tmem_o = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0 && nblk > 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(smem_q, &tmap_qkv, q_full, h * HEAD_DIM, row_base + q0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_2d(smem_k + s * TILE_BYTES, &tmap_qkv, &kv_full[s], p.E + h * HEAD_DIM, row_base + j * BLOCK_KV);
        tma_load_2d(smem_v + s * TILE_BYTES, &tmap_qkv, &kv_full[s], 2 * p.E + h * HEAD_DIM,
                    row_base + j * BLOCK_KV);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nblk > 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128, false);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, true);
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem_q), 1024, 0);
      auto issue_qk = [&](int j) {
        const int s = j % KV_STAGES;
        mbar_wait(&kv_full[s], (j / KV_STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem_k + s * TILE_BYTES), 1024, 0);
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k) umma_ss(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          mbar_wait(s_free, j & 1);  // every softmax thread has read S_j
          tc_fence_after();
          issue_qk(j + 1);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const int s = j % KV_STAGES;
        const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(smem_v + s * TILE_BYTES), 1024, 16384);
#pragma unroll
        for (int k = 0; k < BLOCK_KV / 16; ++k)
          umma_ts(tmem_o, tmem_p + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
        tc_commit(o_done);
        tc_commit(&kv_empty[s]);
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
      const uint4 kw4 = __ldg(reinterpret_cast<const uint4*>(kb_ptr + j * 4));
      const uint32_t kw[4] = {kw4.x, kw4.y, kw4.z, kw4.w};
      mbar_wait(s_full, j & 1);
      tc_fence_after();

      // masked row max of S_j over this block (fp32)
      auto row_max_pass = [&]() -> float {
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t sv[32];
          tmem_ld_32x32b_x32(tmem_s + lane_addr + c * 32, sv);
          tmem_wait_ld_dep(sv);
          const uint32_t w = kw[c];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(sv[i]);
            mx[i & 3] = fmaxf(mx[i & 3], (w == 0xFFFFFFFFu || ((w >> i) & 1u)) ? v : -INFINITY);
          }
        }
        return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      };

      // p = exp(s - mref) for the 128 keys of this block -> fp16 P in TMEM; returns row sum, tracks the block max
      auto exp_pass = [&](float mref, float& bmax_out) -> float {
        const float mneg = -mref * LOG2E;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        uint32_t sv[2][32];
        tmem_ld_32x32b_x32(tmem_s + lane_addr, sv[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_wait_ld_dep(sv[c & 1]);
          if (c < 3) tmem_ld_32x32b_x32(tmem_s + lane_addr + (c + 1) * 32, sv[(c + 1) & 1]);  // prefetch
          const uint32_t w = kw[c];
          uint32_t pk[16];
          if (w == 0xFFFFFFFFu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float s0 = __uint_as_float(sv[c & 1][2 * i]), s1 = __uint_as_float(sv[c & 1][2 * i + 1]);
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
              const float s0 = k0 ? __uint_as_float(sv[c & 1][2 * i]) : -INFINITY;
              const float s1 = k1 ? __uint_as_float(sv[c & 1][2 * i + 1]) : -INFINITY;
              mx[i & 3] = fmaxf(mx[i & 3], fmaxf(s0, s1));
              const float p0 = ex2_approx(fmaf(s0, LOG2E, mneg));  // ex2(-inf) = 0 for masked keys
              const float p1 = ex2_approx(fmaf(s1, LOG2E, mneg));
              sum[i & 3] += p0 + p1;
              pk[i] = pack_half2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(tmem_p + lane_addr + c * 16, pk);
        }
        bmax_out = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        return (sum[0] + sum[1]) + (sum[2] + sum[3]);
      };

      float bmax, rsum;
      if (j == 0) {
        const float m0 = row_max_pass();
        m_ref = (m0 == -INFINITY) ? 0.f : m0;
      }
      // one call site for exp_pass (single inlined copy): first trip is speculative against the running reference,
      // an optional second trip follows a raise of the reference
      for (int trip = 0;; ++trip) {
        rsum = exp_pass(m_ref, bmax);
        if (j == 0 || trip == 1) break;
        const bool raise = bmax > m_ref + attn2_cfg::RESCALE_TAU;
        if (!__any_sync(0xffffffffu, raise)) break;
        // rare: raise the reference for this warp's rows, rescale O (TMEM) and the row sum, redo the block
        const float m_new = fmaxf(m_ref, bmax);
        const float alpha = ex2_approx((m_ref - m_new) * LOG2E);
        mbar_wait(o_done, (j - 1) & 1);  // P_{j-1} V_{j-1} has landed in O
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
      mbar_arrive(s_free);  // all reads of S_j by this thread are complete
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
      l_run += rsum;
    }

    const int t = q0 + row_local;
    if (nblk > 0) {
      mbar_wait(o_done, (nblk - 1) & 1);
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

inline cudaError_t launch_attention_v2(const CUtensorMap& tmap_qkv, const AttnParams& p, cudaStream_t stream) {
  using namespace attn_cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(attention_fwd_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.T + BLOCK_Q - 1) / BLOCK_Q, p.H, p.B);
  attention_fwd_kernel_v2<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmap_qkv, p);
  return cudaGetLastError();
}

}  // namespace esmb200
