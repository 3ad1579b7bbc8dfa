// esm_b200 — attention forward v4 (sm_100a, head_dim 64): PERSISTENT version of the v3 pipeline.
//
// Replaces /root/reference/esm/multihead_attention.py:357-394 (same contract as attention.cuh).
//
// The v3 capture (profiles/r01_ncu_attention_v3_B16.txt) showed the softmax warps in their exponential pass only
// ~45 % of the time: a CTA lives for 16 key blocks (~16 K cycles of work) but pays ~18 K cycles of launch, TMEM
// allocation, descriptor fetch, first-tile TMA latency and tear-down.  v4 keeps 2 CTAs per SM resident for the whole
// launch and streams (sequence, head, 128-query tile) work items through them:
//   * TMEM, mbarriers and tensor-map prefetch are set up once per CTA;
//   * the TMA warp runs ahead across tiles (Q double buffered, K/V 4-stage ring), the MMA thread keeps QK^T two
//     64-key blocks ahead of P.V across tile boundaries, so a new tile's S_0/S_1 are ready when the softmax warps
//     finish writing the previous tile's output;
//   * per-tile hand-off: softmax reads O, arrives on o_free, and the first P.V of the next tile (accumulate = 0)
//     waits for it.
// Inside a tile the pipeline is v3's: 64-key blocks, S and P double buffered in TMEM, O accumulated in TMEM with the
// lazy reference-max rescale (exact softmax up to fp16 rounding of P).
//
// TMEM (256 columns, 2 CTAs/SM): S0 [0,64) S1 [64,128) | P0 [128,160) P1 [160,192) | O [192,256).
#pragma once

#include "attention.cuh"
#include "attention2.cuh"
#include "common.cuh"

namespace esmb200 {

#ifdef ESMB200_TRACE
// developer instrumentation (scripts/attn_trace.py): timestamps of CTA 0's softmax warp 2 / MMA thread
__device__ long long g_attn_trace[8192];
#define ATRACE(slot, idx) do { if (blockIdx.x == 0 && (idx) < 400) g_attn_trace[(slot) * 400 + (idx)] = clock64(); } while (0)
#else
#define ATRACE(slot, idx) do { } while (0)
#endif

namespace attn4_cfg {
constexpr int BLOCK_Q = 128;
constexpr int BLOCK_KV = 64;
constexpr int HEAD_DIM = 64;
constexpr int KV_STAGES = 4;
constexpr int Q_BYTES = 128 * 64 * 2;   // 16 KB, double buffered
constexpr int KV_BYTES = 64 * 64 * 2;   // 8 KB per K tile and per V tile
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int SMEM_BYTES = 2 * Q_BYTES + KV_STAGES * 2 * KV_BYTES + 1024 + 256;
}  // namespace attn4_cfg

__global__ void __launch_bounds__(attn4_cfg::NUM_THREADS, 2)
attention_fwd_kernel_v4(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                        const AttnParams p) {
  using namespace attn4_cfg;
  constexpr float LOG2E = attn_cfg::LOG2E;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // 2 buffers
  uint8_t* smem_k = smem + 2 * Q_BYTES;                     // KV_STAGES buffers
  uint8_t* smem_v = smem + 2 * Q_BYTES + KV_STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * Q_BYTES + KV_STAGES * 2 * KV_BYTES);
  uint64_t* q_full = bars;          // [2] TMA -> MMA
  uint64_t* q_empty = bars + 2;     // [2] MMA -> TMA (all QK^T of the tile have completed)
  uint64_t* kv_full = bars + 4;     // [4]
  uint64_t* kv_empty = bars + 8;    // [4]
  uint64_t* s_full = bars + 12;     // [2] MMA -> softmax: S_g written
  uint64_t* s_free = bars + 14;     // [2] softmax -> MMA: S_g fully read (128 arrivals)
  uint64_t* p_full = bars + 16;     // [2] softmax -> MMA: P_g stored (128 arrivals)
  uint64_t* pv_done = bars + 18;    // [2] MMA -> softmax: P_g V_g accumulated
  uint64_t* o_free = bars + 20;     // [1] softmax -> MMA: O of the finished tile has been read (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;
  const int nqt = (p.T + BLOCK_Q - 1) / BLOCK_Q;
  const int total = p.B * p.H * nqt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(o_free, 128);
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

  // tile index -> (sequence, head, query tile); consecutive tiles share K/V (same sequence and head) for L2 reuse
  auto n_blocks = [&](int w) -> int {
    const int b = w / (nqt * p.H);
    return (p.kvlen[b] + BLOCK_KV - 1) / BLOCK_KV;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t g = 0, tq = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nblk = n_blocks(w);
        if (nblk == 0) continue;
        const int qt = w % nqt, h = (w / nqt) % p.H, b = w / (nqt * p.H);
        const int row_base = (b / p.cols) * p.T;
        const int x0 = (b % p.cols) * 3 * p.E + h * HEAD_DIM;
        const uint32_t qb = tq & 1;
        mbar_wait_relaxed(&q_empty[qb], ((tq >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], Q_BYTES);
        tma_load_2d(smem_q + qb * Q_BYTES, &tmap_q, &q_full[qb], x0, row_base + qt * BLOCK_Q);
        for (int i = 0; i < nblk; ++i, ++g) {
          const uint32_t s = g % KV_STAGES;
          mbar_wait_relaxed(&kv_empty[s], ((g / KV_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * KV_BYTES);
          tma_load_2d(smem_k + s * KV_BYTES, &tmap_kv, &kv_full[s], x0 + p.E, row_base + i * BLOCK_KV);
          tma_load_2d(smem_v + s * KV_BYTES, &tmap_kv, &kv_full[s], x0 + 2 * p.E, row_base + i * BLOCK_KV);
        }
        ++tq;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 64, false);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, true);
      // QK^T cursor: runs two blocks ahead of the P.V cursor, across tile boundaries
      int qw = blockIdx.x, q_nblk = 0, qj = 0;
      uint32_t gq = 0, tqq = 0;
      auto seek_q = [&]() {  // position qw on the next non-empty tile
        while (qw < total && (q_nblk = n_blocks(qw)) == 0) qw += gridDim.x;
      };
      seek_q();
      auto issue_next_qk = [&]() {
        const uint32_t qb = tqq & 1;
        if (qj == 0) {
          mbar_wait(&q_full[qb], (tqq >> 1) & 1);
        }
        if (gq >= 2) mbar_wait(&s_free[gq & 1], ((gq - 2) >> 1) & 1);  // softmax has read the block that used this S buffer
        const uint32_t s = gq % KV_STAGES;
        mbar_wait(&kv_full[s], (gq / KV_STAGES) & 1);
        tc_fence_after();
        const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem_q + qb * Q_BYTES), 1024, 0);
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem_k + s * KV_BYTES), 1024, 0);
        const uint32_t d = tmem_s + (gq & 1) * 64;
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        tc_commit(&s_full[gq & 1]);
        ++gq;
        if (++qj == q_nblk) {  // last QK^T of this tile: its completion releases the Q buffer
          tc_commit(&q_empty[qb]);
          qj = 0;
          ++tqq;
          qw += gridDim.x;
          seek_q();
        }
      };
      uint32_t gp = 0, tp = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nblk = n_blocks(w);
        if (nblk == 0) continue;
        for (int j = 0; j < nblk; ++j, ++gp) {
          while (gq < gp + 2 && qw < total) issue_next_qk();
          const uint32_t bf = gp & 1;
          ATRACE(0, gp);
          mbar_wait(&p_full[bf], (gp >> 1) & 1);
          ATRACE(1, gp);
          if (j == 0 && tp > 0) mbar_wait(o_free, (tp - 1) & 1);  // previous tile's O has been read out
          tc_fence_after();
          const uint32_t s = gp % KV_STAGES;
          const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(smem_v + s * KV_BYTES), 1024, 8192);
#pragma unroll
          for (int k = 0; k < BLOCK_KV / 16; ++k)
            umma_ts(tmem_o, tmem_p + bf * 32 + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
          tc_commit(&pv_done[bf]);
          tc_commit(&kv_empty[s]);
          ATRACE(2, gp);
        }
        ++tp;
      }
    }
  } else {
    // ===================== softmax / output warps: one thread per query row =====================
    const uint32_t quarter = warp % 4;
    const uint32_t row_local = quarter * 32 + lane;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    uint32_t g = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int qt = w % nqt, h = (w / nqt) % p.H, b = w / (nqt * p.H);
      const int nblk = n_blocks(w);
      const int row_base = (b / p.cols) * p.T;
      const int t = qt * BLOCK_Q + row_local;
      float m_ref = 0.f, l_run = 0.f;
      const uint32_t* kb_ptr = p.keybits + (size_t)b * p.words;

      for (int j = 0; j < nblk; ++j, ++g) {
        const uint32_t bf = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        const uint2 kw2 = __ldg(reinterpret_cast<const uint2*>(kb_ptr + j * 2));
        const uint32_t kw[2] = {kw2.x, kw2.y};
        const uint32_t ts = tmem_s + lane_addr + bf * 64;
        const uint32_t tp = tmem_p + lane_addr + bf * 32;
        if (warp == 2 && lane == 0) ATRACE(4, g);
        mbar_wait(&s_full[bf], ph);
        if (warp == 2 && lane == 0) ATRACE(5, g);
        // P buffer bf was last read by P.V(g-2). No wait is needed for it: the MMA thread issued P.V(g-2) BEFORE
        // Q.K^T(g), and s_full[g] is a tcgen05.commit placed after Q.K^T(g) — it fires only when every earlier MMA of
        // that thread, P.V(g-2) included, has completed. (A separate pv_done wait here cost ~130 cycles per block:
        // an mbarrier try_wait takes ~100 cycles even when the phase is long complete.)
        tc_fence_after();
        if (warp == 2 && lane == 0) ATRACE(6, g);
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
#ifdef ESMB200_ATTN_X_SKIPALL  // experiment: the softmax warps only do the barrier handshake
        for (int trip = 0; trip < 0; ++trip) {
#else
        for (int trip = 0;; ++trip) {
#endif
          // ---- p = exp(s - m_ref) for the 64 keys of this block -> fp16 P buffer; row sum; block max
          const float mneg = -m_ref * LOG2E;
          float sum[4] = {0.f, 0.f, 0.f, 0.f};
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          uint32_t sv[2][32];
#ifdef ESMB200_ATTN_X_NOLD  // experiment: no TMEM read of S (values made up from the loop counters)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            sv[0][i] = (__float_as_uint(m_ref) & 0x3fffffffu) ^ (uint32_t)(i << 12) ^ g;
            sv[1][i] = (__float_as_uint(m_ref) & 0x3fffffffu) ^ (uint32_t)(i << 11) ^ g;
          }
          reg_fence(sv[0]);
          reg_fence(sv[1]);
#else
          tmem_ld_32x32b_x32(ts, sv[0]);
          tmem_ld_32x32b_x32(ts + 32, sv[1]);
          tmem_wait_ld_dep(sv[0]);  // ONE tcgen05.wait::ld retires both loads (each extra wait slows the MMA pipe)
          reg_fence(sv[1]);
#endif
          if (warp == 2 && lane == 0) ATRACE(7, g);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t w = kw[c];
            uint32_t pk[16];
            if (w == 0xFFFFFFFFu) {
  #pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float s0 = __uint_as_float(sv[c][2 * i]), s1 = __uint_as_float(sv[c][2 * i + 1]);
                mx[i & 3] = fmaxf(mx[i & 3], fmaxf(s0, s1));
#ifdef ESMB200_ATTN_X_NOEXP  // experiment: no MUFU work
                const float p0 = fmaf(s0, LOG2E, mneg);
                const float p1 = fmaf(s1, LOG2E, mneg);
#else
                const float p0 = ex2_approx(fmaf(s0, LOG2E, mneg));
                const float p1 = ex2_approx(fmaf(s1, LOG2E, mneg));
#endif
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
#ifdef ESMB200_ATTN_X_NOST  // experiment: P is not written
            reg_fence16(pk);
#else
            tmem_st_32x32b_x16(tp + c * 16, pk);
#endif
          }
          rsum = (sum[0] + sum[1]) + (sum[2] + sum[3]);
          if (j == 0 || trip == 1) break;
          const float bmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          const bool raise = bmax > m_ref + attn2_cfg::RESCALE_TAU;
          if (!__any_sync(0xffffffffu, raise)) break;
          // rare: raise the reference of this warp's rows, rescale O (TMEM) and the row sum, redo the block
          const float m_new = fmaxf(m_ref, bmax);
          const float alpha = ex2_approx((m_ref - m_new) * LOG2E);
          mbar_wait(&pv_done[bf ^ 1], ((g - 1) >> 1) & 1);  // P.V(g-1) (and all earlier) have landed in O
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

      // ---- tile epilogue: O / l -> ctx
      uint32_t outv[32];
      if (nblk > 0) {
        mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tmem_o + lane_addr + hlf * 32, ov);
          tmem_wait_ld_dep(ov);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            outv[hlf * 16 + i] = pack_half2(__uint_as_float(ov[2 * i]) * inv, __uint_as_float(ov[2 * i + 1]) * inv);
        }
        tc_fence_before();
        mbar_arrive(o_free);  // the next tile's first P.V may overwrite O
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) outv[i] = 0u;
      }
      if (t < p.T) {
        if (p.row_max != nullptr) {
          const size_t si = ((size_t)b * p.H + h) * p.T + t;
          p.row_max[si] = m_ref;
          p.row_sum[si] = l_run;
        }
        uint4* dst = reinterpret_cast<uint4*>(p.ctx + ((size_t)(row_base + t) * p.cols + b % p.cols) * p.E + h * HEAD_DIM);
#pragma unroll
        for (int v = 0; v < 8; ++v) dst[v] = make_uint4(outv[4 * v], outv[4 * v + 1], outv[4 * v + 2], outv[4 * v + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

inline cudaError_t launch_attention_v4(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const AttnParams& p,
                                       int num_sms, cudaStream_t stream) {
  using namespace attn4_cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(attention_fwd_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long total = (long long)p.B * p.H * ((p.T + BLOCK_Q - 1) / BLOCK_Q);
  const int grid = (int)(total < 2LL * num_sms ? total : 2LL * num_sms);
  attention_fwd_kernel_v4<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmap_q, tmap_kv, p);
  return cudaGetLastError();
}

}  // namespace esmb200
