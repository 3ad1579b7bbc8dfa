// esm_b200 — attention forward v7 (sm_100a, head_dim 64): v4's persistent pipeline with TWO tcgen05.mma issuing threads.
//
// Replaces /root/reference/esm/multihead_attention.py:357-394 (contract: attention_common.cuh).
//
// Measured on B200 (profiles/r01_attention_decomposition.txt): one issuing thread is a serial stream of ~130 cycles per
// tcgen05.mma whatever N is, and v4 issues the 4 Q.K^T and the 4 P.V instructions of every 64-key block from one thread —
// 81 % of v4's time is that chain.  Streams are per thread, so here
//   warp 1 lane 0 issues Q.K^T   (waits: K/V tile landed, P.V(g-2) done      -> commits s_full, kv_empty, q_empty)
//   warp 6 lane 0 issues P.V     (waits: P_g stored, first block: O drained   -> commits pv_done, kv_empty)
// and the two chains run concurrently (4 streams per SM with 2 CTAs).
// RESULT (profiles/r01_attention_decomposition.txt): correct, and the issue chain is gone — with the exponentials
// disabled v7 runs in 0.49 ms where v4 needs 0.63 ms — but with them the softmax warps are now the bottleneck at the same
// ~1560 cycles per block (two CTAs' exponential passes saturate the SM's MUFU pipe, ~80 % of its mixed-instruction rate),
// so the full kernel is only 1 % faster than v4; moving a quarter of the exponentials to the FMA pipe (exp2_fma below)
// and packed FFMA2/FADD2 arithmetic bring it to 0.577 ms (595 TFLOP/s, +8 % over v4).  Superseded as the default by attention8.cuh; ESMB200_ATTN=7 selects it for A/B runs.
// To keep the hand-offs to one barrier per direction, P_g is stored over the first 32 columns of its own S_g buffer
// (every softmax thread has read its whole S row into registers before it writes P), so
//   * Q.K^T(g+3), which overwrites S buffer g%3, is gated by pv_done(g) alone — that also implies S_g was read;
//   * when s_full(g+3) fires the storage of P_g is already dead, the softmax warps need no extra wait before storing P.
// kv_empty takes two arrivals per stage (the K tile is released by Q.K^T, the V tile by P.V).
// Everything else (persistent tiles, Q double buffer, 4-stage K/V ring, lazy reference-max rescale, O in TMEM) is v4's.
//
// S is TRIPLE buffered: Q.K^T(g+3) waits for P.V(g), so the hand-off chain softmax(g) -> P.V(g) -> Q.K^T(g+3) -> softmax(g+3)
// (~1500 cycles measured) has two whole softmax blocks of slack and no longer paces the kernel.
// TMEM (256 columns, 2 CTAs/SM): S0/P0 [0,64) S1/P1 [64,128) S2/P2 [128,192) | O [192,256).
#pragma once

#include "attention_common.cuh"

namespace esmb200 {

namespace attn7_cfg {
using namespace attn4_cfg;
constexpr int NUM_THREADS7 = 224;
}  // namespace attn7_cfg

__global__ void __launch_bounds__(attn7_cfg::NUM_THREADS7, 2)
attention_fwd_kernel_v7(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                        const AttnParams p) {
  using namespace attn4_cfg;
  constexpr float LOG2E = attn_cfg::LOG2E;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // 2 buffers
  uint8_t* smem_k = smem + 2 * Q_BYTES;                     // KV_STAGES buffers
  uint8_t* smem_v = smem + 2 * Q_BYTES + KV_STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * Q_BYTES + KV_STAGES * 2 * KV_BYTES);
  uint64_t* q_full = bars;          // [2] TMA -> QK issuer
  uint64_t* q_empty = bars + 2;     // [2] QK issuer -> TMA (all QK^T of the tile have completed)
  uint64_t* kv_full = bars + 4;     // [4] TMA -> both issuers
  uint64_t* kv_empty = bars + 8;    // [4] both issuers -> TMA (2 arrivals)
  uint64_t* s_full = bars + 12;     // [3] QK issuer -> softmax: S_g written
  uint64_t* p_full = bars + 15;     // [3] softmax -> PV issuer: P_g stored (128 arrivals)
  uint64_t* pv_done = bars + 18;    // [3] PV issuer -> QK issuer / softmax: P_g V_g accumulated
  uint64_t* o_free = bars + 21;     // [1] softmax -> PV issuer: O of the finished tile has been read (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

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
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
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
  pdl_launch_dependents();
  pdl_wait();  // everything below reads the previous kernel's output or writes ctx
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // + 64 * buffer; P_g occupies the first 32 columns of S_g's buffer
  const uint32_t tmem_o = tmem_base + 192;

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
    // ===================== Q.K^T issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 64, false);
      uint32_t g = 0, tq = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nblk = n_blocks(w);
        if (nblk == 0) continue;
        const uint32_t qb = tq & 1;
        mbar_wait(&q_full[qb], (tq >> 1) & 1);
        const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem_q + qb * Q_BYTES), 1024, 0);
        for (int j = 0; j < nblk; ++j, ++g) {
          const uint32_t s = g % KV_STAGES;
          ATRACE(0, g);
          mbar_wait(&kv_full[s], (g / KV_STAGES) & 1);
          const uint32_t sb = g % 3;
          if (g >= 3) mbar_wait(&pv_done[sb], ((g - 3) / 3) & 1);  // P.V(g-3) has consumed the P stored in this buffer
          tc_fence_after();
          ATRACE(1, g);
          const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem_k + s * KV_BYTES), 1024, 0);
          const uint32_t d = tmem_s + sb * 64;
#pragma unroll
          for (int k = 0; k < HEAD_DIM / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
          tc_commit(&s_full[sb]);
          tc_commit(&kv_empty[s]);
          ATRACE(2, g);
        }
        tc_commit(&q_empty[qb]);  // every Q.K^T of this tile has completed
        ++tq;
      }
    }
  } else if (warp == 6) {
    // ===================== P.V issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, true);
      uint32_t g = 0, tp = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nblk = n_blocks(w);
        if (nblk == 0) continue;
        for (int j = 0; j < nblk; ++j, ++g) {
          const uint32_t bf = g % 3;
          const uint32_t s = g % KV_STAGES;
          ATRACE(3, g);
          mbar_wait(&kv_full[s], (g / KV_STAGES) & 1);  // the V tile (long complete: Q.K^T(g) ran before the softmax)
          mbar_wait(&p_full[bf], (g / 3) & 1);
          ATRACE(4, g);
          if (j == 0 && tp > 0) mbar_wait(o_free, (tp - 1) & 1);  // previous tile's O has been read out
          tc_fence_after();
          const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(smem_v + s * KV_BYTES), 1024, 8192);
#pragma unroll
          for (int k = 0; k < BLOCK_KV / 16; ++k)
            umma_ts(tmem_o, tmem_s + bf * 64 + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
          tc_commit(&pv_done[bf]);
          tc_commit(&kv_empty[s]);
          ATRACE(5, g);
        }
        ++tp;
      }
    }
  } else {
    // ===================== softmax / output warps (2-5): one thread per query row =====================
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
        const uint32_t bf = g % 3;
        const uint32_t ph = (g / 3) & 1;
        const uint2 kw2 = __ldg(reinterpret_cast<const uint2*>(kb_ptr + j * 2));
        const uint32_t kw[2] = {kw2.x, kw2.y};
        const uint32_t ts = tmem_s + lane_addr + bf * 64;
        if (warp == 2 && lane == 0) ATRACE(6, g);
        mbar_wait(&s_full[bf], ph);
        tc_fence_after();
        if (warp == 2 && lane == 0) ATRACE(7, g);
        if (j == 0) {  // exact row max of the first block seeds the reference
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t sv[32];
            tmem_ld_32x32b_x32(ts + c * 32, sv);
            tmem_wait_ld_dep(sv);
            const uint32_t wd = kw[c];
#pragma unroll
            for (int i = 0; i < 32; ++i)
              mx[i & 3] = fmaxf(mx[i & 3], (wd == 0xFFFFFFFFu || ((wd >> i) & 1u)) ? __uint_as_float(sv[i]) : -INFINITY);
          }
          const float m0 = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          m_ref = (m0 == -INFINITY) ? 0.f : m0;
        }

        float rsum = 0.f;
        uint32_t pk[2][16];
        for (int trip = 0;; ++trip) {
          // ---- p = exp(s - m_ref) for the 64 keys of this block; row sum; block max
          const float mneg = -m_ref * LOG2E;
          float sum[4] = {0.f, 0.f, 0.f, 0.f};
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          uint32_t sv[2][32];
          tmem_ld_32x32b_x32(ts, sv[0]);
          tmem_ld_32x32b_x32(ts + 32, sv[1]);
          tmem_wait_ld_dep(sv[0]);  // ONE tcgen05.wait::ld retires both loads
          reg_fence(sv[1]);
          if (warp == 2 && lane == 0) ATRACE(8, g);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t wd = kw[c];
            if (wd == 0xFFFFFFFFu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float s0 = __uint_as_float(sv[c][2 * i]), s1 = __uint_as_float(sv[c][2 * i + 1]);
                mx[i & 3] = fmaxf(mx[i & 3], fmaxf(s0, s1));
                float x0, x1, p0, p1;
                fma2(x0, x1, s0, s1, LOG2E, LOG2E, mneg, mneg);  // FFMA2: both keys of the pair in one instruction
#ifdef ESMB200_ATTN_X_NOEXP  // experiment: no MUFU work
                p0 = x0;
                p1 = x1;
#else
                if (ESMB200_ATTN_POLY > 0 && (i % (ESMB200_ATTN_POLY > 0 ? ESMB200_ATTN_POLY : 1)) == 0) {
                  exp2_fma_pair(x0, x1, p0, p1);
                } else {
                  p0 = ex2_approx(x0);
                  p1 = ex2_approx(x1);
                }
#endif
                add2(sum[(i & 1) * 2], sum[(i & 1) * 2 + 1], sum[(i & 1) * 2], sum[(i & 1) * 2 + 1], p0, p1);  // FADD2
                pk[c][i] = pack_half2(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const bool k0 = (wd >> (2 * i)) & 1u, k1 = (wd >> (2 * i + 1)) & 1u;
                const float s0 = k0 ? __uint_as_float(sv[c][2 * i]) : -INFINITY;
                const float s1 = k1 ? __uint_as_float(sv[c][2 * i + 1]) : -INFINITY;
                mx[i & 3] = fmaxf(mx[i & 3], fmaxf(s0, s1));
                const float p0 = ex2_approx(fmaf(s0, LOG2E, mneg));  // ex2(-inf) = 0 for masked keys
                const float p1 = ex2_approx(fmaf(s1, LOG2E, mneg));
                sum[i & 3] += p0 + p1;
                pk[c][i] = pack_half2(p0, p1);
              }
            }
          }
          rsum = (sum[0] + sum[1]) + (sum[2] + sum[3]);
          if (j == 0 || trip == 1) break;
          const float bmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          const bool raise = bmax > m_ref + attn_cfg::RESCALE_TAU;
          if (!__any_sync(0xffffffffu, raise)) break;
          // rare: raise the reference of this warp's rows, rescale O (TMEM) and the row sum, redo the block.
          // S_g is still intact: P_g has not been stored over it yet.
          const float m_new = fmaxf(m_ref, bmax);
          const float alpha = ex2_approx((m_ref - m_new) * LOG2E);
          mbar_wait(&pv_done[(g - 1) % 3], ((g - 1) / 3) & 1);  // P.V(g-1) (and all earlier) have landed in O
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
        // P_g over the first 32 columns of S_g (this thread's own row, already in registers)
        if (warp == 2 && lane == 0) ATRACE(9, g);
        tmem_st_32x32b_x16(ts, pk[0]);
        tmem_st_32x32b_x16(ts + 16, pk[1]);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[bf]);
        l_run += rsum;
      }

      // ---- tile epilogue: O / l -> ctx
      uint32_t outv[32];
      if (nblk > 0) {
        mbar_wait(&pv_done[(g - 1) % 3], ((g - 1) / 3) & 1);
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

inline cudaError_t launch_attention_v7(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const AttnParams& p,
                                       int num_sms, cudaStream_t stream) {
  using namespace attn4_cfg;
  cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel_v7, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const long long total = (long long)p.B * p.H * ((p.T + BLOCK_Q - 1) / BLOCK_Q);
  const int grid = (int)(total < 2LL * num_sms ? total : 2LL * num_sms);
  return launch_pdl(attention_fwd_kernel_v7, dim3(grid), dim3(attn7_cfg::NUM_THREADS7), SMEM_BYTES, stream, tmap_q,
                    tmap_kv, p);
}

}  // namespace esmb200
