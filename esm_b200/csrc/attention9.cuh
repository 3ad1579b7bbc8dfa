// esm_b200 — attention forward v9 (sm_100a, head_dim <= 64, fp16 operands): the softmax warps never wait for the MMAs.
//
// Replaces /root/reference/esm/multihead_attention.py:357-394 (same contract as attention8.cuh).
//
// Why.  v8 (4 CTAs/SM) runs a strictly serial chain per CTA:  S_j -> softmax_j -> [P.V(j); QK^T(j+1)] -> S_{j+1}.  One
// thread issues the eight tcgen05.mma of the bracket at ~133 cycles each (profiles/r01_attention_decomposition.txt), so
// every softmax warp idles > 1000 cycles per block and the kernel needs 875 cycles per (128 x 64) block and SM where the
// exponentials need 384-512 (MUFU) and the tensor core 256.  The TMEM read path is NOT the limit: the r01 micro-benchmark
// that said 47 B/clk/SM was measuring its own local-memory spills; the clean one (scripts/micro/tmem_bench4.cu,
// profiles/r02_tmem_bench4.txt) reaches 445-910 B/clk/SM and ncu shows smsp__mem_tensor_reads_op_ldt at 4 % of peak.
//
// v9 takes QK^T(j+1) off the softmax warps' critical path:
//   * both 32-column halves of S_j are loaded into registers up front (one tcgen05.wait::ld) and the S buffer is handed
//     back at once (s_free); the MMA thread issues QK^T(j+1) right then, i.e. S_{j+1} is produced while the exponentials
//     of block j run.  With all 64 scores in registers the exact row maximum costs 32 FMNMX3 (hidden under the block's
//     MUFU time), so the softmax is ONE pass against a lazily raised reference (raise when a score exceeds it by 2^8)
//     and S is never re-read;
//   * P_j therefore cannot live in the S columns: it gets its own 32 TMEM columns (a second tcgen05.alloc of 32 next to
//     the 128 of S | O: 160 columns per CTA, three CTAs per SM = 480 of 512); P.V(j) stays a TS-MMA (no shared-memory
//     traffic for P).  p_free (one commit per P.V) tells the softmax warps that P.V(j-1) has read P and — for the raise
//     path — that O holds every block up to j-1;
//   * separate K and V rings (3 stages each), the K loads run one block ahead of the V loads because QK^T(j+1) is now
//     issued a whole softmax pass before P.V(j).
// Three CTAs per SM: 18 warps = 5 on two of the four sub-partitions, so 16384 / (5 * 32) -> 96 registers per thread (at
// 112 only two CTAs are resident: measured 435 vs 670 TF/s for v8).
#pragma once

#include "attention8.cuh"

namespace esmb200 {

namespace attn9_cfg {
constexpr int BLOCK_Q = 128;
constexpr int BLOCK_KV = 64;
constexpr int HEAD_DIM = 64;
constexpr int K_STAGES = 3;
constexpr int V_STAGES = 3;
constexpr int Q_BYTES = 128 * 64 * 2;  // 16 KB
constexpr int KV_BYTES = 64 * 64 * 2;  // 8 KB per K tile and per V tile
constexpr int NUM_THREADS = 192;       // warp 0 TMA, warp 1 MMA issuer + TMEM owner, warps 2-5 softmax (thread = query row)
constexpr int CTAS_PER_SM = 3;
constexpr int TMEM_MAIN = 128;         // S [0,64) | O [64,128)
constexpr int TMEM_P = 32;             // P (fp16, 64 keys) — second allocation
constexpr int SMEM_BYTES = Q_BYTES + (K_STAGES + V_STAGES) * KV_BYTES + 1024 + 256;
constexpr float RAISE_TAU = 8.0f / 1.4426950408889634f;  // raise the reference when a score exceeds it by > 8 in log2 units
}  // namespace attn9_cfg

template <int POLY>
__global__ void __launch_bounds__(attn9_cfg::NUM_THREADS, attn9_cfg::CTAS_PER_SM)
attention_fwd_kernel_v9(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                        const AttnParams p) {
  using namespace attn9_cfg;
  constexpr float LOG2E = attn_cfg::LOG2E;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + Q_BYTES;
  uint8_t* smem_v = smem_k + K_STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + V_STAGES * KV_BYTES);
  uint64_t* q_full = bars;          // [1] TMA -> MMA
  uint64_t* q_empty = bars + 1;     // [1] MMA -> TMA (every QK^T of the tile has completed)
  uint64_t* k_full = bars + 2;      // [3] TMA -> MMA
  uint64_t* k_empty = bars + 5;     // [3] MMA -> TMA (QK^T of the block completed)
  uint64_t* v_full = bars + 8;      // [3]
  uint64_t* v_empty = bars + 11;    // [3] MMA -> TMA (P.V of the block completed)
  uint64_t* s_full = bars + 14;     // [1] MMA -> softmax: S_j written
  uint64_t* s_free = bars + 15;     // [1] softmax -> MMA: S_j is in registers (128 arrivals)
  uint64_t* p_full = bars + 16;     // [1] softmax -> MMA: P_j stored (128 arrivals; first block of a tile: and O read out)
  uint64_t* p_free = bars + 17;     // [1] MMA -> softmax: P.V(j) completed (P may be overwritten, O holds blocks <= j)
  uint64_t* o_full = bars + 18;     // [1] MMA -> softmax: last P.V of the tile accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);  // [2]: S | O allocation, P allocation

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;
  const int nqt = (p.T + BLOCK_Q - 1) / BLOCK_Q;
  const int total = p.B * p.H * nqt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < K_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < V_STAGES; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(p_free, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {  // at most 3 CTAs fit an SM (registers, shared memory): 3 x 160 columns never exhaust the 512
    tmem_alloc(&tmem_slot[0], TMEM_MAIN);
    tmem_alloc(&tmem_slot[1], TMEM_P);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();  // everything below reads the previous kernel's output (qkv, key bits) or writes ctx
  const uint32_t tmem_s = tmem_slot[0];
  const uint32_t tmem_o = tmem_s + 64;
  const uint32_t tmem_p = tmem_slot[1];

  auto n_blocks = [&](int w) -> int {
    const int b = w / (nqt * p.H);
    return (p.kvlen[b] + BLOCK_KV - 1) / BLOCK_KV;
  };

  if (warp == 0) {
    // ===================== TMA producer: Q per tile, K one block ahead of V =====================
    if (lane == 0) {
      uint32_t gk = 0, gv = 0, tq = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nblk = n_blocks(w);
        if (nblk == 0) continue;
        const int qt = w % nqt, h = (w / nqt) % p.H, b = w / (nqt * p.H);
        const int row_base = (b / p.cols) * p.T;
        const int x0 = (b % p.cols) * 3 * p.E + h * HEAD_DIM;
        mbar_wait_relaxed(q_empty, (tq & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, Q_BYTES);
        tma_load_2d(smem_q, &tmap_q, q_full, x0, row_base + qt * BLOCK_Q);
        auto load_k = [&](int i) {
          const uint32_t s = gk % K_STAGES;
          mbar_wait_relaxed(&k_empty[s], ((gk / K_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[s], KV_BYTES);
          tma_load_2d(smem_k + s * KV_BYTES, &tmap_kv, &k_full[s], x0 + p.E, row_base + i * BLOCK_KV);
          ++gk;
        };
        load_k(0);
        for (int i = 0; i < nblk; ++i) {
          if (i + 1 < nblk) load_k(i + 1);
          const uint32_t s = gv % V_STAGES;
          mbar_wait_relaxed(&v_empty[s], ((gv / V_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[s], KV_BYTES);
          tma_load_2d(smem_v + s * KV_BYTES, &tmap_kv, &v_full[s], x0 + 2 * p.E, row_base + i * BLOCK_KV);
          ++gv;
        }
        ++tq;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: QK^T(j+1) as soon as S_j is in registers, P.V(j) when P_j is stored =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 64, false);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, true);
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem_q), 1024, 0);
      uint32_t gk = 0, gv = 0, tq = 0;
      auto issue_qk = [&](bool last) {
        const uint32_t s = gk % K_STAGES;
        mbar_wait(&k_full[s], (gk / K_STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem_k + s * KV_BYTES), 1024, 0);
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k) umma_ss(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        tc_commit(s_full);
        tc_commit(&k_empty[s]);
        if (last) tc_commit(q_empty);  // every QK^T of this tile has been issued: Q may be reloaded when they finish
        ++gk;
      };
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nblk = n_blocks(w);
        if (nblk == 0) continue;
        mbar_wait(q_full, tq & 1);
        tc_fence_after();
        issue_qk(nblk == 1);
        for (int j = 0; j < nblk; ++j, ++gv) {
          mbar_wait(s_free, gv & 1);  // S_j has been read into registers by all 128 rows
          tc_fence_after();
          if (j + 1 < nblk) issue_qk(j + 2 == nblk);
          mbar_wait(p_full, gv & 1);  // P_j stored (and, on the first block of a tile, the previous O read out)
          const uint32_t s = gv % V_STAGES;
          mbar_wait(&v_full[s], (gv / V_STAGES) & 1);
          tc_fence_after();
          const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(smem_v + s * KV_BYTES), 1024, 8192);
#pragma unroll
          for (int k = 0; k < BLOCK_KV / 16; ++k)
            umma_ts(tmem_o, tmem_p + 8 * k, vdesc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
          tc_commit(&v_empty[s]);
          tc_commit(p_free);
          if (j + 1 == nblk) tc_commit(o_full);
        }
        ++tq;
      }
    }
  } else {
    // ===================== softmax / output warps (2-5): one thread per query row =====================
    const uint32_t quarter = warp % 4;
    const uint32_t row_local = quarter * 32 + lane;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    const uint32_t ts = tmem_s + lane_addr;
    const uint32_t tp = tmem_p + lane_addr;
    uint32_t ns = 0, nt = 0;  // blocks / tiles consumed so far (barrier phases)
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int qt = w % nqt, h = (w / nqt) % p.H, b = w / (nqt * p.H);
      const int nblk = n_blocks(w);
      const int row_base = (b / p.cols) * p.T;
      const int t = qt * BLOCK_Q + row_local;
      float m_ref = 0.f, l_run = 0.f;
      bool seeded = false;  // m_ref holds the exact maximum of the first block that has an attendable key
      const uint32_t* kb_ptr = p.keybits + (size_t)b * p.words;

      for (int j = 0; j < nblk; ++j, ++ns) {
        const uint2 kw2 = __ldg(reinterpret_cast<const uint2*>(kb_ptr + j * 2));
        const uint32_t kw[2] = {kw2.x, kw2.y};
        mbar_wait(s_full, ns & 1);
        tc_fence_after();
        uint32_t sv0[32], sv1[32];
        tmem_ld_32x32b_x32(ts, sv0);
        tmem_ld_32x32b_x32(ts + 32, sv1);
        tmem_wait_ld_dep(sv0);
        reg_fence(sv1);
        tc_fence_before();
        mbar_arrive(s_free);  // the S buffer may be overwritten by QK^T(j+1)

        // exact maximum of this row's 64 scores (FMNMX3: 32 instructions, hidden under the MUFU time of the block)
        float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if ((kw[0] & kw[1]) == 0xFFFFFFFFu) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            bm[(i >> 1) & 3] = fmaxf(fmaxf(bm[(i >> 1) & 3], __uint_as_float(sv0[i])), __uint_as_float(sv0[i + 1]));
            bm[(i >> 1) & 3] = fmaxf(fmaxf(bm[(i >> 1) & 3], __uint_as_float(sv1[i])), __uint_as_float(sv1[i + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            bm[i & 3] = fmaxf(bm[i & 3], ((kw[0] >> i) & 1u) ? __uint_as_float(sv0[i]) : -INFINITY);
            bm[i & 3] = fmaxf(bm[i & 3], ((kw[1] >> i) & 1u) ? __uint_as_float(sv1[i]) : -INFINITY);
          }
        }
        const float m_blk = fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3]));
        if (!seeded) {  // uniform over the CTA: the key mask is per sequence
          if ((kw[0] | kw[1]) != 0u) {
            m_ref = m_blk;  // the first block with an attendable key seeds the reference with its exact maximum
            seeded = true;
          }
        } else if (__any_sync(0xffffffffu, m_blk > m_ref + RAISE_TAU)) {
          // rare: a score lies more than tau above the reference (P would exceed 2^8 relative to it).  Raise the reference
          // of the rows that need it, rescale O — once P.V(j-1) has completed — and the row sum.
          const float m_new = (m_blk > m_ref + RAISE_TAU) ? m_blk : m_ref;
          const float alpha = ex2_approx((m_ref - m_new) * LOG2E);
          if (j > 0) {
            mbar_wait(p_free, (ns - 1) & 1);  // P.V(j-1) = completion ns-1; P.V(j) cannot start before this row's P_j
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
          }
          l_run *= alpha;
          m_ref = m_new;
        }

        // P = exp(s - m_ref) in one pass; rounded to fp16 relative to the reference (values <= 2^8 keep 11 bits)
        uint32_t pk0[16], pk1[16];
        const float mneg = -m_ref * LOG2E;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        attn8_exp_half<POLY>(sv0, kw[0], mneg, sum, pk0);
        attn8_exp_half<POLY>(sv1, kw[1], mneg, sum, pk1);
        const float rsum = (sum[0] + sum[1]) + (sum[2] + sum[3]);
        // P_j into its own columns, once P.V of the previous block (of this tile or the last one) has read them
        if (ns > 0) {
          mbar_wait(p_free, (ns - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x32b_x16(tp, pk0);
        tmem_st_32x32b_x16(tp + 16, pk1);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(p_full);
        l_run += rsum;
      }

      // ---- tile epilogue: O / l -> ctx
      float inv = 0.f;
      if (nblk > 0) {
        mbar_wait(o_full, nt & 1);
        ++nt;
        tc_fence_after();
        inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      }
      if (t < p.T && p.row_max != nullptr) {
        const size_t si = ((size_t)b * p.H + h) * p.T + t;
        p.row_max[si] = m_ref;
        p.row_sum[si] = l_run;
      }
      uint32_t outv[32];
      if (nblk > 0) {
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tmem_o + lane_addr + hlf * 32, ov);
          tmem_wait_ld_dep(ov);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            outv[hlf * 16 + i] = pack_half2(__uint_as_float(ov[2 * i]) * inv, __uint_as_float(ov[2 * i + 1]) * inv);
        }
        tc_fence_before();  // O has been read: the arrival on p_full of the next tile's first block orders it
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) outv[i] = 0u;
      }
      if (t < p.T) {
        uint4* dst = reinterpret_cast<uint4*>(p.ctx + ((size_t)(row_base + t) * p.cols + b % p.cols) * (size_t)p.E + h * HEAD_DIM);
#pragma unroll
        for (int v = 0; v < 8; ++v) dst[v] = make_uint4(outv[4 * v], outv[4 * v + 1], outv[4 * v + 2], outv[4 * v + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_s, TMEM_MAIN);
    tmem_dealloc(tmem_p, TMEM_P);
  }
}

template <int POLY>
inline cudaError_t launch_attention_v9_poly(const CUtensorMap& tmap_q, const CUtensorMap& tmap_kv, const AttnParams& p,
                                            int num_sms, cudaStream_t stream) {
  using namespace attn9_cfg;
  auto kern = attention_fwd_kernel_v9<POLY>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const long long total = (long long)p.B * p.H * ((p.T + BLOCK_Q - 1) / BLOCK_Q);
  const long long cap = (long long)CTAS_PER_SM * num_sms;
  const int grid = (int)(total < cap ? total : cap);
  return launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmap_q, tmap_kv, p);
}

}  // namespace esmb200
