// esm_b200 — padding-masked FlashAttention-style forward on tcgen05 / TMEM (sm_100a), head_dim = 64.
//
// Replaces /root/reference/esm/multihead_attention.py:357-394 (bmm(q,k^T) -> key-padding -inf mask ->
// fp32 softmax -> bmm(P,v) -> (T,B,E) merge) without ever writing S or P to HBM.
//
// Inputs come from the QKV GEMM epilogue: qkv fp16 [B*T, 3E], q already scaled by d^-1/2 and rotated,
// k rotated.  One CTA = one (sequence, head, 128-query block); 2 CTAs co-reside per SM so that one CTA's
// MMAs overlap the other's softmax.
//
//   warp 0   : TMA producer  — Q tile once, then K/V tiles (128 keys x 64) through a 2-stage ring
//   warp 1   : MMA issuer    — S = Q K^T  (SS, 128x128x64)  ->  TMEM cols [0,128)
//                              O_j = P V  (TS, 128x64x128, P read from TMEM, V MN-major from smem)
//   warps 2-5: softmax       — one thread per query row: tcgen05.ld S, mask, online max/sum (fp32, exp2),
//                              P -> fp16 -> tcgen05.st over S's own columns, O accumulated in registers
//
// TMEM (256 columns): S fp32 [0,128) aliased by P fp16 [0,64); O_j fp32 double buffer [128,192) / [192,256).
#pragma once

#include "common.cuh"

namespace esmb200 {

struct AttnParams {
  int B, T, H, E;           // E = H * 64
  const uint32_t* keybits;  // [B, words]: bit i of word w set <=> key 32*w+i is attendable (not pad, < T)
  const int* kvlen;         // [B]: 1 + index of the last attendable key (0 if none)
  int words;                // words per sequence, multiple of 4
  __half* ctx;              // [B*T, E] attention output, heads merged (column h*64 + j)
  float* row_max;           // optional [B,H,T]: final softmax row max (of the scaled scores) ...
  float* row_sum;           // optional [B,H,T]: ... and row sum of exp(s - max), for attention_probs_kernel
  int cols = 1;             // attention4.cuh only: sequence s = (s / cols, s % cols) of a [B/cols, T, cols, 3E] tensor
                            // (MSA column attention: the T tokens of a sequence are `cols` rows apart)
};

namespace attn_cfg {
constexpr int BLOCK_Q = 128;
constexpr int BLOCK_KV = 128;
constexpr int HEAD_DIM = 64;
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB, every tile is 128 rows x 128 bytes
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * KV_STAGES) + 1024 + 128;
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace attn_cfg

__global__ void __launch_bounds__(attn_cfg::NUM_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  using namespace attn_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + TILE_BYTES;
  uint8_t* smem_v = smem + TILE_BYTES * (1 + KV_STAGES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TILE_BYTES * (1 + 2 * KV_STAGES));
  uint64_t* q_full = bars;             // [1]
  uint64_t* kv_full = bars + 1;        // [2]
  uint64_t* kv_empty = bars + 3;       // [2]
  uint64_t* s_full = bars + 5;         // [1] MMA -> softmax
  uint64_t* p_full = bars + 6;         // [1] softmax -> MMA
  uint64_t* o_full = bars + 7;         // [2] MMA -> softmax
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const uint32_t warp = threadIdx.x / 32;
  const uint32_t lane = threadIdx.x % 32;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * BLOCK_Q;
  const int kvlen = p.kvlen[b];
  const int nblk = (kvlen + BLOCK_KV - 1) / BLOCK_KV;
  const int row_base = b * p.T;  // first row of this sequence in the [B*T, 3E] matrix

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
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
  const uint32_t tmem_s = tmem_base;        // S fp32 128 cols / P fp16 64 cols
  const uint32_t tmem_o = tmem_base + 128;  // 2 x 64 cols

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
      mbar_wait(q_full, 0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem_k + s * TILE_BYTES), 1024, 0);
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k) umma_ss(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        tc_commit(s_full);
        // P_j (fp16, TMEM) ready <=> every softmax thread has read S_j and stored its P row
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        // V tile: rows = keys (K dim), 64 head-dim elements (N) contiguous -> MN-major, 16 keys = 2048 B per UMMA_K
        const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(smem_v + s * TILE_BYTES), 1024, 16384);
        const uint32_t tmem_oj = tmem_o + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < BLOCK_KV / 16; ++k) umma_ts(tmem_oj, tmem_s + 8 * k, vdesc + 128 * k, idesc_pv, k != 0);
        tc_commit(&o_full[j & 1]);
        tc_commit(&kv_empty[s]);
      }
    }
  } else {
    // ===================== softmax / output warps =====================
    const uint32_t quarter = warp % 4;
    const uint32_t row_local = quarter * 32 + lane;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    const uint32_t* kb_ptr = p.keybits + (size_t)b * p.words;

    for (int j = 0; j < nblk; ++j) {
      const uint4 kw4 = __ldg(reinterpret_cast<const uint4*>(kb_ptr + j * 4));
      const uint32_t kw[4] = {kw4.x, kw4.y, kw4.z, kw4.w};
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // ---- pass 1: row max over the 128 keys of this block
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32b_x32(tmem_s + lane_addr + c * 32, sv);
        tmem_wait_ld();
        const uint32_t w = kw[c];
        if (w == 0xFFFFFFFFu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if ((w >> i) & 1u) mx = fmaxf(mx, __uint_as_float(sv[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // row with no attendable key so far
      const float alpha = ex2_approx((m_run - m_use) * LOG2E);      // exp2(-inf) = 0 on the first block
      const float mneg = -m_use * LOG2E;
      // ---- pass 2: p = exp(s - m), fp32 row sum, fp16 P written over S's columns [16c, 16c+16)
      float rsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32b_x32(tmem_s + lane_addr + c * 32, sv);
        tmem_wait_ld();
        const uint32_t w = kw[c];
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), LOG2E, mneg));
          float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), LOG2E, mneg));
          if (w != 0xFFFFFFFFu) {
            p0 = ((w >> (2 * i)) & 1u) ? p0 : 0.f;
            p1 = ((w >> (2 * i + 1)) & 1u) ? p1 : 0.f;
          }
          rsum += p0 + p1;
          pk[i] = pack_half2(p0, p1);
        }
        tmem_st_32x32b_x16(tmem_s + lane_addr + c * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
      l_run = l_run * alpha + rsum;
      m_run = m_new;
      // ---- fold in the previous block's O while this block's PV runs
      if (j > 0) {
        const int jp = j - 1;
        mbar_wait(&o_full[jp & 1], (jp >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tmem_o + lane_addr + (jp & 1) * 64 + hlf * 32, ov);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[hlf * 32 + i] = fmaf(o_acc[hlf * 32 + i], alpha_prev, __uint_as_float(ov[i]));
        }
      }
      alpha_prev = alpha;
    }
    if (nblk > 0) {
      const int jp = nblk - 1;
      mbar_wait(&o_full[jp & 1], (jp >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(tmem_o + lane_addr + (jp & 1) * 64 + hlf * 32, ov);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[hlf * 32 + i] = fmaf(o_acc[hlf * 32 + i], alpha_prev, __uint_as_float(ov[i]));
      }
    }
    const int t = q0 + row_local;
    if (t < p.T) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      if (p.row_max != nullptr) {
        const size_t si = ((size_t)b * p.H + h) * p.T + t;
        p.row_max[si] = (m_run == -INFINITY) ? 0.f : m_run;
        p.row_sum[si] = l_run;
      }
      uint4* dst = reinterpret_cast<uint4*>(p.ctx + (size_t)(row_base + t) * p.E + h * HEAD_DIM);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        uint4 o;
        o.x = pack_half2(o_acc[8 * v + 0] * inv, o_acc[8 * v + 1] * inv);
        o.y = pack_half2(o_acc[8 * v + 2] * inv, o_acc[8 * v + 3] * inv);
        o.z = pack_half2(o_acc[8 * v + 4] * inv, o_acc[8 * v + 5] * inv);
        o.w = pack_half2(o_acc[8 * v + 6] * inv, o_acc[8 * v + 7] * inv);
        dst[v] = o;
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


// ---------------------------------------------------------------------------------------------------------------
// need_head_weights=True: materialise the normalised probabilities (multihead_attention.py:379,397-400).
// One CTA per (key block, query block, sequence*head): S = Q K^T again on the tensor core, then
// p = exp(s - rowmax) / rowsum with the row statistics saved by attention_fwd_kernel, fp32 [B,H,T,T].
// ---------------------------------------------------------------------------------------------------------------
struct ProbsParams {
  int B, T, H, E;
  const uint32_t* keybits;
  const int* kvlen;
  int words;
  const float* row_max;
  const float* row_sum;
  float* probs;  // [B,H,T,T], batch b starting at probs + b * batch_stride (elements)
  long long batch_stride;
  int zero_pad_rows;  // 1: rows of padded query tokens are written as zeros (ESM2.forward's stacked result)
};

namespace probs_cfg {
constexpr int NUM_THREADS = 128;
constexpr int TMEM_COLS = 128;
constexpr int SMEM_BYTES = 2 * attn_cfg::TILE_BYTES + 1024 + 64 + 4 * 32 * 33 * 4;  // + per-warp transpose tiles
}  // namespace probs_cfg

__global__ void __launch_bounds__(probs_cfg::NUM_THREADS, 4)
attention_probs_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const ProbsParams p) {
  using namespace attn_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_BYTES);
  uint64_t* ld_full = bars;
  uint64_t* mma_done = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const uint32_t warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int kb = blockIdx.x, qb = blockIdx.y;
  const int b = blockIdx.z / p.H, h = blockIdx.z % p.H;
  const int q0 = qb * BLOCK_Q, k0 = kb * BLOCK_KV;
  const int row_base = b * p.T;
  const bool live = k0 < p.kvlen[b];  // otherwise every key of this block is masked: probabilities are exactly 0

  if (threadIdx.x == 0) {
    mbar_init(ld_full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, probs_cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;

  if (live && threadIdx.x == 0) {
    mbar_arrive_expect_tx(ld_full, 2 * TILE_BYTES);
    tma_load_2d(smem_q, &tmap_qkv, ld_full, h * HEAD_DIM, row_base + q0);
    tma_load_2d(smem_k, &tmap_qkv, ld_full, p.E + h * HEAD_DIM, row_base + k0);
    mbar_wait(ld_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128, false);
    const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem_q), 1024, 0);
    const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem_k), 1024, 0);
#pragma unroll
    for (int k = 0; k < HEAD_DIM / 16; ++k) umma_ss(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
    tc_commit(mma_done);
  }
  __syncwarp();

  // Each thread owns one query row in TMEM; the 32x32 fp32 piece of a warp is transposed through padded shared
  // memory so that every global store instruction writes one 128-byte row segment (lane = key column).
  const uint32_t quarter = warp % 4;
  const int t_warp0 = q0 + quarter * 32;       // first query row of this warp
  const int t = t_warp0 + lane;
  const bool row_ok = t < p.T;
  const int ncols = min(BLOCK_KV, p.T - k0);
  float* tile = reinterpret_cast<float*>(smem + 2 * TILE_BYTES + 64) + warp * (32 * 33);
  float* base = p.probs + (size_t)b * p.batch_stride + (size_t)h * p.T * p.T + k0;
  float mneg = 0.f, inv = 0.f;
  uint32_t kw[4] = {0u, 0u, 0u, 0u};
  if (live) {
    const size_t si = ((size_t)b * p.H + h) * p.T + (row_ok ? t : 0);
    mneg = -p.row_max[si] * LOG2E;
    const float l = p.row_sum[si];
    inv = l > 0.f ? 1.0f / l : 0.f;
    // esm2.py:135-139: rows of padded QUERY tokens are zero in the stacked result (padded key columns already are)
    if (p.zero_pad_rows && row_ok && !((p.keybits[(size_t)b * p.words + (t >> 5)] >> (t & 31)) & 1u)) inv = 0.f;
    const uint4 kw4 = __ldg(reinterpret_cast<const uint4*>(p.keybits + (size_t)b * p.words + kb * 4));
    kw[0] = kw4.x; kw[1] = kw4.y; kw[2] = kw4.z; kw[3] = kw4.w;
    mbar_wait(mma_done, 0);
    tc_fence_after();
  }
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    if (c * 32 >= ncols) break;
    if (live) {
      uint32_t sv[32];
      tmem_ld_32x32b_x32(tmem_s + ((quarter * 32u) << 16) + c * 32, sv);
      tmem_wait_ld_dep(sv);
      const uint32_t w = kw[c];
#pragma unroll
      for (int i = 0; i < 32; ++i)
        tile[lane * 33 + i] = ((w >> i) & 1u) ? ex2_approx(fmaf(__uint_as_float(sv[i]), LOG2E, mneg)) * inv : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) tile[lane * 33 + i] = 0.f;
    }
    __syncwarp();
    const int col = c * 32 + lane;
    if (col < ncols) {
      const int nrows = min(32, p.T - t_warp0);
      for (int r = 0; r < nrows; ++r) base[(size_t)(t_warp0 + r) * p.T + col] = tile[r * 33 + lane];
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_s, probs_cfg::TMEM_COLS);
  }
}

inline cudaError_t launch_attention_probs(const CUtensorMap& tmap_qkv, const ProbsParams& p, cudaStream_t stream) {
  using namespace attn_cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         probs_cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.T + BLOCK_KV - 1) / BLOCK_KV, (p.T + BLOCK_Q - 1) / BLOCK_Q, p.B * p.H);
  attention_probs_kernel<<<grid, probs_cfg::NUM_THREADS, probs_cfg::SMEM_BYTES, stream>>>(tmap_qkv, p);
  return cudaGetLastError();
}

inline cudaError_t launch_attention(const CUtensorMap& tmap_qkv, const AttnParams& p, cudaStream_t stream) {
  using namespace attn_cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.T + BLOCK_Q - 1) / BLOCK_Q, p.H, p.B);
  attention_fwd_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmap_qkv, p);
  return cudaGetLastError();
}

}  // namespace esmb200
